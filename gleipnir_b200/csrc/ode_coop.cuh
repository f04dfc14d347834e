// Cooperative ODE solver for mid-size mass-action networks (4 < NS <= 32 species).
//
// A thread-per-system Rosenbrock solve keeps an NS x NS matrix in thread-local memory (9 KB stack
// frame at NS = 16; measured 145 us per step, 0.2-1.4 s per solve) -- fine for throughput, useless
// for a Metropolis chain that needs hundreds of DEPENDENT solves.  Here a group of GL lanes
// (16 or 32, GL >= NS) owns one system:
//   * lane i holds species i: y_i, the stage vectors and ROW i of M = I/(gamma h) - J in registers;
//   * the network is data, not code (generated tables: reactant slots, rate constants, CSR rows of
//     the stoichiometric matrix), so all lanes run the same instructions: lanes evaluate reaction
//     rates in parallel, then each lane gathers its own row of f and J from shared memory;
//   * LU (no pivoting: M is the identity plus a small multiple of a kinetics Jacobian, the same
//     assumption KPP-style chemistry solvers make; a vanishing/non-finite pivot fails the step) and
//     the four triangular solves of a Kaps-Rentrop step run over warp shuffles with compile-time
//     column loops, so the matrix never leaves registers.
// Same method, coefficients, step control and misfit forms as ode_like.cuh (thread-per-system).
#pragma once
#include "common.cuh"
#include "likelihoods.cuh"
#include "ode_like.cuh"
#include "engine.cuh"
#include "philox.cuh"

namespace gl {

// Net (generated): NS, NP, NR, NOBS, GL; tables
// accessors (inline functions over __device__ const arrays):
//   rx_a(r), rx_b(r) (second reactant or -1), rx_k(r) (rate parameter), rx_sf(r) (double)
//   row_ptr(i), row_rx(z), row_coef(z)               (stoichiometry, CSR by species)
//   init_kind(i) (0 const, 1 parameter), init_val(i) (double), init_par(i)
//   obs_ptr(o), obs_sp(z), obs_coef(z) (double)
template <class Net>
struct CoopOde {
    static constexpr int NS = Net::NS, GL = Net::GL, NR = Net::NR;

    struct Shared {                     // per group, in shared memory
        double y[NS];                   // state broadcast
        double v[NR];                   // reaction rates
        double da[NR], db[NR];          // d v_r / d y_a, d v_r / d y_b
        double J[NS][NS + 1];           // Jacobian rows (+1: bank-conflict padding)
        double p[Net::NP];              // parameter vector of this system
    };

    // rates (and optionally their derivatives) from sh.y; all lanes of the group
    static __device__ __forceinline__ void rates(Shared &sh, int lane, unsigned gmask, bool deriv) {
        for (int r = lane; r < NR; r += GL) {
            const int a = Net::rx_a(r), b = Net::rx_b(r);
            const double c = Net::rx_sf(r) * sh.p[Net::rx_k(r)];
            const double ya = sh.y[a];
            if (b < 0) {
                sh.v[r] = c * ya;
                if (deriv) { sh.da[r] = c; sh.db[r] = 0.0; }
            } else {
                const double yb = sh.y[b];
                sh.v[r] = c * ya * yb;
                if (deriv) {
                    if (a == b) { sh.da[r] = 2.0 * c * ya; sh.db[r] = 0.0; }
                    else { sh.da[r] = c * yb; sh.db[r] = c * ya; }
                }
            }
        }
        __syncwarp(gmask);
    }

    // f_i at the state held in sh.y (rates must be current)
    static __device__ __forceinline__ double rhs_row(const Shared &sh, int lane) {
        double f = 0.0;
        if (lane < NS)
            for (int z = Net::row_ptr(lane); z < Net::row_ptr(lane + 1); ++z)
                f = fma((double)Net::row_coef(z), sh.v[Net::row_rx(z)], f);
        return f;
    }

    static __device__ __forceinline__ void publish(Shared &sh, int lane, unsigned gmask, double yi) {
        __syncwarp(gmask);
        if (lane < NS) sh.y[lane] = yi;
        __syncwarp(gmask);
    }

    // One Kaps-Rentrop step for the whole group.  yi = this lane's species; returns false on a bad pivot.
    static __device__ bool step(Shared &sh, int lane, unsigned gmask, int gbase, double yi, double h, double &ynew,
                                double &yerr) {
        constexpr double GAM = 0.5, A21 = 2.0, A31 = 48.0 / 25.0, A32 = 6.0 / 25.0;
        constexpr double C21 = -8.0, C31 = 372.0 / 25.0, C32 = 12.0 / 5.0, C41 = -112.0 / 125.0,
                         C42 = -54.0 / 125.0, C43 = -2.0 / 5.0;
        constexpr double B1 = 19.0 / 9.0, B2 = 0.5, B3 = 25.0 / 108.0, B4 = 125.0 / 108.0;
        constexpr double E1 = 17.0 / 54.0, E2 = 7.0 / 36.0, E4 = 125.0 / 108.0;
        const double ih = 1.0 / h, diag = 1.0 / (GAM * h);
        // ---- Jacobian rows into shared memory, then M row into registers ----------------------
        publish(sh, lane, gmask, yi);
        rates(sh, lane, gmask, true);
        if (lane < NS) {
#pragma unroll
            for (int q = 0; q < NS; ++q) sh.J[lane][q] = 0.0;
            for (int z = Net::row_ptr(lane); z < Net::row_ptr(lane + 1); ++z) {
                const int r = Net::row_rx(z);
                const double coef = (double)Net::row_coef(z);
                sh.J[lane][Net::rx_a(r)] += coef * sh.da[r];
                if (Net::rx_b(r) >= 0 && Net::rx_b(r) != Net::rx_a(r)) sh.J[lane][Net::rx_b(r)] += coef * sh.db[r];
            }
        }
        const double f0 = rhs_row(sh, lane);
        double M[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) M[q] = lane < NS ? (q == lane ? diag : 0.0) - sh.J[lane][q] : (q == lane ? 1.0 : 0.0);
        // ---- LU without pivoting over shuffles: after column c, rows > c hold l_rc in M[c] ----
        bool ok = true;
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            const double piv = __shfl_sync(gmask, M[c], gbase + c);
            ok = ok && (fabs(piv) > 1e-300) && (piv == piv);
            const double l = M[c] * (1.0 / piv);
            if (lane > c) M[c] = l;
#pragma unroll
            for (int q = c + 1; q < NS; ++q) {
                const double pq = __shfl_sync(gmask, M[q], gbase + c);
                if (lane > c) M[q] = fma(-l, pq, M[q]);
            }
        }
        if (!__all_sync(gmask, ok)) return false;
        double dsel = 1.0;                                // M[lane] with static indexing (keeps M in registers)
#pragma unroll
        for (int q = 0; q < NS; ++q) dsel = q == lane ? M[q] : dsel;
        const double inv_d = 1.0 / dsel;
        auto solve = [&](double b) -> double {
#pragma unroll
            for (int c = 0; c < NS; ++c) {              // forward: L has unit diagonal
                const double bc = __shfl_sync(gmask, b, gbase + c);
                if (lane > c && lane < NS) b = fma(-M[c], bc, b);
            }
#pragma unroll
            for (int c = NS - 1; c >= 0; --c) {         // backward
                if (lane == c) b *= inv_d;
                const double bc = __shfl_sync(gmask, b, gbase + c);
                if (lane < c) b = fma(-M[c], bc, b);
            }
            return b;
        };
        // ---- stages -----------------------------------------------------------------------------
        const double g1 = solve(f0);
        publish(sh, lane, gmask, fma(A21, g1, yi));
        rates(sh, lane, gmask, false);
        const double g2 = solve(fma(C21 * ih, g1, rhs_row(sh, lane)));
        publish(sh, lane, gmask, fma(A32, g2, fma(A31, g1, yi)));
        rates(sh, lane, gmask, false);
        const double f3 = rhs_row(sh, lane);
        const double g3 = solve(fma(C32 * ih, g2, fma(C31 * ih, g1, f3)));
        const double g4 = solve(fma(C43 * ih, g3, fma(C42 * ih, g2, fma(C41 * ih, g1, f3))));
        ynew = yi + B1 * g1 + B2 * g2 + B3 * g3 + B4 * g4;
        yerr = E1 * g1 + E2 * g2 + E4 * g4;
        return true;
    }

    // Log-likelihood of the system whose parameters are in sh.p (group-uniform result).
    static __device__ double solve_loglike(Shared &sh, const double *B, int lane, unsigned gmask, int gbase) {
        const int nt = (int)B[OB_NTIMES], like = (int)B[OB_LIKE], max_steps = (int)B[OB_MAXSTEPS];
        const int nrows = (int)B[OB_NROWS];
        const double rtol = B[OB_RTOL], atol = B[OB_ATOL];
        const double *tspan = B + (int)B[OB_OFF_TSPAN], *obsidx = B + (int)B[OB_OFF_OBSIDX];
        const double *data = B + (int)B[OB_OFF_DATA], *sd = B + (int)B[OB_OFF_SD], *mask = B + (int)B[OB_OFF_MASK];
        double yi = 0.0;
        if (lane < NS) yi = Net::init_kind(lane) ? sh.p[Net::init_par(lane)] : Net::init_val(lane);
        double acc = 0.0;
        auto misfit = [&](int it) {                      // lane 0 accumulates; sh.y must hold the state
            publish(sh, lane, gmask, yi);
            if (lane == 0) {
                for (int r = 0; r < nrows; ++r) {
                    if (mask[r * nt + it] == 0.0) continue;
                    const int o = (int)obsidx[r];
                    double sim = 0.0;
                    for (int z = Net::obs_ptr(o); z < Net::obs_ptr(o + 1); ++z) sim = fma(Net::obs_coef(z), sh.y[Net::obs_sp(z)], sim);
                    const double mu = data[r * nt + it];
                    if (like == ODE_SNLPDF) {
                        const double s = sd[r * nt + it], z2 = (sim - mu) / s;
                        acc += -0.5 * z2 * z2 - log(s) - GL_HALF_LOG_2PI;
                    } else {
                        const double e = mu - sim;
                        acc -= like == ODE_SSE ? e * e : e * e * sd[r * nt + it];
                    }
                }
            }
        };
        misfit(0);
        double t = tspan[0];
        double h = fmin(1e-3 * (tspan[nt - 1] - tspan[0]), tspan[1] - tspan[0]);
        int steps = 0;
        bool failed = false;
        for (int it = 1; it < nt && !failed; ++it) {
            const double tn = tspan[it];
            while (t < tn) {                             // every quantity below is group-uniform
                const double rem = tn - t;
                const bool last = h >= rem;
                const double hh = last ? rem : h;
                if (++steps > max_steps) { failed = true; break; }
                double ynew, yerr;
                if (!step(sh, lane, gmask, gbase, yi, hh, ynew, yerr)) {
                    if (hh < 1e-300) { failed = true; break; }
                    h = hh * 0.1;
                    continue;
                }
                double e = 0.0;
                bool fin = true;
                if (lane < NS) {
                    e = fabs(yerr) / (atol + rtol * fmax(fabs(yi), fabs(ynew)));
                    fin = isfinite(ynew) && e == e;
                }
#pragma unroll
                for (int o = GL / 2; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(gmask, e, o));
                fin = __all_sync(gmask, fin);
                if (!fin) {
                    if (hh < 1e-300) { failed = true; break; }
                    h = hh * 0.1;
                    continue;
                }
                const double fac = e > 0.0 ? fmin(5.0, fmax(0.2, 0.9 * pow(e, -0.25))) : 5.0;
                if (e <= 1.0) {
                    yi = ynew;
                    t = last ? tn : t + hh;
                    if (!last || fac < 1.0) h = hh * fac;
                    else h = fmax(h, hh * fac);
                } else {
                    h = hh * fac;
                    if (h < 1e-14 * fabs(tn)) { failed = true; break; }
                }
            }
            if (!failed) misfit(it);
        }
        acc = __shfl_sync(gmask, acc, gbase);
        return (failed || acc != acc) ? -INFINITY : acc;
    }

    // params[mask] = 10**Y with component jrep of Y replaced by v (jrep < 0: none); Y in shared memory
    static __device__ __forceinline__ void load_params(Shared &sh, const double *B, const double *Y, int stride, int d,
                                                       int jrep, double v, int lane, unsigned gmask) {
        const int np = (int)B[OB_NPARAMS];
        const double *nominal = B + (int)B[OB_OFF_NOMINAL], *map = B + (int)B[OB_OFF_MAP];
        const double *islog = B + (int)B[OB_OFF_ISLOG];
        __syncwarp(gmask);
        for (int i = lane; i < np; i += GL) sh.p[i] = nominal[i];
        __syncwarp(gmask);
        for (int j = lane; j < d; j += GL) {
            const double yj = j == jrep ? v : Y[j * stride];
            sh.p[(int)map[j]] = islog[j] != 0.0 ? exp10(yj) : yj;
        }
        __syncwarp(gmask);
    }
};

// ---- batch evaluation: one group per parameter set ------------------------------------------------
template <class Net>
__global__ void __launch_bounds__(128) ode_coop_eval_kernel(ModelDev M, const double *__restrict__ x, long long n,
                                                            double *__restrict__ out) {
    using C = CoopOde<Net>;
    constexpr int GL = Net::GL, GPB = 128 / GL;
    __shared__ typename C::Shared sh[GPB];
    const int g = threadIdx.x / GL, lane = threadIdx.x % GL;
    const int gbase = (threadIdx.x & 31) / GL * GL;
    const unsigned gmask = (GL == 32 ? 0xffffffffu : ((1u << GL) - 1u)) << gbase;
    for (long long i = (long long)blockIdx.x * GPB + g; i < n; i += (long long)gridDim.x * GPB) {
        C::load_params(sh[g], M.aux0, x + i * M.d, 1, M.d, -1, 0.0, lane, gmask);
        const double L = C::solve_loglike(sh[g], M.aux0, lane, gmask, gbase);
        if (lane == 0) out[i] = L;
    }
}


// ---- walkers with a cooperative likelihood ---------------------------------------------------------
// Same chain as walk_kernel (walker.cuh) -- same Philox counters, proposal, prior test, hard wall,
// phases and pooled width adaptation -- but the GL lanes of a group evaluate ONE trial move together
// instead of speculating on GL different ones.
template <class Net>
__global__ void __launch_bounds__(128) walk_coop_kernel(WalkParams P, ModelDev M, int phase) {
    using C = CoopOde<Net>;
    constexpr int GL = Net::GL, WPB = 128 / GL;
    __shared__ typename C::Shared sh[WPB];
    extern __shared__ double dyn[];
    EngineState *S = P.S;
    if (!S->do_walk) return;
    const int k = S->k_cur, d = P.d;
    if (blockIdx.x * WPB >= k) return;
    double *xs = dyn;                              // [d][WPB]
    double *wj = xs + (size_t)d * WPB;
    double *p0s = wj + d, *p1s = p0s + d;
    int *kinds = reinterpret_cast<int *>(p1s + d);
    unsigned int *acc = reinterpret_cast<unsigned int *>(kinds + d);
    const int tid = threadIdx.x;
    const bool tuning = phase < P.tuning_cycles;
    for (int j = tid; j < d; j += 128) {
        wj[j] = phase_width(S, j, phase, k);
        p0s[j] = P.priors->p0[j];
        p1s[j] = P.priors->p1[j];
        kinds[j] = P.priors->kind[j];
        acc[j] = 0u;
    }
    __syncthreads();
    const int wl = tid / GL, lane = tid % GL;
    const int w = blockIdx.x * WPB + wl;
    const bool active = w < k;
    const int gbase = (tid & 31) / GL * GL;
    const unsigned gmask = (GL == 32 ? 0xffffffffu : ((1u << GL) - 1u)) << gbase;
    double *x = xs + wl;
    const double lstar = S->lstar;
    const uint32_t step = (uint32_t)S->step;
    if (active) {
        const long long slot = P.dead_list[w];
        const uint32_t g = (uint32_t)(slot + P.row_offset);
        double L;
        if (phase == 0 || !tuning) {
            long long pick = 0;
            if (lane == 0) {
                for (uint32_t a = 0;; ++a) {
                    const U2 u = philox_uniforms(P.seed, g, step, a, GL_TAG_PICK);
                    pick = (long long)(u.a * (double)P.N);
                    if (pick >= P.N) pick = P.N - 1;
                    if (!P.is_dead[pick]) break;
                }
            }
            pick = __shfl_sync(gmask, pick, gbase);
            for (int j = lane; j < d; j += GL) x[j * WPB] = P.live[pick * d + j];
            L = P.logl[pick];
        } else {
            for (int j = lane; j < d; j += GL) x[j * WPB] = P.scratch_x[(size_t)w * d + j];
            L = P.scratch_l[w];
        }
        __syncwarp(gmask);
        const int n_moves = (tuning ? 20 : P.iterations + P.burn_in) * d;
        const uint32_t c2_base = (uint32_t)((tuning ? phase * 20 : P.tuning_cycles * 20) * d);
        const bool normal = P.proposal == GL_PROPOSAL_NORMAL;
        for (int m = 0; m < n_moves; ++m) {
            const int j = m % d;
            const U2 ru = philox_uniforms(P.seed, g, step, c2_base + (uint32_t)m, GL_TAG_MOVE);
            const double width = wj[j], xj = x[j * WPB];
            const double stepv = __dmul_rn(normal ? normcdfinv(ru.a) : ru.a, width);
            const double nj = __dadd_rn(stepv, normal ? xj : __dadd_rn(xj, -0.5 * width));
            if (prior_accept(kinds[j], p0s[j], p1s[j], xj, nj, ru.b)) {        // group-uniform
                C::load_params(sh[wl], M.aux0, x, WPB, d, j, nj, lane, gmask);
                const double Lt = C::solve_loglike(sh[wl], M.aux0, lane, gmask, gbase);
                if (Lt > lstar) {
                    __syncwarp(gmask);
                    if (lane == 0) {
                        x[j * WPB] = nj;
                        if (tuning) atomicAdd(&acc[j], 1u);
                    }
                    L = Lt;
                    __syncwarp(gmask);
                }
            }
        }
        if (tuning) {
            if (phase + 1 < P.tuning_cycles) {
                for (int j = lane; j < d; j += GL) P.scratch_x[(size_t)w * d + j] = x[j * WPB];
                if (lane == 0) P.scratch_l[w] = L;
            }
        } else {
            for (int j = lane; j < d; j += GL) P.live[slot * d + j] = x[j * WPB];
            if (lane == 0) P.logl[slot] = L;
        }
    }
    if (tuning) {
        __syncthreads();
        for (int j = tid; j < d; j += 128)
            if (acc[j]) atomicAdd(&S->acc[phase][j], acc[j]);
    }
}

template <class Net>
inline int walk_coop_launch(const WalkParams &P, const ModelDev &M, cudaStream_t st) {
    constexpr int WPB = 128 / Net::GL;
    const int smem = (int)(sizeof(double) * ((size_t)P.d * WPB + 3 * (size_t)P.d) + 2 * sizeof(int) * P.d + 16);
    const int blocks = (P.k_max + WPB - 1) / WPB;
    for (int phase = 0; phase <= P.tuning_cycles; ++phase)
        walk_coop_kernel<Net><<<blocks, 128, smem, st>>>(P, M, phase);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

template <class Net>
inline int ode_coop_eval_launch(const ModelDev &M, const double *x, long long n, double *out, cudaStream_t st) {
    constexpr int GPB = 128 / Net::GL;
    long long blocks = (n + GPB - 1) / GPB;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    ode_coop_eval_kernel<Net><<<(int)blocks, 128, 0, st>>>(M, x, n, out);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

}  // namespace gl
