// Cooperative ODE solver for mass-action networks above 4 species, and the multi-condition
// likelihood protocol on top of it.
//
// A thread-per-system Rosenbrock solve keeps an NS x NS matrix in thread-local memory (9 KB stack
// frame at NS = 16; measured 145 us per step, 0.2-1.4 s per solve) -- fine for throughput, useless
// for a Metropolis chain that needs hundreds of DEPENDENT solves.  Here a group of GL lanes
// (16 or 32) owns one system:
//   * lane i holds CORE species i: y_i, the stage vectors and ROW i of M = I/(gamma h) - J in
//     registers.  A network with more species than lanes keeps NX = NS - 32 mutually independent
//     low-degree LEAF species in a second register slot of lanes 0..NX-1; their rows stay in shared
//     memory and are eliminated first (block elimination), so the register matrix stays 32 wide;
//   * species order, the leaf set, the symbolic LU (with fill) and the level schedules of the
//     triangular solves are decided ONCE per network by gleipnir_b200/odeplan.py and emitted as
//     straight-line code (Net::lu, Net::solve): one broadcast + one FMA per structural nonzero of a
//     pivot row, nothing for structural zeros; pivot rows and substitution vectors are broadcast
//     through shared memory (128-bit stores by the owning lane, broadcast loads by all);
//   * reaction rates (and their partial derivatives) are evaluated by all lanes in parallel into signed
//     value tables in shared memory; each lane then sums its f_i from an index list and gathers its row
//     of M straight into registers, column by column (Net::load_m) -- no Jacobian rows in shared memory
//     except those of the leaf species, so a block of four 34-species systems needs 40 KB and four
//     blocks (16 warps) are resident per SM at 128 registers;
//   * error control: weighted RMS norm (the reference solver's), step-size factor in single precision;
//   * no pivoting: M is the identity plus a small multiple of a kinetics Jacobian, the same
//     assumption KPP-style chemistry solvers make; a vanishing/non-finite pivot fails the step
//     (the step is retried with h/10).
// Method: RODAS4 (6 stages, order 4(3), L-stable); step control and misfit forms as in ode_like.cuh
// (thread-per-system, NS <= 4, Kaps-Rentrop).
//
// Protocol (blob layout ProtoBlob, written by odegen.py::pack_protocol_blob) -- the structure of
// examples/pysb_jarm/run_jarm_multinest.py:92-143, of which the single-condition likelihoods of
// pysb_utilities/nestedsample_it.py:273-339 are the special case "one condition, no
// pre-equilibration, no cycle terms":
//     params[mask] = 10**Y                                                       (:94)
//     cycle products scored by norm(loc, scale).logpdf, NaN -> -inf              (:72-90, 101-103)
//     per condition: parameter overrides                                          (:110-111, 122-123)
//         pre-equilibration to t_end with the listed constants zeroed; failure -> -inf   (:105-119)
//         main solve from the equilibrated state, constants restored             (:125)
//         rows: norm / halfnorm logpdf of observable / scale at the masked times (:126-134)
//     sum; NaN -> -inf                                                            (:137-143)
// Conditions are independent solves: the walker gives each its own lane group.
#pragma once
#include "common.cuh"
#include "likelihoods.cuh"
#include "engine.cuh"
#include "philox.cuh"

// resident 128-thread blocks per SM the kernels are compiled for (register budget 65536 / (128 * blocks));
// the generator sets it per network (odegen.py)
#ifndef GL_COOP_MIN_BLOCKS
#define GL_COOP_MIN_BLOCKS 3
#endif

namespace gl {

enum ProtoBlob {
    PB_NSAMPLED = 0, PB_NPARAMS = 1, PB_NTIMES = 2, PB_LIKE = 3, PB_MAXSTEPS = 4, PB_RTOL = 5, PB_ATOL = 6,
    PB_NROWS = 7, PB_NCOND = 8, PB_PREEQ = 9, PB_PREEQ_TEND = 10, PB_PREEQ_TOL = 11, PB_NZERO = 12,
    PB_NCYCLES = 13, PB_CYC_LOC = 14, PB_CYC_SCALE = 15,
    PB_OFF_NOMINAL = 16, PB_OFF_MAP = 17, PB_OFF_ISLOG = 18, PB_OFF_TSPAN = 19, PB_OFF_ROWMETA = 20,
    PB_OFF_DATA = 21, PB_OFF_SD = 22, PB_OFF_MASK = 23, PB_OFF_CTERM = 24, PB_OFF_ROWPTR = 25,
    PB_OFF_OVERPTR = 26, PB_OFF_OVERIDX = 27, PB_OFF_OVERVAL = 28, PB_OFF_ZEROIDX = 29, PB_OFF_CYCPTR = 30,
    PB_OFF_CYCFAC = 31, PB_OFF_STOPPTR = 32, PB_OFF_STOPS = 33, PB_HEADER = 48
};
#ifndef GL_ODE_LIKE_TYPES
#define GL_ODE_LIKE_TYPES
enum OdeLikeType { ODE_SNLPDF = 0, ODE_MSE = 1, ODE_SSE = 2 };
#endif
enum OdeRowKind { ODE_ROW_NORM = 0, ODE_ROW_HALFNORM = 1 };

// RODAS4 (Hairer & Wanner, Solving ODEs II, sec. IV.10; the coefficient set of their RODAS code): 6 stages,
// order 4 with an embedded order-3 solution, stiffly accurate and L-stable -- on the JARM network it needs
// 3.4-4.3x fewer steps than the 4-stage Kaps-Rentrop pair of ode_like.cuh at the same tolerance and is more
// accurate (tools/ode_method_study.py).  Row s of ROD_A / ROD_C: coefficients of k_0..k_{s-1}; the sixth
// stage's state is the fifth's plus k_4, its result is the error estimate.
#define ROD_GAMMA 0.25
__constant__ double ROD_A[30] = {
    0.0, 0.0, 0.0, 0.0, 0.0,
    0.1544000000000000e+01, 0.0, 0.0, 0.0, 0.0,
    0.9466785280815826e+00, 0.2557011698983284e+00, 0.0, 0.0, 0.0,
    0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00, 0.0, 0.0,
    0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 0.0,
    0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0};
__constant__ double ROD_C[30] = {
    0.0, 0.0, 0.0, 0.0, 0.0,
    -0.5668800000000000e+01, 0.0, 0.0, 0.0, 0.0,
    -0.2430093356833875e+01, -0.2063599157091915e+00, 0.0, 0.0, 0.0,
    -0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02, 0.0, 0.0,
    0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0.0,
    0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
    -0.6058818238834054e+01};

// ---- helpers of the generated elimination code ------------------------------------------------------
// 64-bit shuffle from two 32-bit ones without the register-pair moves of the toolkit's wrapper
__device__ __forceinline__ double shfl_f64(unsigned mask, double v, int src) {
    const int lo = __shfl_sync(mask, __double2loint(v), src), hi = __shfl_sync(mask, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}
// b - m*t where lane > c (lane < c), else b: ONE predicated FMA
__device__ __forceinline__ double pfma_gt(int lane, int c, double m, double t, double b) {
    asm("{\n\t.reg .pred p;\n\t.reg .f64 nm;\n\tsetp.gt.s32 p, %1, %2;\n\tneg.f64 nm, %3;\n\t@p fma.rn.f64 %0, nm, %4, %0;\n\t}"
        : "+d"(b) : "r"(lane), "r"(c), "d"(m), "d"(t));
    return b;
}
__device__ __forceinline__ double pfma_lt(int lane, int c, double m, double t, double b) {
    asm("{\n\t.reg .pred p;\n\t.reg .f64 nm;\n\tsetp.lt.s32 p, %1, %2;\n\tneg.f64 nm, %3;\n\t@p fma.rn.f64 %0, nm, %4, %0;\n\t}"
        : "+d"(b) : "r"(lane), "r"(c), "d"(m), "d"(t));
    return b;
}
// 1/x from the hardware's 20-bit seed (MUFU.RCP64H) and two Newton steps: ~1 ulp, no slow path, half the
// instructions of the IEEE division -- the elimination takes one reciprocal per pivot on its critical path
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// b * s on the lanes of `mask`, else b
__device__ __forceinline__ double pmul_mask(int lane, unsigned mask, double b, double s) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tshr.b32 t, %2, %1;\n\tand.b32 t, t, 1;\n\tsetp.ne.b32 p, t, 0;\n\t@p mul.rn.f64 %0, %0, %3;\n\t}"
        : "+d"(b) : "r"(lane), "r"(mask), "d"(s));
    return b;
}

// work counters of a plugin (read through gl_plugin_counters): likelihood evaluations actually
// integrated (prior-rejected trial moves are not), ODE solves, Rosenbrock steps attempted
struct CoopCounters {
    unsigned long long evals, solves, steps;
    unsigned long long next;          // work-claim cursor of the batch-evaluation kernel (zeroed before a launch)
    // optional trace of the walker launches (gl_plugin_trace): one record per block
    unsigned long long trace_on, trace_n;
    unsigned long long trace[4 * 4096];   // (launch phase << 32 | SM id, start ns, end ns, ODE steps of the block)
};
constexpr unsigned long long COOP_TRACE_CAP = 4096;
__device__ __forceinline__ unsigned long long coop_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int coop_smid() {
    unsigned int s;
    asm("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}

// Net (generated): NS, NC (core rows), NX (leaves), NXA = max(NX,1), NP, NR, NOBS, GL, LD (row stride of the leaf
// rows of J), NRM = ceil(NR/GL), NRP = NRM*GL, EF/EF2/ED/EM/EJ2 (list depths);
// all species indices in the plan's INTERNAL numbering (core 0..NC-1, leaves NC..NS-1);
//   rx_k(r) (rate parameter), rx_sf(r) (double)
//   rtab(r)   = a | b << 16: reactant slots of reaction r in y[] (y[NS] = 1.0 stands in for "no second reactant")
//   ftab(e*GL + i)            position in sv[] of the e-th term of f_i (padding: the zero slot 2*NRP)
//   dtab(e*GL + i)            position in sd[] of the e-th term of -J_ii (padding: the zero slot 4*NRP)
//   mtab(z*GL + i)            column lists of M = diag*I - J: load_m() sums sd[mtab(..)] per column into the register
//                             row; the diagonal entry is the lane's own slot sd[4*NRP + 2 + i]
//   ftab2                     ftab for the leaf species (width NXA)
//   jtab2(e*NXA + l)          leaf ROWS (kept in shared memory): position in sd[] | column << 16 | "first term of
//                             this J_l,column" << 24; every entry stores the running sum
//   load_m(M, A, sd, mtab, lane)
//   init_kind(i) (0 const, 1 parameter), init_val(i) (double), init_par(i)
//   obs_ptr(o), obs_sp(z), obs_coef(z) (double)
//   lu(M, A, J, diag, lane, gmask, gbase, inv_d, inv_d2, xs) -> pivots ok;  solve(M, A, J, inv_d, inv_d2, b, b2, ...)
template <class Net>
struct CoopOde {
    static constexpr int NS = Net::NS, NC = Net::NC, NX = Net::NX, NXA = Net::NXA, GL = Net::GL, NR = Net::NR;
    static constexpr int LD = Net::LD, NRM = Net::NRM, NRP = Net::NRP, XS = Net::XS;
    static constexpr int DG = 4 * NRP + 2;              // sd[DG + lane] = diag - J_ii of the lane's species
    static constexpr int SD = (DG + GL > 2 * XS ? DG + GL : 2 * XS);

    struct alignas(16) Tables {         // per block, in shared memory: the network's index lists
        unsigned int rtab[NRP];
        unsigned int jtab2[Net::EJ2 * NXA];
        unsigned short ftab[Net::EF * GL];
        unsigned short mtab[Net::EM * GL];
        unsigned short dtab[Net::ED * GL];
        unsigned short ftab2[Net::EF2 * NXA];
    };
    struct alignas(16) Shared {         // per group, in shared memory
        double J[NXA * LD];             // Jacobian rows of the LEAF species; entries outside the pattern stay 0
        double sv[2 * NRP + 2];         // v_r, -v_r; [2*NRP] = 0
        double sd[SD];                  // dv_r/dy_a, -, dv_r/dy_b, -; [4*NRP] = 0; [DG + lane] = diag - J_ii.  Dead once
                                        // the matrix row is in registers: the substitutions use it as their two
                                        // broadcast buffers
        double c[NRP];                  // sf_r * p[k_r] of the running solve (0 for the padding reactions)
        double k[6 * NS];               // stage vectors of the running step
        double y[NS + 2];               // state broadcast; y[NS] = 1
        double p[Net::NP];              // parameter vector of the running solve (overrides applied)
        double pb[Net::NP];             // parameter vector of the evaluation (params[mask] = 10**Y)
        unsigned long long n_solves, n_steps;
    };

    static __device__ __forceinline__ void load_tables(Tables &T) {         // whole block
        for (int i = threadIdx.x; i < NRP; i += blockDim.x) T.rtab[i] = Net::rtab(i);
        for (int i = threadIdx.x; i < Net::EM * GL; i += blockDim.x) T.mtab[i] = Net::mtab(i);
        for (int i = threadIdx.x; i < Net::ED * GL; i += blockDim.x) T.dtab[i] = Net::dtab(i);
        for (int i = threadIdx.x; i < Net::EJ2 * NXA; i += blockDim.x) T.jtab2[i] = Net::jtab2(i);
        for (int i = threadIdx.x; i < Net::EF * GL; i += blockDim.x) T.ftab[i] = Net::ftab(i);
        for (int i = threadIdx.x; i < Net::EF2 * NXA; i += blockDim.x) T.ftab2[i] = Net::ftab2(i);
    }
    static __device__ __forceinline__ void init_shared(Shared &sh, int lane, unsigned gmask) {   // one group
        for (int i = lane; i < NXA * LD; i += GL) sh.J[i] = 0.0;
        if (lane == 0) {
            sh.sv[2 * NRP] = sh.sv[2 * NRP + 1] = 0.0;
            sh.sd[4 * NRP] = sh.sd[4 * NRP + 1] = 0.0;
            sh.y[NS] = 1.0;
            sh.y[NS + 1] = 0.0;
            sh.n_solves = sh.n_steps = 0ull;
        }
        __syncwarp(gmask);
    }

    static __device__ __forceinline__ void rate_constants(Shared &sh, int lane, unsigned gmask) {
        __syncwarp(gmask);
        for (int r = lane; r < NRP; r += GL) sh.c[r] = r < NR ? Net::rx_sf(r) * sh.p[Net::rx_k(r)] : 0.0;
        __syncwarp(gmask);
    }

    // rates (and their derivatives) from sh.y into the signed value tables; all lanes of the group
    template <bool DERIV>
    static __device__ __forceinline__ void rates(Shared &sh, const Tables &T, int lane, unsigned gmask) {
#pragma unroll
        for (int m = 0; m < NRM; ++m) {
            const int r = m * GL + lane;
            const unsigned int t = T.rtab[r];
            const double c = sh.c[r], ya = sh.y[t & 0xffffu], yb = sh.y[t >> 16];
            const double cy = c * yb, v = cy * ya;
            *reinterpret_cast<double2 *>(&sh.sv[2 * r]) = make_double2(v, -v);
            if (DERIV) {
                const double db = c * ya;
                *reinterpret_cast<double2 *>(&sh.sd[4 * r]) = make_double2(cy, -cy);
                *reinterpret_cast<double2 *>(&sh.sd[4 * r + 2]) = make_double2(db, -db);
            }
        }
        __syncwarp(gmask);
    }

    // f of this lane's core / leaf species at the state the rates were evaluated at
    static __device__ __forceinline__ void rhs_rows(const Shared &sh, const Tables &T, int lane, double &f, double &f2) {
        double fa[4] = {0.0, 0.0, 0.0, 0.0};               // four short dependent chains instead of one long one
#pragma unroll
        for (int e = 0; e < Net::EF; ++e) fa[e & 3] += sh.sv[T.ftab[e * GL + lane]];
        f = (fa[0] + fa[1]) + (fa[2] + fa[3]);
        f2 = 0.0;
        if (NX > 0 && lane < NX) {
#pragma unroll
            for (int e = 0; e < Net::EF2; ++e) f2 += sh.sv[T.ftab2[e * NXA + lane]];
        }
    }

    static __device__ __forceinline__ void publish(Shared &sh, int lane, unsigned gmask, double yi, double y2) {
        __syncwarp(gmask);
        if (lane < NC) sh.y[lane] = yi;
        if (NX > 0 && lane < NX) sh.y[NC + lane] = y2;
        __syncwarp(gmask);
    }

    // Jacobian rows of the leaf species (derivative table must be current); lanes < NX
    static __device__ __forceinline__ void leaf_rows(Shared &sh, const Tables &T, int lane) {
        if (NX > 0 && lane < NX) {
            double acc = 0.0;
            double *row = sh.J + lane * LD;
#pragma unroll
            for (int e = 0; e < Net::EJ2; ++e) {
                const unsigned int t = T.jtab2[e * NXA + lane];
                acc = sh.sd[t & 0xffffu] + ((t >> 24) ? 0.0 : acc);
                row[(t >> 16) & 0xffu] = acc;
            }
        }
    }

    // One RODAS4 step for the whole group.  (yi, y2) = this lane's core / leaf species; returns false on a
    // bad pivot.  Stage vectors live in shared memory (each lane touches only its own entries), so the
    // six stages run through ONE rolled loop: the generated substitution code exists once per kernel.
    static __device__ bool step(Shared &sh, const Tables &T, int lane, unsigned gmask, int gbase, double yi, double y2,
                                double h, double &ynew, double &ynew2, double &yerr, double &yerr2) {
        const double ih = 1.0 / h, diag = 1.0 / (ROD_GAMMA * h);
        // ---- the lane's row of M = diag*I - J straight into registers -----------------------------
        publish(sh, lane, gmask, yi, y2);
        if (lane == 0) sh.sd[4 * NRP] = sh.sd[4 * NRP + 1] = 0.0;     // (the previous step's substitutions used sd)
        rates<true>(sh, T, lane, gmask);
        {
            double da = diag, db = 0.0;                  // diag - J_ii into the lane's own slot of the table
#pragma unroll
            for (int e = 0; e < Net::ED; ++e) {
                const double t = sh.sd[T.dtab[e * GL + lane]];
                if (e & 1) db += t;
                else da += t;
            }
            sh.sd[DG + lane] = da + db;
        }
        leaf_rows(sh, T, lane);
        double F, F2;
        rhs_rows(sh, T, lane, F, F2);
        __syncwarp(gmask);
        double M[NC], A[NXA];
        Net::load_m(M, A, sh.sd, T.mtab, lane);
        __syncwarp(gmask);                               // sd is reused by the substitutions
        double inv_d = 1.0, inv_d2 = 0.0;
        const bool ok = Net::lu(M, A, sh.J, diag, lane, GL == 32 ? 0xffffffffu : gmask, gbase, inv_d, inv_d2, sh.sd);
        if (!__all_sync(gmask, ok)) return false;
        const unsigned gtm = (1u << lane) - 1u;          // bit c set: lane > c
        const unsigned ltm = ~((2u << lane) - 1u);       // bit c set: lane < c
        Net::scale_u(M, inv_d, ltm);
        // ---- stages: k_s = M^-1 (f(y + sum_j a_sj k_j) + (1/h) sum_j c_sj k_j) ------------------
        const bool core = lane < NC, leaf = NX > 0 && lane < NX;
        double *kc = sh.k + (core ? lane : 0), *kl = sh.k + NC + (leaf ? lane : 0);
        double u = yi, u2 = y2, b = 0.0, b2 = 0.0;
#pragma unroll 1
        for (int stage = 0; stage < 6; ++stage) {
            double r = 0.0, r2 = 0.0;
            u = yi;
            u2 = y2;
#pragma unroll 1
            for (int j = 0; j < stage; ++j) {
                const double a = ROD_A[stage * 5 + j], c = ROD_C[stage * 5 + j];
                const double kj = core ? kc[j * NS] : 0.0, kj2 = leaf ? kl[j * NS] : 0.0;
                u = fma(a, kj, u);
                r = fma(c, kj, r);
                u2 = fma(a, kj2, u2);
                r2 = fma(c, kj2, r2);
            }
            if (stage > 0) {
                publish(sh, lane, gmask, u, u2);
                rates<false>(sh, T, lane, gmask);
                rhs_rows(sh, T, lane, F, F2);
            }
            b = fma(ih, r, F);
            b2 = fma(ih, r2, F2);
            Net::solve(M, A, sh.J, inv_d, inv_d2, b, b2, lane, GL == 32 ? 0xffffffffu : gmask, gbase, gtm, ltm, sh.sd);
            if (core) kc[stage * NS] = b;
            if (leaf) kl[stage * NS] = b2;
        }
        ynew = u + b;                                    // the sixth stage is the error estimate
        yerr = b;
        ynew2 = u2 + b2;
        yerr2 = b2;
        return true;
    }

    struct Control {
        double rtol, atol, h;
        int steps, max_steps;
    };

    // Integrate from t to tn (adaptive steps, the last one clipped to land on tn); group-uniform verdict.
    static __device__ __noinline__ bool advance(Shared &sh, const Tables &T, int lane, unsigned gmask, int gbase, double &yi,
                                   double &y2, double &t, double tn, Control &ctl) {
        double h = ctl.h;
        while (t < tn) {                             // every quantity below is group-uniform
            const double rem = tn - t;
            const bool last = h >= rem;
            const double hh = last ? rem : h;
            if (++ctl.steps > ctl.max_steps) return false;
            double ynew, ynew2, yerr, yerr2;
            if (!step(sh, T, lane, gmask, gbase, yi, y2, hh, ynew, ynew2, yerr, yerr2)) {
                if (hh < 1e-300) return false;
                h = hh * 0.1;
                continue;
            }
            // error norm: weighted RMS over the species, the norm of the reference's solver (VODE behind
            // ScipyOdeSimulator); GL_COOP_MAX_NORM selects the max norm.  The group reduction and the step-size
            // factor run in single precision (an error ESTIMATE steering h; identical on every lane)
            float e = 0.f;
            bool fin = true;
            if (lane < NC) {
                const double q = fabs(yerr) / (ctl.atol + ctl.rtol * fmax(fabs(yi), fabs(ynew)));
                fin = isfinite(ynew) && q == q;
                e = (float)q;
            }
            if (NX > 0 && lane < NX) {
                const double q2 = fabs(yerr2) / (ctl.atol + ctl.rtol * fmax(fabs(y2), fabs(ynew2)));
                fin = fin && isfinite(ynew2) && q2 == q2;
#ifdef GL_COOP_MAX_NORM
                e = fmaxf(e, (float)q2);
#else
                e = sqrtf(e * e + (float)q2 * (float)q2);
#endif
            }
#ifdef GL_COOP_MAX_NORM
#pragma unroll
            for (int o = GL / 2; o > 0; o >>= 1) e = fmaxf(e, __shfl_xor_sync(gmask, e, o));
#else
            e = e * e;
#pragma unroll
            for (int o = GL / 2; o > 0; o >>= 1) e += __shfl_xor_sync(gmask, e, o);
            e = sqrtf(e * (1.0f / NS));
#endif
            fin = __all_sync(gmask, fin);
            if (!fin) {
                if (hh < 1e-300) return false;
                h = hh * 0.1;
                continue;
            }
            // 0.9 e^(-1/4), clamped to [0.2, 5]
            const double fac = e > 0.f ? (double)fminf(5.0f, fmaxf(0.2f, 0.9f * rsqrtf(sqrtf(e)))) : 5.0;
            if (e <= 1.0) {
                yi = ynew;
                y2 = ynew2;
                t = last ? tn : t + hh;
                // a step clipped to land on an output time only lowers the running step size if its
                // own error asks for that; otherwise the unclipped proposal stands
                if (!last || fac < 1.0) h = hh * fac;
                else h = fmax(h, hh * fac);
            } else {
                h = hh * fac;
                if (h < 1e-14 * fabs(tn)) return false;
            }
        }
        ctl.h = h;
        return true;
    }

    // sh.p = sh.pb with the overrides of condition c; optionally the pre-equilibration zeros
    static __device__ __forceinline__ void condition_params(Shared &sh, const double *B, int c, bool zeroed, int lane,
                                                            unsigned gmask) {
        const int np = (int)B[PB_NPARAMS];
        const double *optr = B + (int)B[PB_OFF_OVERPTR], *oidx = B + (int)B[PB_OFF_OVERIDX];
        const double *oval = B + (int)B[PB_OFF_OVERVAL], *zidx = B + (int)B[PB_OFF_ZEROIDX];
        __syncwarp(gmask);
        for (int i = lane; i < np; i += GL) sh.p[i] = sh.pb[i];
        __syncwarp(gmask);
        for (int z = (int)optr[c] + lane; z < (int)optr[c + 1]; z += GL) sh.p[(int)oidx[z]] = oval[z];
        __syncwarp(gmask);
        if (zeroed)
            for (int z = lane; z < (int)B[PB_NZERO]; z += GL) sh.p[(int)zidx[z]] = 0.0;
        rate_constants(sh, lane, gmask);
    }

    // Contribution of condition c to the log-likelihood for the parameters in sh.pb (group-uniform):
    // pre-equilibration, main solve, data rows.  -inf on integrator failure.
    static __device__ double condition_loglike(Shared &sh, const Tables &T, const double *B, int c, int lane,
                                               unsigned gmask, int gbase) {
        const int nt = (int)B[PB_NTIMES], like = (int)B[PB_LIKE];
        const double *tspan = B + (int)B[PB_OFF_TSPAN], *meta = B + (int)B[PB_OFF_ROWMETA];
        const double *data = B + (int)B[PB_OFF_DATA], *sd = B + (int)B[PB_OFF_SD], *mask = B + (int)B[PB_OFF_MASK];
        const double *cterm = B + (int)B[PB_OFF_CTERM], *rowptr = B + (int)B[PB_OFF_ROWPTR];
        const double *stopptr = B + (int)B[PB_OFF_STOPPTR], *stops = B + (int)B[PB_OFF_STOPS];
        const int r0 = (int)rowptr[c], r1 = (int)rowptr[c + 1];
        const int s0 = (int)stopptr[c], s1 = (int)stopptr[c + 1];
        Control ctl;
        ctl.rtol = B[PB_RTOL];
        ctl.atol = B[PB_ATOL];
        ctl.steps = 0;
        ctl.max_steps = (int)B[PB_MAXSTEPS];
        const bool preeq = B[PB_PREEQ] != 0.0;
        condition_params(sh, B, c, preeq, lane, gmask);
        double yi = 0.0, y2 = 0.0;
        if (lane < NC) yi = Net::init_kind(lane) ? sh.p[Net::init_par(lane)] : Net::init_val(lane);
        if (NX > 0 && lane < NX) y2 = Net::init_kind(NC + lane) ? sh.p[Net::init_par(NC + lane)] : Net::init_val(NC + lane);
        bool failed = false;
        if (preeq) {
            const double te = B[PB_PREEQ_TEND];
            double t = 0.0;
            ctl.h = 1e-3 * te;
            failed = !advance(sh, T, lane, gmask, gbase, yi, y2, t, te, ctl);
            if (lane == 0) sh.n_solves += 1;
            if (!failed && B[PB_PREEQ_TOL] >= 0.0) {        // equilibrium test: max |dy/dt| at t_end
                publish(sh, lane, gmask, yi, y2);
                rates<false>(sh, T, lane, gmask);
                double f, f2;
                rhs_rows(sh, T, lane, f, f2);
                double e = fmax(fabs(f), fabs(f2));
                if (e != e) e = INFINITY;
#pragma unroll
                for (int o = GL / 2; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(gmask, e, o));
                failed = !(e <= B[PB_PREEQ_TOL]);
            }
            if (!failed) condition_params(sh, B, c, false, lane, gmask);
        }
        double acc = 0.0;
        auto misfit = [&](int it) {                      // lane 0 accumulates
            publish(sh, lane, gmask, yi, y2);
            if (lane == 0) {
                for (int r = r0; r < r1; ++r) {
                    if (mask[r * nt + it] == 0.0) continue;
                    const int o = (int)meta[4 * r + 1];
                    double sim = 0.0;
                    for (int z = Net::obs_ptr(o); z < Net::obs_ptr(o + 1); ++z) sim = fma(Net::obs_coef(z), sh.y[Net::obs_sp(z)], sim);
                    sim = sim / meta[4 * r + 2];
                    const double mu = data[r * nt + it];
                    if (like == ODE_SNLPDF) {
                        const double zz = (sim - mu) / sd[r * nt + it];
                        if ((int)meta[4 * r + 3] == ODE_ROW_HALFNORM && !(zz >= 0.0)) acc = zz == zz ? -INFINITY : zz;
                        else acc += fma(-0.5 * zz, zz, cterm[r * nt + it]);
                    } else {
                        const double e = mu - sim;
                        acc -= like == ODE_SSE ? e * e : e * e * sd[r * nt + it];
                    }
                }
            }
        };
        if (!failed && s1 > s0) {
            double t = tspan[0];
            const double t_end = tspan[(int)stops[s1 - 1]];
            int s = s0;
            if ((int)stops[s] == 0) { misfit(0); ++s; }
            if (s < s1) {
                ctl.h = fmin(1e-3 * (t_end - t), tspan[(int)stops[s]] - t);
                if (lane == 0) sh.n_solves += 1;
            }
            for (; s < s1 && !failed; ++s) {
                const int it = (int)stops[s];
                failed = !advance(sh, T, lane, gmask, gbase, yi, y2, t, tspan[it], ctl);
                if (!failed) misfit(it);
            }
        }
        if (lane == 0) sh.n_steps += (unsigned long long)ctl.steps;
        acc = __shfl_sync(gmask, acc, gbase);
        return (failed || acc != acc) ? -INFINITY : acc;
    }

    // sum of norm(loc, scale).logpdf over the cycle products of sh.pb (run_jarm_multinest.py:72-103)
    static __device__ double cycle_logp(const Shared &sh, const double *B) {
        const int nc = (int)B[PB_NCYCLES];
        if (nc == 0) return 0.0;
        const double *ptr = B + (int)B[PB_OFF_CYCPTR], *fac = B + (int)B[PB_OFF_CYCFAC];
        const double loc = B[PB_CYC_LOC], scale = B[PB_CYC_SCALE];
        double lp = 0.0;
        for (int i = 0; i < nc; ++i) {
            double box = 1.0;
            for (int z = (int)ptr[i]; z < (int)ptr[i + 1]; ++z) {
                const double t = sh.pb[(int)fac[3 * z]] / sh.pb[(int)fac[3 * z + 1]];
                box = box * (fac[3 * z + 2] != 0.0 ? 1.0 / t : t);
            }
            const double zz = (box - loc) / scale;
            lp += -0.5 * zz * zz - log(scale) - GL_HALF_LOG_2PI;
        }
        return lp != lp ? -INFINITY : lp;
    }

    // pb[mask] = 10**Y with component jrep of Y replaced by v (jrep < 0: none)
    static __device__ __forceinline__ void load_params(Shared &sh, const double *B, const double *Y, int stride, int d,
                                                       int jrep, double v, int lane, unsigned gmask) {
        const int np = (int)B[PB_NPARAMS];
        const double *nominal = B + (int)B[PB_OFF_NOMINAL], *map = B + (int)B[PB_OFF_MAP];
        const double *islog = B + (int)B[PB_OFF_ISLOG];
        __syncwarp(gmask);
        for (int i = lane; i < np; i += GL) sh.pb[i] = nominal[i];
        __syncwarp(gmask);
        for (int j = lane; j < d; j += GL) {
            const double yj = j == jrep ? v : Y[j * stride];
            sh.pb[(int)map[j]] = islog[j] != 0.0 ? exp10(yj) : yj;
        }
        __syncwarp(gmask);
    }

    // conditions sub, sub + nsub, ... of the evaluation whose parameters are in sh.pb; subset 0 adds the cycle terms
    static __device__ double subset_loglike(Shared &sh, const Tables &T, const double *B, int sub, int nsub, int lane,
                                            unsigned gmask, int gbase) {
        const int ncond = (int)B[PB_NCOND];
        double total = 0.0;
        if (sub == 0) {
            double lp = 0.0;
            if (lane == 0) lp = cycle_logp(sh, B);
            total = __shfl_sync(gmask, lp, gbase);
        }
        for (int c = sub; c < ncond; c += nsub) {
            if (total == -INFINITY) break;                                   // nothing can lift it
            total += condition_loglike(sh, T, B, c, lane, gmask, gbase);
        }
        return total != total ? -INFINITY : total;
    }
};

// device-resident counters of one plugin (allocated on first use, never freed)
template <class Net>
inline CoopCounters *coop_counter_store() {
    static CoopCounters *p = nullptr;
    if (!p) {
        if (cudaMalloc(&p, sizeof(CoopCounters)) != cudaSuccess) { p = nullptr; return nullptr; }
        cudaMemset(p, 0, sizeof(CoopCounters));
    }
    return p;
}

// ---- batch evaluation: one group per parameter set ------------------------------------------------
template <class Net>
__global__ void __launch_bounds__(128, GL_COOP_MIN_BLOCKS) ode_coop_eval_kernel(ModelDev M, const double *__restrict__ x, long long n,
                                                            double *__restrict__ out, CoopCounters *ctr,
                                                            unsigned long long *next) {
    using C = CoopOde<Net>;
    constexpr int GL = Net::GL, GPB = 128 / GL;
    extern __shared__ double2 dyn2[];
    typename C::Tables &T = *reinterpret_cast<typename C::Tables *>(dyn2);
    typename C::Shared *shs = reinterpret_cast<typename C::Shared *>(&T + 1);
    const int g = threadIdx.x / GL, lane = threadIdx.x % GL;
    const int gbase = (threadIdx.x & 31) / GL * GL;
    const unsigned gmask = (GL == 32 ? 0xffffffffu : ((1u << GL) - 1u)) << gbase;
    typename C::Shared &sh = shs[g];
    C::load_tables(T);
    C::init_shared(sh, lane, gmask);
    __syncthreads();
    unsigned long long evals = 0ull;
    // solves differ in length by orders of magnitude: groups claim parameter sets one at a time
    for (;;) {
        long long i = 0;
        if (lane == 0) i = (long long)atomicAdd(next, 1ull);
        i = __shfl_sync(gmask, i, gbase);
        if (i >= n) break;
        C::load_params(sh, M.aux0, x + i * M.d, 1, M.d, -1, 0.0, lane, gmask);
        const double L = C::subset_loglike(sh, T, M.aux0, 0, 1, lane, gmask, gbase);
        if (lane == 0) out[i] = L;
        ++evals;
    }
    if (lane == 0 && evals) {
        atomicAdd(&ctr->evals, evals);
        atomicAdd(&ctr->solves, sh.n_solves);
        atomicAdd(&ctr->steps, sh.n_steps);
    }
}


// ---- walkers with a cooperative likelihood ---------------------------------------------------------
// Same chain as walk_kernel (walker.cuh) -- same Philox counters, proposal, prior test, hard wall,
// phases and pooled width adaptation -- but the lanes of a group integrate ONE system together
// instead of speculating on different trial moves, and a walker owns `gpw` groups, one per subset of
// the protocol's conditions (independent solves), whose partial sums meet in shared memory.
template <class Net>
__global__ void __launch_bounds__(128, GL_COOP_MIN_BLOCKS) walk_coop_kernel(WalkParams P, ModelDev M, int phase, int gpw,
                                                        CoopCounters *ctr) {
    using C = CoopOde<Net>;
    constexpr int GL = Net::GL, GPB = 128 / GL;
    extern __shared__ double2 dyn2[];
    EngineState *S = P.S;
    if (!S->do_walk) return;
    const int k = S->k_cur, d = P.d;
    const int WPB = GPB / gpw;                     // walkers per block
    if (blockIdx.x * WPB >= k) return;
    typename C::Tables &T = *reinterpret_cast<typename C::Tables *>(dyn2);
    typename C::Shared *shs = reinterpret_cast<typename C::Shared *>(&T + 1);
    double *xs = reinterpret_cast<double *>(shs + GPB);   // [d][WPB]
    double *wj = xs + (size_t)d * WPB;
    double *p0s = wj + d, *p1s = p0s + d;
    double *part = p1s + d;                        // [2][GPB] partial sums, double-buffered by move parity
    int *kinds = reinterpret_cast<int *>(part + 2 * GPB);
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
    C::load_tables(T);
    __syncthreads();
    const int g = tid / GL, lane = tid % GL;
    const int wl = g / gpw, sub = g % gpw;         // walker of this block, condition subset of this group
    const int wlane = tid - wl * gpw * GL;         // index inside the walker's threads
    const int wthreads = gpw * GL;
    const int w = blockIdx.x * WPB + wl;
    const bool active = wl < WPB && w < k;
    const int gbase = (tid & 31) / GL * GL;
    const unsigned gmask = (GL == 32 ? 0xffffffffu : ((1u << GL) - 1u)) << gbase;
    // barrier over the threads of one walker: a warp-level sync when they share a warp
    const unsigned wmask = wthreads >= 32 ? 0xffffffffu : ((1u << wthreads) - 1u) << ((tid & 31) / wthreads * wthreads);
    auto walker_sync = [&]() {
        if (wthreads <= 32) __syncwarp(wmask);
        else asm volatile("barrier.sync %0, %1;" ::"r"(1 + wl), "r"(wthreads) : "memory");   // non-aligned form: the
        // two groups of a 16-lane pair sharing a warp may arrive from different iterations of their step loops
    };
    typename C::Shared &sh = shs[g];
    C::init_shared(sh, lane, gmask);
    unsigned long long evals = 0ull;
    const bool tracing = ctr->trace_on != 0ull;
    const unsigned long long t_begin = tracing ? coop_globaltimer() : 0ull;
    double *x = xs + wl;
    const double lstar = S->lstar;
    const uint32_t step = (uint32_t)S->step;
    if (active) {
        const long long slot = P.dead_list[w];
        const uint32_t gid = (uint32_t)(slot + P.row_offset);
        double L;
        if (phase == 0 || !tuning) {
            long long pick = 0;
            if (wlane == 0) {
                for (uint32_t a = 0;; ++a) {
                    const U2 u = philox_uniforms(P.seed, gid, step, a, GL_TAG_PICK);
                    pick = (long long)(u.a * (double)P.N);
                    if (pick >= P.N) pick = P.N - 1;
                    if (!P.is_dead[pick]) break;
                }
                part[g] = __longlong_as_double(pick);
            }
            walker_sync();
            pick = __double_as_longlong(part[wl * gpw]);
            walker_sync();
            for (int j = wlane; j < d; j += wthreads) x[j * WPB] = P.live[pick * d + j];
            L = P.logl[pick];
        } else {
            for (int j = wlane; j < d; j += wthreads) x[j * WPB] = P.scratch_x[(size_t)w * d + j];
            L = P.scratch_l[w];
        }
        walker_sync();
        const int n_moves = (tuning ? 20 : P.iterations + P.burn_in) * d;
        const uint32_t c2_base = (uint32_t)((tuning ? phase * 20 : P.tuning_cycles * 20) * d);
        const bool normal = P.proposal == GL_PROPOSAL_NORMAL;
        int parity = 0;
        for (int m = 0; m < n_moves; ++m) {
            const int j = m % d;
            const U2 ru = philox_uniforms(P.seed, gid, step, c2_base + (uint32_t)m, GL_TAG_MOVE);
            const double width = wj[j], xj = x[j * WPB];
            const double stepv = __dmul_rn(normal ? normcdfinv(ru.a) : ru.a, width);
            const double nj = __dadd_rn(stepv, normal ? xj : __dadd_rn(xj, -0.5 * width));
            if (prior_accept(kinds[j], p0s[j], p1s[j], xj, nj, ru.b)) {        // walker-uniform
                C::load_params(sh, M.aux0, x, WPB, d, j, nj, lane, gmask);
                double Lt = C::subset_loglike(sh, T, M.aux0, sub, gpw, lane, gmask, gbase);
                ++evals;
                if (gpw > 1) {
                    double *pp = part + parity * GPB + wl * gpw;
                    if (lane == 0) pp[sub] = Lt;
                    walker_sync();
                    Lt = pp[0];
                    for (int s = 1; s < gpw; ++s) Lt += pp[s];
                    if (Lt != Lt) Lt = -INFINITY;
                    parity ^= 1;
                }
                if (Lt > lstar) {
                    walker_sync();
                    if (wlane == 0) {
                        x[j * WPB] = nj;
                        if (tuning) atomicAdd(&acc[j], 1u);
                    }
                    L = Lt;
                    walker_sync();
                }
            }
        }
        if (tuning) {
            if (phase + 1 < P.tuning_cycles) {
                for (int j = wlane; j < d; j += wthreads) P.scratch_x[(size_t)w * d + j] = x[j * WPB];
                if (wlane == 0) P.scratch_l[w] = L;
            }
        } else {
            for (int j = wlane; j < d; j += wthreads) P.live[slot * d + j] = x[j * WPB];
            if (wlane == 0) P.logl[slot] = L;
        }
        if (lane == 0) {
            if (sub == 0) atomicAdd(&ctr->evals, evals);
            atomicAdd(&ctr->solves, sh.n_solves);
            atomicAdd(&ctr->steps, sh.n_steps);
        }
    }
    if (tuning) {
        __syncthreads();
        for (int j = tid; j < d; j += 128)
            if (acc[j]) atomicAdd(&S->acc[phase][j], acc[j]);
    }
    if (tracing) {
        __syncthreads();
        if (tid == 0) {
            unsigned long long steps = 0ull;
            for (int q = 0; q < GPB; ++q) steps += shs[q].n_steps;
            const unsigned long long slot = atomicAdd(&ctr->trace_n, 1ull);
            if (slot < COOP_TRACE_CAP) {
                unsigned long long *r = ctr->trace + 4 * slot;
                r[0] = ((unsigned long long)phase << 32) | coop_smid();
                r[1] = t_begin;
                r[2] = coop_globaltimer();
                r[3] = steps;
            }
        }
    }
}

// groups a walker spreads its conditions over: as many as there are conditions while a block still
// holds whole walkers (1, 2 or 4 groups of a 128-thread block)
template <class Net>
inline int coop_groups_per_walker(const ModelDev &M, int ncond) {
    constexpr int GPB = 128 / Net::GL;
    int gpw = 1;
    while (gpw * 2 <= GPB && gpw * 2 <= ncond && ncond % (gpw * 2) == 0) gpw *= 2;
    return gpw;
}

template <class Net>
inline int walk_coop_launch(const WalkParams &P, const ModelDev &M, int ncond, cudaStream_t st) {
    using C = CoopOde<Net>;
    constexpr int GPB = 128 / Net::GL;
    const int gpw = coop_groups_per_walker<Net>(M, ncond);
    const int WPB = GPB / gpw;
    const size_t smem = sizeof(typename C::Tables) + sizeof(typename C::Shared) * GPB +
                        sizeof(double) * ((size_t)P.d * WPB + 3 * (size_t)P.d + 2 * GPB) + 2 * sizeof(int) * P.d + 16;
    GL_CUDA(cudaFuncSetAttribute(walk_coop_kernel<Net>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CoopCounters *ctr = coop_counter_store<Net>();
    GL_REQUIRE(ctr != nullptr, "counter allocation failed");
    const int blocks = (P.k_max + WPB - 1) / WPB;
    for (int phase = 0; phase <= P.tuning_cycles; ++phase)
        walk_coop_kernel<Net><<<blocks, 128, smem, st>>>(P, M, phase, gpw, ctr);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

template <class Net>
inline int ode_coop_eval_launch(const ModelDev &M, const double *x, long long n, double *out, cudaStream_t st) {
    using C = CoopOde<Net>;
    constexpr int GPB = 128 / Net::GL;
    const size_t smem = sizeof(typename C::Tables) + sizeof(typename C::Shared) * GPB;
    GL_CUDA(cudaFuncSetAttribute(ode_coop_eval_kernel<Net>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CoopCounters *ctr = coop_counter_store<Net>();
    GL_REQUIRE(ctr != nullptr, "counter allocation failed");
    long long blocks = (n + GPB - 1) / GPB;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    GL_CUDA(cudaMemsetAsync(&ctr->next, 0, sizeof(unsigned long long), st));
    ode_coop_eval_kernel<Net><<<(int)blocks, 128, smem, st>>>(M, x, n, out, ctr, &ctr->next);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

// switch the walker trace on/off (clears it), or read it: returns the number of records copied to out (4 words each)
template <class Net>
inline long long coop_trace(int enable, unsigned long long *out, long long max_records) {
    CoopCounters *ctr = coop_counter_store<Net>();
    if (!ctr) return -1;
    if (out == nullptr) {
        unsigned long long v[2] = {enable ? 1ull : 0ull, 0ull};
        if (cudaMemcpy(&ctr->trace_on, v, sizeof(v), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        return 0;
    }
    unsigned long long n = 0;
    if (cudaMemcpy(&n, &ctr->trace_n, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (n > COOP_TRACE_CAP) n = COOP_TRACE_CAP;
    if ((long long)n > max_records) n = (unsigned long long)max_records;
    if (n && cudaMemcpy(out, ctr->trace, 4 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)n;
}

template <class Net>
inline int coop_counters(unsigned long long *out3, int reset) {
    CoopCounters *ctr = coop_counter_store<Net>();
    GL_REQUIRE(ctr != nullptr, "counter allocation failed");
    unsigned long long h3[3];
    GL_CUDA(cudaMemcpy(h3, ctr, sizeof(h3), cudaMemcpyDeviceToHost));
    out3[0] = h3[0]; out3[1] = h3[1]; out3[2] = h3[2];
    if (reset) GL_CUDA(cudaMemset(ctr, 0, 3 * sizeof(unsigned long long)));
    return GL_OK;
}

}  // namespace gl
