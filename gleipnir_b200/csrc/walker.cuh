// Constrained component-wise Metropolis walkers: the B200 form of
// MetropolisComponentWiseHardNSRejection.__call__ (samplers.py:56-152) applied to a whole
// batch of dead slots at once.  Each walker replaces one dead slot:
//   * seed  = uniformly chosen survivor, int(U*N) redrawn while it hits a dead slot
//             (nested_sampling.py:147-149);
//   * move  = the reference's *tuning-loop* move (samplers.py:95-114): propose component j
//             from a symmetric uniform/normal window of width w_j, accept iff
//             u < prior_j(new)/prior_j(cur) AND logL(new) > L*  (hard NS wall);
//   * phases: `tuning_cycles` launches of 20 sweeps whose positions are discarded and whose
//             pooled acceptance adapts w (see phase_width), then one launch of
//             iterations+burn_in sweeps from the seed again whose end point is written
//             into the dead slot (samplers.py:124-152 without quirk Q1).
//
// Speculative prefetch.  A Metropolis chain is one dependent sequence of trial moves, so a
// walker-per-thread kernel is bound by the latency of that chain (r01 baseline: 423 warp
// instructions and ~2500 cycles per trial move with one warp per SM sub-partition).  Here a
// GROUP of G lanes (G = 1..32, chosen on the host from the batch size) serves one walker:
// lane i evaluates trial move m+i from the CURRENT state, which is exactly the state move
// m+i would see if moves m..m+i-1 are all rejected.  A ballot finds the first accepting
// lane f; moves m..m+f-1 were rejections, move m+f is committed, and the window restarts
// at m+f+1.  Randomness is Philox4x32-10 keyed by (global slot, step, move index) -- one
// block yields the proposal uniform r and the acceptance uniform u of a trial move -- so the
// retired sequence of moves, decisions and values is IDENTICAL to the sequential chain
// for every G; speculation only shortens the critical path (~1/p_accept moves per round).
//
// State: the walker's point x[d], the likelihood cache (see likelihoods.cuh) and logL live
// in shared memory, column-per-walker (conflict-free; lanes of a group broadcast-read).
// Bound: FP64 pipe / dependent-chain latency; HBM traffic is 8(2d+1) B per replacement.
#pragma once
#include "engine.cuh"
#include "philox.cuh"

namespace gl {

constexpr int WALK_THREADS = 128;

// moves pre-generated per refill of a walker's random-number ring
__host__ __device__ inline int walker_chunk(int G) { return 4 * G > 8 ? 4 * G : 8; }

// FULL = true: G == 32, one group per warp, every intra-group primitive uses the constant full
// mask (no mask-validation code); FULL = false: runtime power-of-two G < 32.
template <class Like, bool FULL>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(WalkParams P, ModelDev M, int phase, int lgG) {
    extern __shared__ double sm[];
    EngineState *S = P.S;
    pdl_wait();                                    // launched with launch_pdl (common.cuh): only the launch
    pdl_trigger();                                 // latency of the next kernel overlaps this one
    if (!S->do_walk) return;                       // grid-uniform
    const int k = S->k_cur;
    const int d = P.d;
    const int G = FULL ? 32 : (1 << lgG);
    const int WPB = WALK_THREADS >> lgG;           // walkers per block
    if (blockIdx.x * WPB >= k) return;             // block-uniform
    const int nc = Like::cache_size(d);
    const int CH = walker_chunk(G);
    // shared layout: x[d][WPB] | cache[nc][WPB] | step[WPB][CH] | u[WPB][CH] | width[d] | p0[d] | p1[d]
    //                | kind[d] | acc[d]
    double *xs = sm;
    double *cs = xs + (size_t)d * WPB;
    double *sbuf = cs + (size_t)nc * WPB;
    double *ubuf = sbuf + (size_t)WPB * CH;
    double *wj = ubuf + (size_t)WPB * CH;
    double *p0s = wj + d;
    double *p1s = p0s + d;
    int *kinds = reinterpret_cast<int *>(p1s + d);
    unsigned int *acc = reinterpret_cast<unsigned int *>(kinds + d);

    const int tid = threadIdx.x;
    const bool tuning = phase < P.tuning_cycles;
    for (int j = tid; j < d; j += WALK_THREADS) {
        wj[j] = phase_width(S, j, phase, k);
        p0s[j] = P.priors->p0[j];
        p1s[j] = P.priors->p1[j];
        kinds[j] = P.priors->kind[j];
        acc[j] = 0u;
    }
    __syncthreads();

    const int wl = tid >> lgG;                     // walker within block
    const int lane = tid & (G - 1);                // lane within group
    const int w = blockIdx.x * WPB + wl;           // walker within batch
    const bool active = w < k;                     // group-uniform
    const unsigned gbase = FULL ? 0u : (unsigned)((tid & 31) & ~(G - 1));
    const unsigned gmask = FULL ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    double *x = xs + wl;
    double *c = cs + wl;
    double *sb = sbuf + (size_t)wl * CH;
    double *ub = ubuf + (size_t)wl * CH;
    const Like like(M);
    const double lstar = S->lstar;
    const uint32_t step = (uint32_t)S->step;
    // n % d == n - umulhi(n, dmagic)*d for n < 2^32/d (dmagic = ceil(2^32/d); d = 1 -> 0 -> n - 0*1 handled below)
    const uint32_t dmagic = d > 1 ? 0xFFFFFFFFu / (uint32_t)d + 1u : 0u;

    if (active) {
        const long long slot = P.dead_list[w];
        const uint32_t g = (uint32_t)(slot + P.row_offset);
        // ---- start point ------------------------------------------------------------------
        if (phase == 0 || !tuning) {
            long long pick = 0;
            if (lane == 0) {
                for (uint32_t a = 0;; ++a) {       // survivor pick, redraw while dead
                    const U2 u = philox_uniforms(P.seed, g, step, a, GL_TAG_PICK);
                    pick = (long long)(u.a * (double)P.N);
                    if (pick >= P.N) pick = P.N - 1;
                    if (!P.is_dead[pick]) break;
                }
            }
            pick = __shfl_sync(gmask, pick, gbase);
            const double *src = P.live + pick * d;
            for (int j = lane; j < d; j += G) x[j * WPB] = src[j];
        } else {
            const double *src = P.scratch_x + (size_t)w * d;
            for (int j = lane; j < d; j += G) x[j * WPB] = src[j];
        }
        __syncwarp(gmask);
        double L = like.init(x, c, WPB, lane, G, gmask);

        // ---- speculative rounds --------------------------------------------------------------
        const int n_moves = (tuning ? 20 : P.iterations + P.burn_in) * d;
        const uint32_t c2_base = (uint32_t)((tuning ? phase * 20 : P.tuning_cycles * 20) * d);
        const bool normal = P.proposal == GL_PROPOSAL_NORMAL;
        int m = 0, base = -CH;                     // ring holds moves [base, base+CH)
        // functors that track their value incrementally (dense Gaussian: O(1) update of the quadratic form)
        // recompute it exactly every RESYNC_SWEEPS sweeps, so rounding drift cannot accumulate over a chain
        constexpr int RESYNC = resync_sweeps<Like>::value;
        int next_resync = RESYNC > 0 ? RESYNC * d : 0x7fffffff;
        while (m < n_moves) {                      // group-uniform
            if (RESYNC > 0 && m >= next_resync) {
                L = like.refresh(x, c, WPB, lane, G, gmask, L);
                next_resync += RESYNC * d;
            }
            if (m + G > base + CH) {
                // refill: everything about moves [m, m+CH) that does not depend on the walker's
                // state -- Philox block, proposal step r*w (or ndtri(r)*w), acceptance uniform.
                // Independent per move: the CH/G Philox chains of a lane overlap (ILP).
                base = m;
                __syncwarp(gmask);
                for (int q = lane; q < CH; q += G) {
                    const int mv = base + q;
                    if (mv < n_moves) {
                        const uint32_t jj = d > 1 ? (uint32_t)mv - __umulhi((uint32_t)mv, dmagic) * (uint32_t)d : 0u;
                        const U2 ru = philox_uniforms(P.seed, g, step, c2_base + (uint32_t)mv, GL_TAG_MOVE);
                        sb[q] = __dmul_rn(normal ? normcdfinv(ru.a) : ru.a, wj[jj]);
                        ub[q] = ru.b;
                    }
                }
                __syncwarp(gmask);
            }
            const int mine = m + lane;
            bool ok = false;
            int j = 0;
            double nj = 0.0, Lt = 0.0, cnew = 0.0;
            if (mine < n_moves) {
                j = d > 1 ? (int)((uint32_t)mine - __umulhi((uint32_t)mine, dmagic) * (uint32_t)d) : 0;
                const double xj = x[j * WPB];
                // uniform: r*w + (x - w/2)   normal: ndtri(r)*w + x     (samplers.py:100,102)
                const double loc = normal ? xj : __dadd_rn(xj, -0.5 * wj[j]);
                nj = __dadd_rn(sb[mine - base], loc);
                if (prior_accept(kinds[j], p0s[j], p1s[j], xj, nj, ub[mine - base])) {
                    Lt = like.trial(x, c, WPB, j, nj, cnew);
                    ok = Lt > lstar;
                }
            }
            const unsigned bal = __ballot_sync(gmask, ok) & gmask;
            if (bal == 0u) {
                m += G;
                continue;
            }
            const int f = __ffs(bal) - 1;          // absolute lane of the first accept
            // broadcast the accepted move to the group and commit cooperatively
            j = __shfl_sync(gmask, j, f);
            nj = __shfl_sync(gmask, nj, f);
            Lt = __shfl_sync(gmask, Lt, f);
            cnew = __shfl_sync(gmask, cnew, f);
            like.commit(x, c, WPB, j, nj, cnew, lane, G, gmask);
            L = Lt;
            if (tuning && lane == 0) atomicAdd(&acc[j], 1u);
            __syncwarp(gmask);
            m += (f - (int)gbase) + 1;
        }

        // ---- write back ----------------------------------------------------------------------
        if (tuning) {
            if (phase + 1 < P.tuning_cycles) {
                double *dst = P.scratch_x + (size_t)w * d;
                for (int j = lane; j < d; j += G) dst[j] = x[j * WPB];
            }
        } else {
            L = like.refresh(x, c, WPB, lane, G, gmask, L);
            double *dst = P.live + slot * d;
            for (int j = lane; j < d; j += G) dst[j] = x[j * WPB];
            if (lane == 0) P.logl[slot] = L;
        }
    }
    if (tuning) {
        __syncthreads();
        for (int j = tid; j < d; j += WALK_THREADS)
            if (acc[j]) atomicAdd(&S->acc[phase][j], acc[j]);
    }
}

// ---- register-resident walker for small d ------------------------------------------------------
// The r01_g profile of walk_kernel at d = 2 shows a round of ~170 dependent, in-order-issued
// instructions at ~8.6 cycles each, more than half of them shared-memory address arithmetic,
// LDS/STS round trips of the walker state and runtime-d loop control.  For d <= REG_MAX_D (8) the
// point, the likelihood cache, widths and prior rows live in REGISTERS, replicated in every lane
// of the group (commit = a shuffle from the accepting lane; no store, no group barrier), and d is
// a template parameter.  Same moves, same arithmetic, same Philox counters: results are
// bit-identical to walk_kernel (tests compare them).
template <int D>
__device__ __forceinline__ double reg_sel(const double (&a)[D], int j) {
    double r = a[0];
#pragma unroll
    for (int i = 1; i < D; ++i) r = j == i ? a[i] : r;
    return r;
}

// ---- deferred bookkeeping of a batch (see kill_cluster_kernel / shard_commit_cluster_kernel) ------
// BOOK_BLOCKS extra 128-thread blocks of the first walker launch do the evidence / information
// update of the batch the walkers are refilling: together they are the 1024 threads of
// accumulate_kernel (thread vt sums elements vt, vt+1024, ...; warp tree; the 32 warp partials meet
// in global scratch and the last block to finish reduces them with the same 32-lane tree), so the
// run stays bit-identical to every other path.  The maximum is order-independent: every block
// takes it over the whole batch by itself.
constexpr int BOOK_BLOCKS = 1024 / WALK_THREADS;

__device__ inline void book_accumulate_block(EngineState *S, const WalkParams &P, int bb) {
    __shared__ double scr[5];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = P.book;
    const double logw = S->book_logw, w = S->book_w, m_old = S->M;
    double mx = -INFINITY;
    for (int i = tid; i < k; i += WALK_THREADS) {
        const double lw = logw + P.book_vals[P.book_idx[i]];
        if (isfinite(lw)) mx = fmax(mx, lw);
    }
    mx = warp_max(mx);
    if (lane == 0) scr[wid] = mx;
    __syncthreads();
    mx = fmax(fmax(scr[0], scr[1]), fmax(scr[2], scr[3]));
    const double m_new = fmax(m_old, mx);
    const int vt = bb * WALK_THREADS + tid;        // virtual thread of the 1024-thread tree
    double sz = 0.0, sh = 0.0, zl = 0.0, hl = 0.0;
    for (int i = vt; i < k; i += 1024) {
        const double L = P.book_vals[P.book_idx[i]];
        const double lw = logw + L;
        if (isfinite(lw)) {
            const double t = exp(lw - m_new);
            sz += t;
            sh += t * L;
        }
        const double dZ = w * exp(L);
        zl += dZ;
        const double dH = dZ * L;
        if (dH == dH) hl += dH;
    }
    sz = warp_sum(sz);
    sh = warp_sum(sh);
    zl = warp_sum(zl);
    hl = warp_sum(hl);
    const int vw = vt >> 5;                        // virtual warp
    if (lane == 0) {
        S->book_part[0][vw] = sz;
        S->book_part[1][vw] = sh;
        S->book_part[2][vw] = zl;
        S->book_part[3][vw] = hl;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&S->book_count, 1u) == BOOK_BLOCKS - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (wid == 0) {
        volatile double *part = &S->book_part[0][0];
        const double szs = warp_sum(part[lane]), shs = warp_sum(part[32 + lane]);
        const double zls = warp_sum(part[64 + lane]), hls = warp_sum(part[96 + lane]);
        if (lane == 0) {
            if (mx > -INFINITY) {
                const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
                S->Zs = S->Zs * scale + szs;
                S->Hs = S->Hs * scale + shs;
                S->M = m_new;
            }
            S->Z_lin += zls;
            S->H_lin += hls;
            S->book_count = 0u;
        }
    }
}

template <class Like, int D, bool FULL>
__global__ void __launch_bounds__(WALK_THREADS) walk_reg_kernel(WalkParams P, ModelDev M, int phase, int lgG) {
    extern __shared__ double sm[];                 // r[WPB][CH] | u[WPB][CH]: raw Philox uniforms of the ring
    __shared__ unsigned int acc_s[D];
    EngineState *S = P.S;
    // PDL (common.cuh): phase 0 follows the kill/commit kernel whose outputs everything below
    // needs, so it waits first.  Later phases follow the previous walker launch of the SAME
    // refill: all they need from it are the pooled acceptance counts (-> widths) and the carried
    // positions, so their whole state-independent prologue -- batch header, survivor pick, seed
    // point, likelihood cache, the first ring of Philox uniforms -- runs BEFORE the wait and
    // overlaps the tail of that launch.
    const bool after_walk = phase > 0;
    if (!after_walk) { pdl_wait(); pdl_trigger(); }
    const bool book = phase == 0 && P.book != 0;   // this launch also does the batch's deferred bookkeeping
    if (book && blockIdx.x + BOOK_BLOCKS >= gridDim.x) {       // (extra blocks; the host only asks when do_walk is certain)
        book_accumulate_block(S, P, (int)(blockIdx.x + BOOK_BLOCKS - gridDim.x));
        return;
    }
    const bool do_walk = S->do_walk != 0;          // grid-uniform
    const int k = S->k_cur;
    const int G = FULL ? 32 : (1 << lgG);
    const int WPB = WALK_THREADS >> lgG;
    if (!do_walk || blockIdx.x * WPB >= k) {       // block-uniform
        if (after_walk) { pdl_wait(); pdl_trigger(); }
        return;
    }
    const int CH = walker_chunk(G);
    double *sbuf = sm;
    double *ubuf = sm + (size_t)WPB * CH;
    const int tid = threadIdx.x;
    const bool tuning = phase < P.tuning_cycles;
    if (tid < D) acc_s[tid] = 0u;
    double p0v[D], p1v[D];
    int kindv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        p0v[j] = P.priors->p0[j];
        p1v[j] = P.priors->p1[j];
        kindv[j] = P.priors->kind[j];
    }

    const int wl = tid >> lgG;
    const int lane = tid & (G - 1);
    const int w = blockIdx.x * WPB + wl;
    const bool active = w < k;                     // group-uniform
    const unsigned gbase = FULL ? 0u : (unsigned)((tid & 31) & ~(G - 1));
    const unsigned gmask = FULL ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    double *sb = sbuf + (size_t)wl * CH;
    double *ub = ubuf + (size_t)wl * CH;
    const Like like(M);
    const double lstar = S->lstar;
    const uint32_t step = (uint32_t)S->step;
    const int n_moves = (tuning ? 20 : P.iterations + P.burn_in) * D;
    const uint32_t c2_base = (uint32_t)((tuning ? phase * 20 : P.tuning_cycles * 20) * D);
    const bool normal = P.proposal == GL_PROPOSAL_NORMAL;
    const bool from_seed = phase == 0 || !tuning;  // else: continue from the previous tuning cycle
    long long slot = 0;
    uint32_t g = 0u;
    double x[D], c[D];
    double L = 0.0;
    int base = -CH;                                // ring holds moves [base, base+CH)
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = c[j] = 0.0;

    // state-independent per move: Philox block -> proposal variate r (or ndtri(r)) and acceptance
    // uniform u.  Independent per move: the CH/G Philox chains of a lane overlap (ILP).
    auto fill_ring = [&](int m0) {
        base = m0;
        __syncwarp(gmask);
        for (int q = lane; q < CH; q += G) {
            const int mv = base + q;
            if (mv < n_moves) {
                const U2 ru = philox_uniforms(P.seed, g, step, c2_base + (uint32_t)mv, GL_TAG_MOVE);
                sb[q] = normal ? normcdfinv(ru.a) : ru.a;
                ub[q] = ru.b;
            }
        }
        __syncwarp(gmask);
    };

    if (active) {
        slot = P.dead_list[w];
        g = (uint32_t)(slot + P.row_offset);
        if (book && S->book_room) {                // record_kernel's job: each walker logs the row it replaces
            const long long row = S->book_base_row + w;
            for (int j = lane; j < D; j += G) P.dead_x[row * D + j] = P.live[slot * D + j];
            if (lane == 0) {
                P.dead_logl[row] = P.logl[slot];
                P.dead_w[row] = S->book_w;
                P.dead_idx[row] = slot + P.row_offset;
                P.dead_step[row] = S->book_step;
            }
        }
        if (from_seed) {
            // survivor pick, redrawn while it hits a dead slot.  The first draw's dead flag and seed row
            // are requested BEFORE the ring is filled and consumed after it, so the two dependent L2
            // round trips hide behind the ~600 instructions of Philox; only a redraw (probability k/N)
            // pays them again.
            long long pick = 0;
            {
                const U2 u = philox_uniforms(P.seed, g, step, 0u, GL_TAG_PICK);    // same value in every lane
                pick = (long long)(u.a * (double)P.N);
                if (pick >= P.N) pick = P.N - 1;
            }
            const unsigned char dead0 = P.is_dead[pick];
            {
                const double *src = P.live + pick * D;
#pragma unroll
                for (int j = 0; j < D; ++j) x[j] = src[j];
            }
            fill_ring(0);
            if (dead0) {                           // group-uniform
                if (lane == 0) {
                    for (uint32_t a = 1;; ++a) {
                        const U2 u = philox_uniforms(P.seed, g, step, a, GL_TAG_PICK);
                        pick = (long long)(u.a * (double)P.N);
                        if (pick >= P.N) pick = P.N - 1;
                        if (!P.is_dead[pick]) break;
                    }
                }
                pick = __shfl_sync(gmask, pick, gbase);
                const double *src = P.live + pick * D;
#pragma unroll
                for (int j = 0; j < D; ++j) x[j] = src[j];
            }
            L = like.template init_reg<D>(x, c);
        } else {
            fill_ring(0);
        }
    }
    if (after_walk) { pdl_wait(); pdl_trigger(); }
    __syncthreads();                               // acc_s cleared

    double wv[D], hw[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        wv[j] = phase_width(S, j, phase, k);
        hw[j] = normal ? -0.0 : -0.5 * wv[j];      // x + (-0) == x exactly, for every x
    }

    if (active) {
        if (!from_seed) {
            const double *src = P.scratch_x + (size_t)w * D;
#pragma unroll
            for (int j = 0; j < D; ++j) x[j] = src[j];
            L = like.template init_reg<D>(x, c);
        }
        unsigned int accr[D];
#pragma unroll
        for (int j = 0; j < D; ++j) accr[j] = 0u;

        // shared-window addresses of the walker's two rings, computed once (the generic->shared
        // conversion otherwise sits at the head of every round's dependent chain)
        const uint32_t sb32 = (uint32_t)__cvta_generic_to_shared(sb), ub32 = (uint32_t)__cvta_generic_to_shared(ub);
        bool all_uniform = true;
#pragma unroll
        for (int i = 0; i < D; ++i) all_uniform = all_uniform && kindv[i] == GL_PRIOR_UNIFORM;
        int m = 0;
        while (m < n_moves) {                      // group-uniform
            if (m + G > base + CH) fill_ring(m);
            const int mine = m + lane;
            bool ok = false;
            double nj = 0.0, Lt = 0.0, cnew = 0.0;
            if (mine < n_moves) {
                const int j = (int)((uint32_t)mine % (uint32_t)D);
                const uint32_t off = (uint32_t)(mine - base) * 8u;
                double r;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(sb32 + off) : "memory");
                const double xj = reg_sel<D>(x, j);
                // uniform: r*w + (x - w/2)   normal: ndtri(r)*w + x     (samplers.py:100,102); hw = -0 for normal
                const double loc = __dadd_rn(xj, reg_sel<D>(hw, j));
                nj = __dadd_rn(__dmul_rn(r, reg_sel<D>(wv, j)), loc);
                bool pa;
                if (all_uniform) {                 // block-uniform: the common case needs no acceptance uniform
                    const double t = nj - reg_sel<D>(p0v, j);
                    pa = t >= 0.0 && t <= reg_sel<D>(p1v, j);
                } else {
                    double u;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(u) : "r"(ub32 + off) : "memory");
                    int kind = kindv[0];
#pragma unroll
                    for (int i = 1; i < D; ++i) kind = j == i ? kindv[i] : kind;
                    pa = prior_accept(kind, reg_sel<D>(p0v, j), reg_sel<D>(p1v, j), xj, nj, u);
                }
                // the trial is evaluated by every lane (no divergent region to reconverge from): a
                // proposal outside the prior is a finite number, its logL is simply not used
                Lt = like.template trial_reg<D>(c, j, nj, cnew);
                ok = pa && Lt > lstar;
            }
            const unsigned bal = __ballot_sync(gmask, ok) & gmask;
            if (bal == 0u) {
                m += G;
                continue;
            }
            const int f = __ffs(bal) - 1;          // absolute lane of the first accept
            const int fm = m + (f - (int)gbase);   // the accepted move; its component is group-uniform
            const int jf = (int)((uint32_t)fm % (uint32_t)D);
            nj = __shfl_sync(gmask, nj, f);
            Lt = __shfl_sync(gmask, Lt, f);
            cnew = __shfl_sync(gmask, cnew, f);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                if (jf == i) { x[i] = nj; c[i] = cnew; accr[i] += 1u; }
            }
            L = Lt;
            m = fm + 1;
        }

        // ---- write back ----------------------------------------------------------------------
        if (tuning) {
            if (lane == 0) {
                if (phase + 1 < P.tuning_cycles) {
                    double *dst = P.scratch_x + (size_t)w * D;
#pragma unroll
                    for (int j = 0; j < D; ++j) dst[j] = x[j];
                }
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (accr[j]) atomicAdd(&acc_s[j], accr[j]);
            }
        } else if (lane == 0) {
            double *dst = P.live + slot * D;
#pragma unroll
            for (int j = 0; j < D; ++j) dst[j] = x[j];
            P.logl[slot] = L;
        }
    }
    if (tuning) {
        __syncthreads();
        if (tid < D && acc_s[tid]) atomicAdd(&S->acc[phase][tid], acc_s[tid]);
    }
}

inline int walker_smem_bytes(int d, int nc, int G) {
    const int WPB = WALK_THREADS / G;
    return (int)(sizeof(double) * ((size_t)(d + nc) * WPB + 2 * (size_t)WPB * walker_chunk(G) + 3 * (size_t)d) +
                 2 * sizeof(int) * d + 16);
}

// lanes per walker: the largest power of two that keeps the whole batch resident
// (k_max * G threads <= one wave at 1024 threads per SM), capped at 32.
inline int walker_group_size(int k_max, int requested) {
    if (requested >= 1) {
        int g = 1;
        while (g * 2 <= requested && g < 32) g *= 2;
        return g;
    }
    const long long cap = (long long)sm_count() * 1024;
    int g = 32;
    while (g > 1 && (long long)k_max * g > cap) g >>= 1;
    return g;
}

struct WalkLauncher {
    const WalkParams &P;
    const ModelDev &M;
    cudaStream_t st;
    template <class Like, int D>
    int launch_reg(int G) {
        const int WPB = WALK_THREADS / G;
        const int smem = (int)(sizeof(double) * 2 * (size_t)WPB * walker_chunk(G));
        const int blocks = (P.k_max + WPB - 1) / WPB;
        int lgG = 0;
        while ((1 << lgG) < G) ++lgG;
        for (int phase = 0; phase <= P.tuning_cycles; ++phase) {
            const int nb = blocks + (phase == 0 && P.book ? BOOK_BLOCKS : 0);        // + the bookkeeping blocks
            if (G == 32) GL_CUDA(launch_pdl(walk_reg_kernel<Like, D, true>, dim3(nb), dim3(WALK_THREADS), smem, st, P, M, phase, lgG));
            else GL_CUDA(launch_pdl(walk_reg_kernel<Like, D, false>, dim3(nb), dim3(WALK_THREADS), smem, st, P, M, phase, lgG));
        }
        return GL_OK;
    }
    template <class Like>
    int operator()() {
        int G = walker_group_size(P.k_max, P.lanes);
        if (P.lanes < 1 && G < min_lanes<Like>::value) G = min_lanes<Like>::value;   // several waves beat one lane
        if constexpr (reg_max_d<Like>::value > 0) {
            if (!P.force_smem_state && P.d <= reg_max_d<Like>::value) {
                switch (P.d) {
                    case 1: return launch_reg<Like, 1>(G);
                    case 2: return launch_reg<Like, 2>(G);
                    case 3: return launch_reg<Like, 3>(G);
                    case 4: return launch_reg<Like, 4>(G);
                    case 5: return launch_reg<Like, 5>(G);
                    case 6: return launch_reg<Like, 6>(G);
                    case 7: return launch_reg<Like, 7>(G);
                    case 8: return launch_reg<Like, 8>(G);
                    default: break;
                }
            }
        }
        int smem = walker_smem_bytes(P.d, Like::cache_size(P.d), G);
        while (smem > 200 * 1024 && G < 32) {      // very large d: fewer walkers per block
            G *= 2;
            smem = walker_smem_bytes(P.d, Like::cache_size(P.d), G);
        }
        GL_REQUIRE(smem <= 227 * 1024, "walker state does not fit in shared memory (d=%d)", P.d);
        static thread_local int configured[64] = {0};      // per device: function attributes belong to a device
        int dev = 0;
        cudaGetDevice(&dev);
        int &configured_smem = configured[dev & 63];
        if (smem > 48 * 1024 && smem > configured_smem) {
            GL_CUDA(cudaFuncSetAttribute(walk_kernel<Like, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            GL_CUDA(cudaFuncSetAttribute(walk_kernel<Like, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured_smem = smem;
        }
        const int WPB = WALK_THREADS / G;
        const int blocks = (P.k_max + WPB - 1) / WPB;
        int lgG = 0;
        while ((1 << lgG) < G) ++lgG;
        for (int phase = 0; phase <= P.tuning_cycles; ++phase) {
            if (G == 32) GL_CUDA(launch_pdl(walk_kernel<Like, true>, dim3(blocks), dim3(WALK_THREADS), (size_t)smem, st, P, M, phase, lgG));
            else GL_CUDA(launch_pdl(walk_kernel<Like, false>, dim3(blocks), dim3(WALK_THREADS), (size_t)smem, st, P, M, phase, lgG));
        }
        return GL_OK;
    }
};

}  // namespace gl
