// The Nested Sampling engine: device-resident outer loop of NestedSampling.run
// (nested_sampling.py:83-213), k deaths per step.  All run state (counters, prior mass,
// evidence/information accumulators, proposal widths, batch size) lives in device memory,
// so a whole run is enqueued without host round trips.
//
// Per step:  begin -> select (8 launches, select.cu) -> record -> accumulate
//            -> walkers (tuning_cycles + 1 launches, walker.cu) -> end_walk
#include <new>
#include <vector>
#include <string>
#include <thread>
#include <stdlib.h>
#include "engine.cuh"
#include "philox.cuh"
#include <cooperative_groups.h>

namespace gl {

// ---- step begin: batch size for this step ------------------------------------------------
__global__ void begin_kernel(EngineState *S, int k_req, long long n_target, long long N) {
    long long k = n_target - S->n_dead;
    if (k > k_req) k = k_req;
    if (k > N - 1) k = N - 1;          // at least one survivor to seed from
    if (k < 0) k = 0;
    S->k_cur = (int)k;
    S->do_walk = 0;
}

// ---- record: append the batch to the dead-point log (nested_sampling.py:130-133,170-173) --
__global__ void __launch_bounds__(256) record_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                     const double *__restrict__ live,
                                                     const double *__restrict__ logl, int d, long long row_offset,
                                                     double log_np1, long long capacity, double *__restrict__ dead_x,
                                                     double *__restrict__ dead_logl, double *__restrict__ dead_w,
                                                     long long *__restrict__ dead_idx, int *__restrict__ dead_step) {
    const int k = S->k_cur;
    const long long base = S->n_rows;
    if (k <= 0 || base + k > capacity) return;
    const double w = exp(S->logX - log_np1);          // X/(N+1): every death of a batch
    const int step = S->step;
    const long long total = (long long)k * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long q = e / d;
        const int j = (int)(e - q * d);
        const long long slot = dead_list[q];
        dead_x[(base + q) * d + j] = live[slot * d + j];
        if (j == 0) {
            dead_logl[base + q] = logl[slot];
            dead_w[base + q] = w;
            dead_idx[base + q] = slot + row_offset;
            dead_step[base + q] = step;
        }
    }
}

// ---- accumulate: evidence / information update for the batch (:118-125,160-167) ----------
// k_override < 0: the batch is the engine's own (k = S->k_cur, all rows local).  Sharded run:
// k_override = global batch size, dead_list/logl index the all-gathered candidate array and the
// local row count S->k_cur was set by shard_local_kernel.
__global__ void __launch_bounds__(1024) accumulate_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                          const double *__restrict__ logl, double n_global,
                                                          long long n_target, long long capacity, int k_override) {
    __shared__ double scratch[33];
    __shared__ unsigned long long kscratch[33];
    const int k = k_override >= 0 ? k_override : S->k_cur;
    const int k_rows = S->k_cur;
    if (k <= 0) return;
    const double logw = S->logX - log(n_global + 1.0);
    const double w = exp(logw);
    double mx = -INFINITY;
    unsigned long long mk = 0ull;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double L = logl[dead_list[i]];
        const double lw = logw + L;
        if (isfinite(lw)) mx = fmax(mx, lw);
        const unsigned long long key = sort_key(L);
        mk = key > mk ? key : mk;
    }
    mx = block_max(mx, scratch);
    {   // block max of the key -> L*
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        mk = warp_max_u64(mk);
        __syncthreads();
        if (lane == 0) kscratch[wid] = mk;
        __syncthreads();
        if (wid == 0) {
            unsigned long long t = lane < nw ? kscratch[lane] : 0ull;
            t = warp_max_u64(t);
            if (lane == 0) kscratch[32] = t;
        }
        __syncthreads();
        mk = kscratch[32];
    }
    const double m_old = S->M;
    const double m_new = fmax(m_old, mx);
    double sz = 0.0, sh = 0.0, zl = 0.0, hl = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double L = logl[dead_list[i]];
        const double lw = logw + L;
        if (isfinite(lw)) {
            const double t = exp(lw - m_new);
            sz += t;
            sh += t * L;
        }
        const double dZ = w * exp(L);
        zl += dZ;
        const double dH = dZ * L;
        if (dH == dH) hl += dH;                       // isnan(dH) -> 0  (:122,164)
    }
    sz = block_sum(sz, scratch);
    sh = block_sum(sh, scratch);
    zl = block_sum(zl, scratch);
    hl = block_sum(hl, scratch);
    if (threadIdx.x == 0) {
        if (mx > -INFINITY) {
            const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
            S->Zs = S->Zs * scale + sz;
            S->Hs = S->Hs * scale + sh;
            S->M = m_new;
        }
        S->Z_lin += zl;
        S->H_lin += hl;
        S->logX += log((n_global - (double)k + 1.0) / (n_global + 1.0));
        S->n_dead += k;
        if (S->n_rows + k_rows <= capacity) S->n_rows += k_rows;
        S->lstar = key_to_double(mk);
        S->do_walk = 1;       // whether a refill follows is the host's call: a run driven in slices refills this batch in its NEXT gl_ns_advance
    }
}

// ---- sharded run helpers -------------------------------------------------------------------
// candidates: logL of this rank's k worst slots, in ascending slot order
__global__ void __launch_bounds__(256) shard_gather_kernel(const long long *__restrict__ cand_list,
                                                           const double *__restrict__ logl, int k,
                                                           double *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) out[i] = logl[cand_list[i]];
}

// Truncated exchange: every rank contributes only its m <= k worst slots.  The global selection
// over the [world*m] gathered keys is provably the true k worst iff no rank had ALL m of its
// candidates selected (a rank whose m-th worst survived cannot hide anything below the threshold).
// `positions` is ascending, so rank r's share is a contiguous run found by two binary searches.
__device__ __forceinline__ int lower_bound_pos(const long long *__restrict__ positions, int k, long long v) {
    int a = 0, b = k;
    while (a < b) { const int mid = (a + b) >> 1; if (positions[mid] < v) a = mid + 1; else b = mid; }
    return a;
}
__device__ __forceinline__ bool shard_exchange_overflowed(const long long *__restrict__ positions, int k, int m, int r) {
    const int lo = lower_bound_pos(positions, k, (long long)r * m), hi = lower_bound_pos(positions, k, (long long)(r + 1) * m);
    return hi - lo >= m;
}

// `positions` = the global k worst as ascending positions in the gathered [G*m] array (rank-major,
// slot-ascending inside a rank, i.e. global-index order).  This rank's share is the contiguous run
// of positions inside [rank*m, (rank+1)*m): translate it back to local slots, flag them dead and
// publish the local batch size.
__global__ void __launch_bounds__(256) shard_local_kernel(EngineState *S, const long long *__restrict__ positions,
                                                          const long long *__restrict__ cand_list, int k, int m,
                                                          int world, int rank,
                                                          long long *__restrict__ dead_list,
                                                          unsigned char *__restrict__ is_dead) {
    __shared__ int s_start, s_end;
    if (m < k && threadIdx.x < world && shard_exchange_overflowed(positions, k, m, threadIdx.x)) atomicMax(&S->shard_overflow, 1);
    if (threadIdx.x == 0) {
        const long long lo_pos = (long long)rank * m, hi_pos = lo_pos + m;
        int a = 0, b = k;                          // lower_bound(positions, lo_pos)
        while (a < b) { const int m = (a + b) >> 1; if (positions[m] < lo_pos) a = m + 1; else b = m; }
        s_start = a;
        b = k;
        while (a < b) { const int m = (a + b) >> 1; if (positions[m] < hi_pos) a = m + 1; else b = m; }
        s_end = a;
        S->k_cur = s_end - s_start;
        S->do_walk = 0;
    }
    __syncthreads();
    const int start = s_start, n_loc = s_end - s_start;
    for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
        const long long slot = cand_list[positions[start + i] - (long long)rank * m];
        dead_list[i] = slot;
        is_dead[slot] = 1;
    }
}

// ---- fused kill for small populations ---------------------------------------------------
// begin + select + record + accumulate in ONE single-CTA launch when the N ordering keys fit
// in shared memory (N <= KILL_SMALL_MAX_N): the keys are loaded once, the six radix-select
// passes (4-bit digits counted in registers and reduced with warp REDUX -- shared-memory atomics
// cost ~2 cycles per lane on this part and made an 11-bit histogram version 4x slower), the
// ballot-based index-ordered compaction, the dead-row copy and the evidence update all run out
// of shared memory.  Same arithmetic and same (value, index) order as the multi-launch path
// (select.cu + record_kernel + accumulate_kernel); replaces 11 launches (~90 us per step at
// N = 10^4, r01 baseline launch list) by one.
constexpr int KILL_SMALL_THREADS = 1024;
constexpr int KILL_SMALL_MAX_N = 24 * 1024;      // 192 KB of keys + 8 KB histogram + scratch

// Single-CTA (1024 threads) selection of the k smallest of keys[0..N) -- ordering keys held in
// shared memory -- under (key, index) order.  Writes the selected indices in ascending order to
// out_list[0..k), optionally flags them in is_dead, and returns the largest selected key to
// every thread.  `hist` is 2048 words of shared scratch.
__device__ unsigned long long cta_select_kworst(const unsigned long long *keys, int N, int k, unsigned int *hist,
                                                long long *__restrict__ dead_list,
                                                unsigned char *__restrict__ is_dead) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = KILL_SMALL_THREADS / 32;
    // ---- radix select on 4-bit digits, NO atomics ------------------------------------------------
    // (1) min/max key -> the leading nibbles they share are skipped; (2) per pass each thread counts
    // its candidates' digits in two 64-bit words of 8-bit counters (<= 255 keys per thread), widened
    // to 16-bit lanes and summed with warp REDUX, one thread per digit sums the 32 warps, and every
    // thread scans the 16 totals redundantly (prefix / k_rem stay in registers: two barriers per
    // pass); (3) the loop stops as soon as the boundary bucket is needed in full (always true once
    // it holds one key), so a typical live set needs 3-5 passes, not 16.
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (int i = tid; i < N; i += KILL_SMALL_THREADS) {
        const unsigned long long key = keys[i];
        kmin = key < kmin ? key : kmin;
        kmax = key > kmax ? key : kmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b2 = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b2 > kmax ? b2 : kmax;
    }
    unsigned long long *mm = reinterpret_cast<unsigned long long *>(hist);      // [2*NW] scratch
    if (lane == 0) { mm[wid] = kmin; mm[NW + wid] = kmax; }
    __syncthreads();
    for (int w = 0; w < NW; ++w) {
        const unsigned long long a = mm[w], b2 = mm[NW + w];
        kmin = a < kmin ? a : kmin;
        kmax = b2 > kmax ? b2 : kmax;
    }
    __syncthreads();
    const unsigned long long diff = kmin ^ kmax;
    const int first_pass = diff ? __clzll((long long)diff) >> 2 : 16;            // nibbles shared by all keys
    unsigned long long prefix = first_pass ? kmin >> (64 - 4 * first_pass) : 0ull;
    int k_rem = k;
    int sel_shift = diff ? 0 : 0;              // granularity of the final comparison (bits dropped)
    if (!diff) prefix = kmin;                  // every key equal: lowest indices die
    for (int pass = first_pass; pass < 16; ++pass) {
        const int shift = 60 - 4 * pass;
        unsigned long long lo8 = 0ull, hi8 = 0ull;     // 8-bit counters: digits 0-7, 8-15
        for (int i = tid; i < N; i += KILL_SMALL_THREADS) {
            const unsigned long long key = keys[i];
            if (pass == 0 || (key >> (shift + 4)) == prefix) {
                const unsigned int dg = (unsigned int)(key >> shift) & 15u;
                const unsigned long long inc = 1ull << ((dg & 7u) << 3);
                if (dg < 8u) lo8 += inc; else hi8 += inc;
            }
        }
        // widen to 16-bit lanes: pk[q] = (count of digit 2q) | (count of digit 2q+1) << 16
        unsigned int pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pk[q] = (unsigned int)((lo8 >> (16 * q)) & 0xffull) | ((unsigned int)((lo8 >> (16 * q + 8)) & 0xffull) << 16);
            pk[4 + q] = (unsigned int)((hi8 >> (16 * q)) & 0xffull) | ((unsigned int)((hi8 >> (16 * q + 8)) & 0xffull) << 16);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) pk[q] = __reduce_add_sync(0xffffffffu, pk[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) hist[wid * 8 + q] = pk[q];
        }
        __syncthreads();
        if (tid < 16) {
            unsigned int tot = 0u;
            for (int w = 0; w < NW; ++w) tot += (hist[w * 8 + (tid >> 1)] >> ((tid & 1) << 4)) & 0xffffu;
            hist[NW * 8 + tid] = tot;
        }
        __syncthreads();
        int run = 0, digit = 15, run_at = 0, c_at = 0;
        bool found = false;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int cq = (int)hist[NW * 8 + q];
            if (!found && run + cq >= k_rem) { found = true; digit = q; run_at = run; c_at = cq; }
            run += cq;
        }
        prefix = (prefix << 4) | (unsigned long long)digit;
        k_rem -= run_at;
        sel_shift = shift;
        __syncthreads();                           // hist is rewritten by the next pass
        if (c_at == k_rem) break;                  // the whole boundary bucket dies: no finer digits needed
    }
    // selection: (key >> sel_shift) < prefix, plus the first r (by slot) of those == prefix
    const unsigned long long T = prefix;
    const int r = k_rem;

    // ---- index-ordered compaction -> dead_list (ascending slot), is_dead flags -----------------
    // warp w owns the contiguous slots [w*C, (w+1)*C): count, exchange totals, then place with
    // ballot prefixes -- one barrier, no atomics.  Also finds L* = largest selected key.
    const int C = (N + NW - 1) / NW;
    const int lo = wid * C, hi = lo + C < N ? lo + C : N;
    int my_lt = 0, my_eq = 0;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        const unsigned long long kh = i < hi ? keys[i] >> sel_shift : ~0ull;
        my_lt += __popc(__ballot_sync(0xffffffffu, i < hi && kh < T));
        my_eq += __popc(__ballot_sync(0xffffffffu, i < hi && kh == T));
    }
    if (lane == 0) { hist[wid] = (unsigned)my_lt; hist[NW + wid] = (unsigned)my_eq; }
    __syncthreads();
    int run_lt = 0, run_eq = 0;
    for (int w = 0; w < wid; ++w) { run_lt += (int)hist[w]; run_eq += (int)hist[NW + w]; }
    const unsigned lanemask_lt = (1u << lane) - 1u;
    unsigned long long kstar = 0ull;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        const unsigned long long key = i < hi ? keys[i] : ~0ull;
        const unsigned long long kh = i < hi ? key >> sel_shift : ~0ull;
        const bool is_lt = i < hi && kh < T, is_eq = i < hi && kh == T;
        const unsigned b_lt = __ballot_sync(0xffffffffu, is_lt), b_eq = __ballot_sync(0xffffffffu, is_eq);
        const int lt_before = run_lt + __popc(b_lt & lanemask_lt);
        const int eq_before = run_eq + __popc(b_eq & lanemask_lt);
        if (is_lt || (is_eq && eq_before < r)) {
            const int pos = lt_before + (eq_before < r ? eq_before : r);
            dead_list[pos] = i;
            if (is_dead) is_dead[i] = 1;
            kstar = key > kstar ? key : kstar;
        }
        run_lt += __popc(b_lt);
        run_eq += __popc(b_eq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kstar, o);
        kstar = a > kstar ? a : kstar;
    }
    __syncthreads();                               // hist (totals) fully consumed
    if (lane == 0) mm[wid] = kstar;
    __syncthreads();
    for (int w = 0; w < NW; ++w) {
        const unsigned long long a = mm[w];
        kstar = a > kstar ? a : kstar;
    }

    return kstar;
}

// ---- incrementally sorted live set (optional path 3; NOT the default, see kill_launch) -------------
// Only the k slots refilled by the last batch change between two kills, so the live set is kept
// sorted by (key, slot) in two global (L2-resident) buffers: the refilled entries are sorted with a
// bitonic network in shared memory and merge-pathed into the surviving run; the k worst are then
// simply the FRONT of the array -- no pass over all N keys, no histogram.  Same (key, slot) order
// as the radix paths, and the selected slots are re-sorted ascending so everything downstream
// (record order, accumulation order, walker assignment) is bit-identical to them.
constexpr int SORTED_MAX_N = 16384;

__device__ __forceinline__ bool pair_less(unsigned long long ka, unsigned int ia, unsigned long long kb,
                                          unsigned int ib) {
    return ka < kb || (ka == kb && ia < ib);
}

// Bitonic sort of 1024*E (key, slot) pairs, ascending, with element i = e*1024 + tid held in
// registers: partners at strides >= 1024 are other registers of the same thread, strides < 32 are
// exchanged with warp shuffles, and only strides 32..512 go through shared memory (15 barrier
// stages for 1024 elements instead of 55).  sk/sv: 1024*E entries of shared scratch.
template <int E>
__device__ void cta_bitonic_regs(unsigned long long (&k)[E], unsigned int (&v)[E], unsigned long long *sk,
                                 unsigned int *sv) {
    const int tid = threadIdx.x;
    constexpr int T = KILL_SMALL_THREADS;
#pragma unroll 1
    for (int size = 2; size <= T * E; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= T) {                               // register <-> register, same thread
                const int es = stride / T;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & es) == 0 && (e | es) < E) {
                        const int i = e * T + tid;
                        const bool up = (i & size) == 0;
                        const int f = e | es;
                        if (pair_less(k[f], v[f], k[e], v[e]) == up) {
                            const unsigned long long tk = k[e]; k[e] = k[f]; k[f] = tk;
                            const unsigned int tv = v[e]; v[e] = v[f]; v[f] = tv;
                        }
                    }
                }
            } else if (stride >= 32) {                       // through shared memory
#pragma unroll
                for (int e = 0; e < E; ++e) { sk[e * T + tid] = k[e]; sv[e * T + tid] = v[e]; }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * T + tid, j = i ^ stride;
                    const unsigned long long pk = sk[j];
                    const unsigned int pv = sv[j];
                    const bool keep_min = ((i & stride) == 0) == ((i & size) == 0);
                    const bool p_less = pair_less(pk, pv, k[e], v[e]);
                    if (p_less == keep_min) { k[e] = pk; v[e] = pv; }
                }
                __syncthreads();
            } else {                                         // inside the warp
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * T + tid;
                    const unsigned long long pk = __shfl_xor_sync(0xffffffffu, k[e], stride);
                    const unsigned int pv = __shfl_xor_sync(0xffffffffu, v[e], stride);
                    const bool keep_min = ((i & stride) == 0) == ((i & size) == 0);
                    const bool p_less = pair_less(pk, pv, k[e], v[e]);
                    if (p_less == keep_min) { k[e] = pk; v[e] = pv; }
                }
            }
        }
    }
}

// Load the nB new (key, slot) pairs, sort them, leave them sorted in shared memory Bk/Bi[0..1024*E).
template <int E>
__device__ void sorted_load_and_sort(const double *__restrict__ logl, const long long *__restrict__ dead_list,
                                     bool ready, int nB, unsigned long long *Bk, unsigned int *Bi) {
    constexpr int T = KILL_SMALL_THREADS;
    unsigned long long k[E];
    unsigned int v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * T + threadIdx.x;
        if (i < nB) {
            const unsigned int slot = ready ? (unsigned int)dead_list[i] : (unsigned int)i;
            k[e] = sort_key(logl[slot]);
            v[e] = slot;
        } else {
            k[e] = ~0ull;
            v[e] = 0xffffffffu;
        }
    }
    cta_bitonic_regs<E>(k, v, Bk, Bi);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) { Bk[e * T + threadIdx.x] = k[e]; Bi[e * T + threadIdx.x] = v[e]; }
    __syncthreads();
}

// Returns the largest selected key; writes dead_list (ascending slot) and is_dead.
// smem_raw: 12 B x (1024*E) for the sorted new entries, then N bytes of dead flags.
__device__ unsigned long long cta_sorted_select(EngineState *S, const double *__restrict__ logl, int N, int k,
                                                unsigned long long *sk0, unsigned long long *sk1, unsigned int *si0,
                                                unsigned int *si1, unsigned char *smem_raw, int n2,
                                                long long *__restrict__ dead_list,
                                                unsigned char *__restrict__ is_dead) {
    constexpr int T = KILL_SMALL_THREADS, NW = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool ready = S->sorted_ready != 0;
    const int cur = S->sort_cur;
    const int nB = ready ? S->front_dead : N;           // entries whose logL is new
    const int nA = N - nB;
    const unsigned long long *Ak = (cur ? sk1 : sk0) + (N - nA);
    const unsigned int *Ai = (cur ? si1 : si0) + (N - nA);
    unsigned long long *Ok = cur ? sk0 : sk1;
    unsigned int *Oi = cur ? si0 : si1;
    unsigned long long *Bk = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned int *Bi = reinterpret_cast<unsigned int *>(Bk + n2);
    unsigned char *flags = reinterpret_cast<unsigned char *>(Bi + n2);          // [N]
    __shared__ int s_cnt[NW];
    switch (n2 / T) {
        case 1: sorted_load_and_sort<1>(logl, dead_list, ready, nB, Bk, Bi); break;
        case 2: sorted_load_and_sort<2>(logl, dead_list, ready, nB, Bk, Bi); break;
        case 4: sorted_load_and_sort<4>(logl, dead_list, ready, nB, Bk, Bi); break;
        case 8: sorted_load_and_sort<8>(logl, dead_list, ready, nB, Bk, Bi); break;
        default: sorted_load_and_sort<16>(logl, dead_list, ready, nB, Bk, Bi); break;
    }
    // ---- merge: thread t owns a contiguous chunk of the surviving run A and the new entries that
    // fall inside it (binary searches only touch shared memory; the A loads are independent) -----
    if (nA == 0) {
        for (int i = tid; i < N; i += T) { Ok[i] = Bk[i]; Oi[i] = Bi[i]; }
    } else {
        constexpr int PA = SORTED_MAX_N / T;              // 16: max chunk of A per thread
        const int per = (nA + T - 1) / T;
        const int a_lo = tid * per < nA ? tid * per : nA, a_hi = a_lo + per < nA ? a_lo + per : nA;
        unsigned long long ak[PA + 1];
        unsigned int ai[PA + 1];
#pragma unroll
        for (int q = 0; q <= PA; ++q) {
            const int a = a_lo + q;
            if (q <= per && a < nA && a <= a_hi) { ak[q] = Ak[a]; ai[q] = Ai[a]; }
            else { ak[q] = ~0ull; ai[q] = 0xffffffffu; }
        }
        if (a_lo < a_hi) {
            auto lower_bound_B = [&](unsigned long long key, unsigned int idx) {   // # of B entries < (key, idx)
                int lo = 0, hi = nB;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (pair_less(Bk[mid], Bi[mid], key, idx)) lo = mid + 1;
                    else hi = mid;
                }
                return lo;
            };
            int b = tid == 0 ? 0 : lower_bound_B(ak[0], ai[0]);
            const int n_own = a_hi - a_lo;
            const int b_hi = a_hi < nA ? lower_bound_B(ak[n_own], ai[n_own]) : nB;
            int o = a_lo + b, q = 0;
            while (q < n_own || b < b_hi) {
                bool takeA;
                if (b >= b_hi) takeA = true;
                else if (q >= n_own) takeA = false;
                else takeA = pair_less(ak[q], ai[q], Bk[b], Bi[b]);
                if (takeA) {
                    unsigned long long kk = ak[0];
                    unsigned int ii = ai[0];
#pragma unroll
                    for (int z = 1; z < PA; ++z) if (z == q) { kk = ak[z]; ii = ai[z]; }   // static register select
                    Ok[o] = kk; Oi[o] = ii; ++q;
                } else { Ok[o] = Bk[b]; Oi[o] = Bi[b]; ++b; }
                ++o;
            }
        }
    }
    for (int i = tid; i < N; i += T) flags[i] = 0;
    __syncthreads();                                    // Ok/Oi visible block-wide; flags cleared
    // ---- the k worst are the front of the merged array: flag them, then emit them in slot order --
    for (int i = tid; i < k; i += T) flags[Oi[i]] = 1;
    const unsigned long long kstar = Ok[k - 1];
    __syncthreads();
    const int C = (N + NW - 1) / NW;
    const int lo = wid * C, hi = lo + C < N ? lo + C : N;
    int cnt = 0;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        cnt += __popc(__ballot_sync(0xffffffffu, i < hi && flags[i]));
    }
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    int run = 0;
    for (int w = 0; w < wid; ++w) run += s_cnt[w];
    const unsigned lanemask_lt = (1u << lane) - 1u;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        const bool f = i < hi && flags[i];
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (f) {
            dead_list[run + __popc(bal & lanemask_lt)] = i;
            is_dead[i] = 1;
        }
        run += __popc(bal);
    }
    __syncthreads();
    if (tid == 0) {
        S->sorted_ready = 1;
        S->sort_cur = cur ^ 1;
        S->front_dead = k;
    }
    return kstar;
}

template <bool SORTED>
__global__ void __launch_bounds__(KILL_SMALL_THREADS) kill_small_kernel(
    EngineState *S, int end_walk_tc /* >= 0: first finish the previous refill (end_walk_kernel's job) */,
    long long evals_per_walker, int k_req, long long n_target, int N, int d, long long row_offset, double n_global,
    long long capacity, const double *__restrict__ live, const double *__restrict__ logl,
    long long *__restrict__ dead_list, unsigned char *__restrict__ is_dead, double *__restrict__ dead_x,
    double *__restrict__ dead_logl, double *__restrict__ dead_w, long long *__restrict__ dead_idx,
    int *__restrict__ dead_step, unsigned long long *sk0, unsigned long long *sk1, unsigned int *si0,
    unsigned int *si1, int sort_n2 /* entries of bitonic scratch: 1024 * 2^e */) {
    extern __shared__ unsigned long long keys[];                 // radix: [N] keys; sorted: bitonic scratch
    __shared__ unsigned int hist[SORTED ? 64 : SEL_BINS];       // (radix scratch; the sorted path needs none)
    __shared__ double scratch[33];
    const int tid = threadIdx.x;

    // ---- epilogue of the previous refill, folded in (saves a launch per step) -----------------
    if (end_walk_tc >= 0) {
        const int kp = S->k_cur;
        for (int i = tid; i < kp; i += KILL_SMALL_THREADS) is_dead[dead_list[i]] = 0;
        const double wnew = tid < d ? phase_width(S, tid, end_walk_tc, kp) : 0.0;
        __syncthreads();
        if (tid < d) {
            S->widths[tid] = wnew;
            for (int c = 0; c < end_walk_tc; ++c) S->acc[c][tid] = 0u;
        }
        if (tid == 0) {
            S->n_evals += (unsigned long long)((long long)kp * evals_per_walker);
            S->step += 1;
        }
        __syncthreads();
    }

    // ---- begin: batch size -------------------------------------------------------------------
    long long kk = n_target - S->n_dead;
    if (kk > k_req) kk = k_req;
    if (kk > N - 1) kk = N - 1;
    if (kk < 0) kk = 0;
    const int k = (int)kk;
    if (tid == 0) { S->k_cur = k; S->do_walk = 0; }
    if (k <= 0) return;

    unsigned long long kstar;
    if (SORTED) {
        kstar = cta_sorted_select(S, logl, N, k, sk0, sk1, si0, si1, reinterpret_cast<unsigned char *>(keys), sort_n2,
                                  dead_list, is_dead);
    } else {
        for (int i = tid; i < N; i += KILL_SMALL_THREADS) keys[i] = sort_key(logl[i]);
        __syncthreads();
        kstar = cta_select_kworst(keys, N, k, hist, dead_list, is_dead);
    }
    __syncthreads();

    // ---- record (nested_sampling.py:130-133,170-173) --------------------------------------------
    const long long base_row = S->n_rows;
    __shared__ double s_logw, s_w;                  // scalar math once, not in all 1024 threads
    if (tid == 0) {
        s_logw = S->logX - log(n_global + 1.0);
        s_w = exp(s_logw);
    }
    __syncthreads();
    const double logw = s_logw, w = s_w;
    const bool room = base_row + k <= capacity;
    if (room) {
        const int step = S->step;
        for (int e = tid; e < k * d; e += KILL_SMALL_THREADS) {
            const int q = e / d, j = e - q * d;
            const long long slot = dead_list[q];
            dead_x[(base_row + q) * d + j] = live[slot * d + j];
            if (j == 0) {
                dead_logl[base_row + q] = logl[slot];
                dead_w[base_row + q] = w;
                dead_idx[base_row + q] = slot + row_offset;
                dead_step[base_row + q] = step;
            }
        }
    }

    // ---- accumulate (nested_sampling.py:118-125,160-167); same forms as accumulate_kernel -------
    double mx = -INFINITY;
    for (int i = tid; i < k; i += KILL_SMALL_THREADS) {
        const double lw = logw + logl[dead_list[i]];
        if (isfinite(lw)) mx = fmax(mx, lw);
    }
    mx = block_max(mx, scratch);
    const double m_old = S->M;
    const double m_new = fmax(m_old, mx);
    double sz = 0.0, sh = 0.0, zl = 0.0, hl = 0.0;
    for (int i = tid; i < k; i += KILL_SMALL_THREADS) {
        const double L = logl[dead_list[i]];
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
    sz = block_sum(sz, scratch);
    sh = block_sum(sh, scratch);
    zl = block_sum(zl, scratch);
    hl = block_sum(hl, scratch);
    if (tid == 0) {
        if (mx > -INFINITY) {
            const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
            S->Zs = S->Zs * scale + sz;
            S->Hs = S->Hs * scale + sh;
            S->M = m_new;
        }
        S->Z_lin += zl;
        S->H_lin += hl;
        S->logX += log((n_global - (double)k + 1.0) / (n_global + 1.0));
        S->n_dead += k;
        if (room) S->n_rows += k;
        S->lstar = key_to_double(kstar);
        S->do_walk = 1;       // whether a refill follows is the host's call: a run driven in slices refills this batch in its NEXT gl_ns_advance
    }
}

// ---- fused kill on a thread-block cluster -------------------------------------------------------
// The single-CTA kill is bound by the instruction issue rate of the ONE SM it runs on (r01_d
// profile: 124 k warp-instructions, 32 us).  Here a cluster of KC_CTAS CTAs splits the live set into
// contiguous chunks (keys in each CTA's shared memory); every radix pass exchanges the 16 digit
// counts of each CTA through distributed shared memory (one cluster barrier, ~0.2 us), every CTA
// takes the same decision, the index-ordered compaction gets its global offsets from an exclusive
// prefix over the CTA totals, the dead-row copy is spread over the cluster, and CTA 0 does the
// evidence update with the SAME summation tree as the 1024-thread single-CTA kernel (two virtual
// threads per thread), so every selection path leaves bit-identical state.
constexpr int KC_CTAS = 8;                      // portable cluster size
constexpr int KC_THREADS = 512;
constexpr int KC_NW = KC_THREADS / 32;
constexpr int KC_MAX_N = KC_CTAS * 24 * 1024;   // 192 KB of keys per CTA; <= 48 keys per thread (8-bit counters)

struct ClusterXchg {                            // per-CTA mailbox, read by the other CTAs via DSMEM
    unsigned long long kmin, kmax, kstar;
    unsigned int hist[2][16];
    unsigned int lt, eq;
};

// deterministic sums/max over 1024 virtual threads held two per thread (virtual tid, tid + 512):
// identical tree to block_sum / block_max of a 1024-thread block
__device__ __forceinline__ double block_sum_v2(double a, double b, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { scratch[wid] = a; scratch[KC_NW + wid] = b; }
    __syncthreads();
    if (wid == 0) {
        double t = warp_sum(scratch[lane]);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ double block_max_v2(double a, double b, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_max(a);
    b = warp_max(b);
    __syncthreads();
    if (lane == 0) { scratch[wid] = a; scratch[KC_NW + wid] = b; }
    __syncthreads();
    if (wid == 0) {
        double t = warp_max(scratch[lane]);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

struct ClusterSelSmem {                         // static shared memory of a cluster-select kernel
    ClusterXchg X;
    unsigned int wh[KC_NW * 8];                 // per-warp packed digit counts / lt,eq counts
    unsigned int tot[16];
    unsigned long long mm[2 * KC_NW];
    unsigned long long bc[4];                   // broadcast slots
};

// Cluster-wide selection of the k smallest of `n` ordering keys under (key, index) order.  CTA c
// holds the keys of global indices [c*Nc, c*Nc + n_loc) in `keys` (its shared memory), Nc =
// ceil(n / KC_CTAS).  The selected global indices go to out_list[0..k) ascending (global memory),
// optionally flagged in is_dead; the largest selected key is returned to every thread of every
// CTA.  On return (a cluster barrier has been passed) out_list is complete and visible
// cluster-wide.  Same decisions as cta_select_kworst.
__device__ unsigned long long cluster_select_kworst(cooperative_groups::cluster_group &cluster, ClusterSelSmem &sm,
                                                    const unsigned long long *keys, int n, int k,
                                                    long long *__restrict__ out_list,
                                                    unsigned char *__restrict__ is_dead) {
    const int c = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    ClusterXchg &X = sm.X;
    unsigned int *wh = sm.wh, *tot = sm.tot;
    unsigned long long *mm = sm.mm, *bc = sm.bc;
    const int Nc = (n + KC_CTAS - 1) / KC_CTAS;
    const int lo_g = c * Nc;
    const int n_loc = n - lo_g < 0 ? 0 : (n - lo_g < Nc ? n - lo_g : Nc);
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (int i = tid; i < n_loc; i += KC_THREADS) {
        const unsigned long long key = keys[i];
        kmin = key < kmin ? key : kmin;
        kmax = key > kmax ? key : kmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b2 = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b2 > kmax ? b2 : kmax;
    }
    if (lane == 0) { mm[wid] = kmin; mm[KC_NW + wid] = kmax; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 0; w < KC_NW; ++w) {
            const unsigned long long a = mm[w], b2 = mm[KC_NW + w];
            kmin = a < kmin ? a : kmin;
            kmax = b2 > kmax ? b2 : kmax;
        }
        X.kmin = kmin;
        X.kmax = kmax;
    }
    cluster.sync();
    if (wid == 0) {
        unsigned long long a = ~0ull, b2 = 0ull;
        if (lane < KC_CTAS) {
            const ClusterXchg *R = cluster.map_shared_rank(&X, lane);
            a = R->kmin;
            b2 = R->kmax;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o), b3 = __shfl_xor_sync(0xffffffffu, b2, o);
            a = a2 < a ? a2 : a;
            b2 = b3 > b2 ? b3 : b2;
        }
        if (lane == 0) { bc[0] = a; bc[1] = b2; }
    }
    __syncthreads();
    kmin = bc[0];
    kmax = bc[1];

    // ---- radix select, 4-bit digits ---------------------------------------------------------------
    const unsigned long long diff = kmin ^ kmax;
    const int first_pass = diff ? __clzll((long long)diff) >> 2 : 16;
    unsigned long long prefix = first_pass ? kmin >> (64 - 4 * first_pass) : 0ull;
    int k_rem = k;
    int sel_shift = 0;
    if (!diff) prefix = kmin;
    for (int pass = first_pass; pass < 16; ++pass) {
        const int shift = 60 - 4 * pass, par = pass & 1;
        unsigned long long lo8 = 0ull, hi8 = 0ull;
        for (int i = tid; i < n_loc; i += KC_THREADS) {
            const unsigned long long key = keys[i];
            if (pass == 0 || (key >> (shift + 4)) == prefix) {
                const unsigned int dg = (unsigned int)(key >> shift) & 15u;
                const unsigned long long inc = 1ull << ((dg & 7u) << 3);
                if (dg < 8u) lo8 += inc; else hi8 += inc;
            }
        }
        unsigned int pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pk[q] = (unsigned int)((lo8 >> (16 * q)) & 0xffull) | ((unsigned int)((lo8 >> (16 * q + 8)) & 0xffull) << 16);
            pk[4 + q] = (unsigned int)((hi8 >> (16 * q)) & 0xffull) | ((unsigned int)((hi8 >> (16 * q + 8)) & 0xffull) << 16);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) pk[q] = __reduce_add_sync(0xffffffffu, pk[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) wh[wid * 8 + q] = pk[q];
        }
        __syncthreads();
        if (tid < 16) {
            unsigned int t = 0u;
#pragma unroll
            for (int w = 0; w < KC_NW; ++w) t += (wh[w * 8 + (tid >> 1)] >> ((tid & 1) << 4)) & 0xffffu;
            X.hist[par][tid] = t;
        }
        cluster.sync();
        if (wid == 0) {
            // one warp sums the cluster's counts per digit (lane = digit), scans them with shuffles and
            // publishes the decision: first digit whose running count reaches k_rem, the count below it
            // and the bucket size (a redundant 16-step scan in every thread was 37 % of a pass)
            unsigned int t = 0u;
            if (lane < 16) {
#pragma unroll
                for (int r = 0; r < KC_CTAS; ++r) t += cluster.map_shared_rank(&X, r)->hist[par][lane];
            }
            unsigned int incl = t;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned int hit = __ballot_sync(0xffffffffu, lane < 16 && (int)incl >= k_rem);
            const int dg = hit ? __ffs(hit) - 1 : 15;
            const unsigned int incl_d = __shfl_sync(0xffffffffu, incl, dg), t_d = __shfl_sync(0xffffffffu, t, dg);
            if (lane == 0) {
                tot[0] = (unsigned int)dg;
                tot[1] = incl_d - t_d;
                tot[2] = t_d;
            }
        }
        __syncthreads();
        const int digit = (int)tot[0], run_at = (int)tot[1], c_at = (int)tot[2];
        prefix = (prefix << 4) | (unsigned long long)digit;
        k_rem -= run_at;
        sel_shift = shift;
        if (c_at == k_rem) break;
    }
    const unsigned long long T = prefix;
    const int r_ties = k_rem;

    // ---- index-ordered compaction: warp w owns local slots [w*C, (w+1)*C) ---------------------------
    const int C = (Nc + KC_NW - 1) / KC_NW;
    const int lo = wid * C, hi = lo + C < n_loc ? lo + C : n_loc;
    int my_lt = 0, my_eq = 0;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        const unsigned long long kh = i < hi ? keys[i] >> sel_shift : ~0ull;
        my_lt += __popc(__ballot_sync(0xffffffffu, i < hi && kh < T));
        my_eq += __popc(__ballot_sync(0xffffffffu, i < hi && kh == T));
    }
    if (lane == 0) { wh[wid] = (unsigned)my_lt; wh[KC_NW + wid] = (unsigned)my_eq; }
    __syncthreads();
    if (tid == 0) {
        unsigned int a = 0u, b2 = 0u;
        for (int w = 0; w < KC_NW; ++w) { a += wh[w]; b2 += wh[KC_NW + w]; }
        X.lt = a;
        X.eq = b2;
    }
    cluster.sync();
    if (tid == 0) {
        unsigned int a = 0u, b2 = 0u;
        for (int r = 0; r < c; ++r) {
            const ClusterXchg *R = cluster.map_shared_rank(&X, r);
            a += R->lt;
            b2 += R->eq;
        }
        tot[0] = a;
        tot[1] = b2;
    }
    __syncthreads();
    int run_lt = (int)tot[0], run_eq = (int)tot[1];
    for (int w = 0; w < wid; ++w) { run_lt += (int)wh[w]; run_eq += (int)wh[KC_NW + w]; }
    const unsigned lanemask_lt = (1u << lane) - 1u;
    unsigned long long kstar = 0ull;
    for (int t0 = lo; t0 < hi; t0 += 32) {
        const int i = t0 + lane;
        const unsigned long long key = i < hi ? keys[i] : ~0ull;
        const unsigned long long kh = i < hi ? key >> sel_shift : ~0ull;
        const bool is_lt = i < hi && kh < T, is_eq = i < hi && kh == T;
        const unsigned b_lt = __ballot_sync(0xffffffffu, is_lt), b_eq = __ballot_sync(0xffffffffu, is_eq);
        const int lt_before = run_lt + __popc(b_lt & lanemask_lt);
        const int eq_before = run_eq + __popc(b_eq & lanemask_lt);
        if (is_lt || (is_eq && eq_before < r_ties)) {
            const int pos = lt_before + (eq_before < r_ties ? eq_before : r_ties);
            out_list[pos] = lo_g + i;
            if (is_dead) is_dead[lo_g + i] = 1;
            kstar = key > kstar ? key : kstar;
        }
        run_lt += __popc(b_lt);
        run_eq += __popc(b_eq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kstar, o);
        kstar = a > kstar ? a : kstar;
    }
    if (lane == 0) mm[wid] = kstar;
    __syncthreads();
    if (tid == 0) {
        for (int w = 0; w < KC_NW; ++w) { const unsigned long long a = mm[w]; kstar = a > kstar ? a : kstar; }
        X.kstar = kstar;
    }
    cluster.sync();                                              // out_list complete and visible cluster-wide
    if (wid == 0) {
        unsigned long long a = lane < KC_CTAS ? cluster.map_shared_rank(&X, lane)->kstar : 0ull;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o);
            a = a2 > a ? a2 : a;
        }
        if (lane == 0) bc[2] = a;
    }
    __syncthreads();
    return bc[2];
}

// evidence / information update of one batch by ONE 512-thread CTA with the summation tree of the
// 1024-thread kernels (accumulate_kernel, kill_small_kernel): L(i) = vals[idx[i]], i < k
__device__ void cta_accumulate_v2(EngineState *S, const long long *__restrict__ idx, const double *__restrict__ vals,
                                  int k, double logw, double w, double m_old, double *scratch, double &mx_out,
                                  double &m_new_out, double (&sums)[4]) {
    const int tid = threadIdx.x;
    double mxa = -INFINITY, mxb = -INFINITY;
    for (int i = tid; i < k; i += 2 * KC_THREADS) {
        const double lw = logw + vals[idx[i]];
        if (isfinite(lw)) mxa = fmax(mxa, lw);
    }
    for (int i = tid + KC_THREADS; i < k; i += 2 * KC_THREADS) {
        const double lw = logw + vals[idx[i]];
        if (isfinite(lw)) mxb = fmax(mxb, lw);
    }
    const double mx = block_max_v2(mxa, mxb, scratch);
    const double m_new = fmax(m_old, mx);
    double sz[2] = {0.0, 0.0}, sh[2] = {0.0, 0.0}, zl[2] = {0.0, 0.0}, hl[2] = {0.0, 0.0};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        for (int i = tid + v * KC_THREADS; i < k; i += 2 * KC_THREADS) {
            const double L = vals[idx[i]];
            const double lw = logw + L;
            if (isfinite(lw)) {
                const double t = exp(lw - m_new);
                sz[v] += t;
                sh[v] += t * L;
            }
            const double dZ = w * exp(L);
            zl[v] += dZ;
            const double dH = dZ * L;
            if (dH == dH) hl[v] += dH;
        }
    }
    sums[0] = block_sum_v2(sz[0], sz[1], scratch);
    sums[1] = block_sum_v2(sh[0], sh[1], scratch);
    sums[2] = block_sum_v2(zl[0], zl[1], scratch);
    sums[3] = block_sum_v2(hl[0], hl[1], scratch);
    mx_out = mx;
    m_new_out = m_new;
}

__global__ void __cluster_dims__(KC_CTAS, 1, 1) __launch_bounds__(KC_THREADS) kill_cluster_kernel(
    EngineState *S, int end_walk_tc, long long evals_per_walker, int k_req, long long n_target, int N, int d,
    long long row_offset, double n_global, long long capacity, const double *__restrict__ live,
    const double *__restrict__ logl, long long *__restrict__ dead_list, unsigned char *__restrict__ is_dead,
    double *__restrict__ dead_x, double *__restrict__ dead_logl, double *__restrict__ dead_w,
    long long *__restrict__ dead_idx, int *__restrict__ dead_step,
    int defer_book /* a batch that will be refilled leaves record + accumulate to the first walker launch */) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ unsigned long long keys[];                 // this CTA's chunk of ordering keys
    __shared__ ClusterSelSmem sm;
    __shared__ double scratch[33];
    const int tid = threadIdx.x;
    pdl_wait();                                                  // launched with launch_pdl: the refill before us
    pdl_trigger();                                               // must be complete; the next launch may queue up

    // ---- snapshot of the state every CTA needs (CTA 0 rewrites S only at the very end) -----------
    const int kp = S->k_cur;
    const int step_now = S->step + (end_walk_tc >= 0 ? 1 : 0);
    const long long n_dead0 = S->n_dead, base_row = S->n_rows;
    const double logX0 = S->logX, m_old = S->M;

    // ---- epilogue of the previous refill (end_walk_kernel's job), spread over the cluster ---------
    if (end_walk_tc >= 0) {
        for (int i = c * KC_THREADS + tid; i < kp; i += KC_CTAS * KC_THREADS) is_dead[dead_list[i]] = 0;
        if (c == 0) {
            const double wnew = tid < d ? phase_width(S, tid, end_walk_tc, kp) : 0.0;
            __syncthreads();
            if (tid < d) {
                S->widths[tid] = wnew;
                for (int q = 0; q < end_walk_tc; ++q) S->acc[q][tid] = 0u;
            }
        }
    }
    long long kk = n_target - n_dead0;
    if (kk > k_req) kk = k_req;
    if (kk > N - 1) kk = N - 1;
    if (kk < 0) kk = 0;
    const int k = (int)kk;
    if (k <= 0) {
        if (c == 0 && tid == 0) {
            if (end_walk_tc >= 0) { S->n_evals += (unsigned long long)((long long)kp * evals_per_walker); S->step = step_now; }
            S->k_cur = 0;
            S->do_walk = 0;
        }
        return;
    }

    // ---- this CTA's chunk of keys, cluster-wide selection ----------------------------------------------
    const int Nc = (N + KC_CTAS - 1) / KC_CTAS;
    const int lo_g = c * Nc;
    const int n_loc = N - lo_g < 0 ? 0 : (N - lo_g < Nc ? N - lo_g : Nc);
    for (int i = tid; i < n_loc; i += KC_THREADS) keys[i] = sort_key(logl[lo_g + i]);
    __syncthreads();
    // (the first cluster barrier inside also orders every is_dead clear above before the new flags)
    const unsigned long long kstar = cluster_select_kworst(cluster, sm, keys, N, k, dead_list, is_dead);

    const double logw = logX0 - log(n_global + 1.0);
    const double w = exp(logw);
    const bool room = base_row + k <= capacity;
    if (defer_book && n_dead0 + k < n_target) {
        // the walkers only need dead_list / is_dead / L* / the batch header: publish those and leave
        // the dead-row copy and the evidence sums to the first walker launch (walk_reg_kernel, phase 0),
        // where they run beside the walkers instead of in front of them
        if (c == 0 && tid == 0) {
            S->book_logw = logw;
            S->book_w = w;
            S->book_base_row = base_row;
            S->book_step = step_now;
            S->book_room = room ? 1 : 0;
            S->logX = logX0 + log((n_global - (double)k + 1.0) / (n_global + 1.0));
            S->n_dead = n_dead0 + k;
            if (room) S->n_rows = base_row + k;
            S->lstar = key_to_double(kstar);
            S->k_cur = k;
            S->do_walk = 1;
            if (end_walk_tc >= 0) {
                S->n_evals += (unsigned long long)((long long)kp * evals_per_walker);
                S->step = step_now;
            }
        }
        cluster.sync();
        return;
    }

    // ---- record: rows spread over the cluster --------------------------------------------------------
    if (room) {
        for (int e = c * KC_THREADS + tid; e < k * d; e += KC_CTAS * KC_THREADS) {
            const int q = e / d, j = e - q * d;
            const long long slot = dead_list[q];
            dead_x[(base_row + q) * d + j] = live[slot * d + j];
            if (j == 0) {
                dead_logl[base_row + q] = logl[slot];
                dead_w[base_row + q] = w;
                dead_idx[base_row + q] = slot + row_offset;
                dead_step[base_row + q] = step_now;
            }
        }
    }

    // ---- accumulate on CTA 0 ----------------------------------------------------------------------------
    if (c == 0) {
        double mx, m_new, sums[4];
        cta_accumulate_v2(S, dead_list, logl, k, logw, w, m_old, scratch, mx, m_new, sums);
        if (tid == 0) {
            if (mx > -INFINITY) {
                const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
                S->Zs = S->Zs * scale + sums[0];
                S->Hs = S->Hs * scale + sums[1];
                S->M = m_new;
            }
            S->Z_lin += sums[2];
            S->H_lin += sums[3];
            S->logX = logX0 + log((n_global - (double)k + 1.0) / (n_global + 1.0));
            S->n_dead = n_dead0 + k;
            if (room) S->n_rows = base_row + k;
            S->lstar = key_to_double(kstar);
            S->k_cur = k;
            S->do_walk = 1;   // (the host decides whether a refill follows: see accumulate_kernel)
            if (end_walk_tc >= 0) {
                S->n_evals += (unsigned long long)((long long)kp * evals_per_walker);
                S->step = step_now;
            }
        }
    }
    cluster.sync();                                              // nobody exits while its mailbox may still be read
}

// Rank boundaries of the ascending `positions[0..k)` (gathered index = source rank * m + i):
// bound[r] = first index whose position belongs to rank >= r, r = 0..world.  One parallel pass with
// one global round trip (a binary search per boundary is ~22 dependent L2 loads, ~7 us).
__device__ __forceinline__ void shard_rank_bounds(const long long *__restrict__ positions, int k, int m, int world,
                                                  int *bound /* shared, world+1 */) {
    const int tid = threadIdx.x;
    if (tid <= world) bound[tid] = k;
    __syncthreads();
    for (int i = tid; i < k; i += blockDim.x) {
        const int r = (int)((unsigned int)positions[i] / (unsigned int)m);
        const int rprev = i > 0 ? (int)((unsigned int)positions[i - 1] / (unsigned int)m) : -1;
        for (int q = rprev + 1; q <= r; ++q) bound[q] = i;
    }
    __syncthreads();
}

// ---- sharded step on a cluster: candidates and commit ---------------------------------------------
__global__ void __cluster_dims__(KC_CTAS, 1, 1) __launch_bounds__(KC_THREADS) shard_candidates_cluster_kernel(
    EngineState *S, int end_walk_tc /* >= 0: first finish the previous refill */, long long evals_per_walker, int d,
    const long long *__restrict__ dead_list, unsigned char *__restrict__ is_dead, const double *__restrict__ logl, int N,
    int m, long long *__restrict__ cand_list, double *__restrict__ cand_logl) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ unsigned long long keys[];
    __shared__ ClusterSelSmem sm;
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int kp = S->k_cur;
    if (end_walk_tc >= 0) {                                      // end_walk_kernel's job, spread over the cluster
        for (int i = c * KC_THREADS + tid; i < kp; i += KC_CTAS * KC_THREADS) is_dead[dead_list[i]] = 0;
        if (c == 0) {
            const double wnew = tid < d ? phase_width(S, tid, end_walk_tc, kp) : 0.0;
            __syncthreads();
            if (tid < d) {
                S->widths[tid] = wnew;
                for (int q = 0; q < end_walk_tc; ++q) S->acc[q][tid] = 0u;
            }
        }
    }
    const int Nc = (N + KC_CTAS - 1) / KC_CTAS;
    const int lo_g = c * Nc;
    const int n_loc = N - lo_g < 0 ? 0 : (N - lo_g < Nc ? N - lo_g : Nc);
    for (int i = tid; i < n_loc; i += KC_THREADS) keys[i] = sort_key(logl[lo_g + i]);
    __syncthreads();
    (void)cluster_select_kworst(cluster, sm, keys, N, m, cand_list, nullptr);
    for (int i = c * KC_THREADS + tid; i < m; i += KC_CTAS * KC_THREADS) cand_logl[i] = logl[cand_list[i]];
    if (end_walk_tc >= 0 && c == 0 && tid == 0) {                // every CTA read S->k_cur before the first barrier
        S->n_evals += (unsigned long long)((long long)kp * evals_per_walker);
        S->step += 1;
        S->do_walk = 0;
        S->k_cur = 0;
    }
    cluster.sync();
}

__global__ void __cluster_dims__(KC_CTAS, 1, 1) __launch_bounds__(KC_THREADS) shard_commit_cluster_kernel(
    EngineState *S, const double *__restrict__ gathered, int world, int rank, int k, int m, int d, long long row_offset,
    double n_global, long long capacity, const long long *__restrict__ cand_list, long long *__restrict__ positions,
    const double *__restrict__ live, const double *__restrict__ logl, long long *__restrict__ dead_list,
    unsigned char *__restrict__ is_dead, double *__restrict__ dead_x, double *__restrict__ dead_logl,
    double *__restrict__ dead_w, long long *__restrict__ dead_idx, int *__restrict__ dead_step,
    int defer_book /* record + accumulate are left to the first walker launch of the refill that follows */) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ unsigned long long keys[];
    __shared__ ClusterSelSmem sm;
    __shared__ double scratch[33];
    __shared__ int s_bound[GL_MAX_PEERS + 1];
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const long long n_dead0 = S->n_dead, base_row = S->n_rows;
    const int step_now = S->step;
    const double logX0 = S->logX, m_old = S->M;
    const int n = world * m;
    const int Nc = (n + KC_CTAS - 1) / KC_CTAS;
    const int lo_g = c * Nc;
    const int n_loc = n - lo_g < 0 ? 0 : (n - lo_g < Nc ? n - lo_g : Nc);
    for (int i = tid; i < n_loc; i += KC_THREADS) keys[i] = sort_key(gathered[lo_g + i]);
    __syncthreads();
    const unsigned long long kstar = cluster_select_kworst(cluster, sm, keys, n, k, positions, nullptr);
    shard_rank_bounds(positions, k, m, world, s_bound);          // every rank's contiguous share of `positions`
    if (c == 0 && m < k && tid < world && s_bound[tid + 1] - s_bound[tid] >= m) atomicMax(&S->shard_overflow, 1);
    const int start = s_bound[rank], n_own = s_bound[rank + 1] - s_bound[rank];
    const double logw = logX0 - log(n_global + 1.0);
    const double w = exp(logw);
    const bool room = base_row + n_own <= capacity;
    if (defer_book) {
        for (int i = c * KC_THREADS + tid; i < n_own; i += KC_CTAS * KC_THREADS) {
            const long long slot = cand_list[positions[start + i] - (long long)rank * m];
            dead_list[i] = slot;
            is_dead[slot] = 1;
        }
        if (c == 0 && tid == 0) {
            S->book_logw = logw;
            S->book_w = w;
            S->book_base_row = base_row;
            S->book_step = step_now;
            S->book_room = room ? 1 : 0;
            S->logX = logX0 + log((n_global - (double)k + 1.0) / (n_global + 1.0));
            S->n_dead = n_dead0 + k;
            if (room) S->n_rows = base_row + n_own;
            S->lstar = key_to_double(kstar);
            S->k_cur = n_own;
            S->do_walk = 1;
        }
        cluster.sync();
        return;
    }
    for (int i = c * KC_THREADS + tid; i < n_own; i += KC_CTAS * KC_THREADS) {
        const long long slot = cand_list[positions[start + i] - (long long)rank * m];
        dead_list[i] = slot;
        is_dead[slot] = 1;
        if (room) {
            dead_logl[base_row + i] = logl[slot];
            dead_w[base_row + i] = w;
            dead_idx[base_row + i] = slot + row_offset;
            dead_step[base_row + i] = step_now;
        }
    }
    if (room) {
        for (int e = c * KC_THREADS + tid; e < n_own * d; e += KC_CTAS * KC_THREADS) {
            const int q = e / d, j = e - q * d;
            const long long slot = cand_list[positions[start + q] - (long long)rank * m];
            dead_x[(base_row + q) * d + j] = live[slot * d + j];
        }
    }
    if (c == 0) {
        double mx, m_new, sums[4];
        cta_accumulate_v2(S, positions, gathered, k, logw, w, m_old, scratch, mx, m_new, sums);
        if (tid == 0) {
            if (mx > -INFINITY) {
                const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
                S->Zs = S->Zs * scale + sums[0];
                S->Hs = S->Hs * scale + sums[1];
                S->M = m_new;
            }
            S->Z_lin += sums[2];
            S->H_lin += sums[3];
            S->logX = logX0 + log((n_global - (double)k + 1.0) / (n_global + 1.0));
            S->n_dead = n_dead0 + k;
            if (room) S->n_rows = base_row + n_own;
            S->lstar = key_to_double(kstar);
            S->k_cur = n_own;
            S->do_walk = 1;
        }
    }
    cluster.sync();
}

// ---- fused sharded step over NVLink peer memory ----------------------------------------------------
// candidates -> exchange -> commit in ONE launch, no NCCL call on the step path: the cluster selects
// this rank's m worst, STORES their logL straight into every peer's exchange buffer (peer pointers
// from CUDA IPC, comm.cu), fences, raises its arrival flag on every peer (st.release.sys), waits for
// the world's flags on its own (ld.acquire.sys), and runs the global selection on what arrived.
// Buffers are double-buffered by epoch parity: a rank can be at most one step ahead of the slowest
// (its next arrival wait needs everybody's next flag), so a buffer is never overwritten while read.
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Watchdog of the peer-memory arrival wait, in ns: GLEIPNIR_PEER_TIMEOUT_S seconds (default 600 -- the wait
// includes the other ranks' walker time of the step, which is tens of seconds for ODE likelihoods).
static unsigned long long peer_timeout_ns_host() {
    static unsigned long long cached = 0;
    if (!cached) {
        const char *e = getenv("GLEIPNIR_PEER_TIMEOUT_S");
        double s = e ? atof(e) : 600.0;
        if (!(s > 0.0)) s = 600.0;
        cached = (unsigned long long)(s * 1e9);
    }
    return cached;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __cluster_dims__(KC_CTAS, 1, 1) __launch_bounds__(KC_THREADS) shard_step_p2p_kernel(
    EngineState *S, int end_walk_tc, long long evals_per_walker, int N, int m, int world, int rank, int k, int d,
    long long row_offset, double n_global, long long capacity, PeerTables T, unsigned int epoch,
    long long *__restrict__ cand_list, long long *__restrict__ positions, const double *__restrict__ live,
    const double *__restrict__ logl, long long *__restrict__ dead_list, unsigned char *__restrict__ is_dead,
    double *__restrict__ dead_x, double *__restrict__ dead_logl, double *__restrict__ dead_w,
    long long *__restrict__ dead_idx, int *__restrict__ dead_step, int defer_book,
    unsigned long long peer_timeout_ns) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ unsigned long long keys[];
    __shared__ ClusterSelSmem sm;
    __shared__ double scratch[33];
    __shared__ int s_bound[GL_MAX_PEERS + 1];
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int kp = S->k_cur;
    const int step_now = S->step + (end_walk_tc >= 0 ? 1 : 0);
    const long long n_dead0 = S->n_dead, base_row = S->n_rows;
    const double logX0 = S->logX, m_old = S->M;
    if (end_walk_tc >= 0) {                                      // end_walk_kernel's job, spread over the cluster
        for (int i = c * KC_THREADS + tid; i < kp; i += KC_CTAS * KC_THREADS) is_dead[dead_list[i]] = 0;
        if (c == 0) {
            const double wnew = tid < d ? phase_width(S, tid, end_walk_tc, kp) : 0.0;
            __syncthreads();
            if (tid < d) {
                S->widths[tid] = wnew;
                for (int q = 0; q < end_walk_tc; ++q) S->acc[q][tid] = 0u;
            }
        }
    }
    // ---- local candidates ------------------------------------------------------------------------------
    {
        const int Nc = (N + KC_CTAS - 1) / KC_CTAS;
        const int lo_g = c * Nc;
        const int n_loc = N - lo_g < 0 ? 0 : (N - lo_g < Nc ? N - lo_g : Nc);
        for (int i = tid; i < n_loc; i += KC_THREADS) keys[i] = sort_key(logl[lo_g + i]);
        __syncthreads();
        (void)cluster_select_kworst(cluster, sm, keys, N, m, cand_list, nullptr);
    }
    // ---- exchange over peer memory ---------------------------------------------------------------------
    const long long buf_off = (long long)(epoch & 1u) * (long long)world * T.cap;
    for (int i = c * KC_THREADS + tid; i < m; i += KC_CTAS * KC_THREADS) {
        const double v = logl[cand_list[i]];
        for (int p = 0; p < world; ++p) T.gather[p][buf_off + (long long)rank * m + i] = v;
    }
    __threadfence_system();
    cluster.sync();
    if (c == 0 && tid < world) {
        st_release_sys(T.flags[tid] + rank, epoch);               // "rank's candidates of this epoch are in your buffer"
        const unsigned int *mine = T.flags[rank] + tid;
        // wall-clock watchdog (%globaltimer, ns): the wait also absorbs the ranks' walker skew, which is
        // seconds for ODE likelihoods, so the limit is generous and configurable (GLEIPNIR_PEER_TIMEOUT_S)
        const unsigned long long t0 = global_timer_ns();
        int timeout = 0;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {         // flags only grow (wrap-safe compare)
            __nanosleep(40);
            if (global_timer_ns() - t0 > peer_timeout_ns) { timeout = 1; break; }    // a peer died
        }
        if (timeout) atomicMax(&S->shard_overflow, 2);            // sticky and never downgraded to 1
    }
    cluster.sync();
    // ---- global selection on what arrived ------------------------------------------------------------------
    const double *gathered = T.gather[rank] + buf_off;
    const int n = world * m;
    {
        const int Nc = (n + KC_CTAS - 1) / KC_CTAS;
        const int lo_g = c * Nc;
        const int n_loc = n - lo_g < 0 ? 0 : (n - lo_g < Nc ? n - lo_g : Nc);
        for (int i = tid; i < n_loc; i += KC_THREADS) keys[i] = sort_key(__ldcg(gathered + lo_g + i));
        __syncthreads();
    }
    const unsigned long long kstar = cluster_select_kworst(cluster, sm, keys, n, k, positions, nullptr);
    shard_rank_bounds(positions, k, m, world, s_bound);          // every rank's contiguous share of `positions`
    if (c == 0 && m < k && tid < world && s_bound[tid + 1] - s_bound[tid] >= m) atomicMax(&S->shard_overflow, 1);
    const int start = s_bound[rank], n_own = s_bound[rank + 1] - s_bound[rank];
    const double logw = logX0 - log(n_global + 1.0);
    const double w = exp(logw);
    const bool room = base_row + n_own <= capacity;
    for (int i = c * KC_THREADS + tid; i < n_own; i += KC_CTAS * KC_THREADS) {
        const long long slot = cand_list[positions[start + i] - (long long)rank * m];
        dead_list[i] = slot;
        is_dead[slot] = 1;
        if (!defer_book && room) {
            dead_logl[base_row + i] = logl[slot];
            dead_w[base_row + i] = w;
            dead_idx[base_row + i] = slot + row_offset;
            dead_step[base_row + i] = step_now;
        }
    }
    if (!defer_book && room) {
        for (int e = c * KC_THREADS + tid; e < n_own * d; e += KC_CTAS * KC_THREADS) {
            const int q = e / d, j = e - q * d;
            const long long slot = cand_list[positions[start + q] - (long long)rank * m];
            dead_x[(base_row + q) * d + j] = live[slot * d + j];
        }
    }
    if (c == 0) {
        double mx = -INFINITY, m_new = m_old, sums[4] = {0.0, 0.0, 0.0, 0.0};
        if (!defer_book) cta_accumulate_v2(S, positions, gathered, k, logw, w, m_old, scratch, mx, m_new, sums);
        if (tid == 0) {
            if (defer_book) {
                S->book_logw = logw;
                S->book_w = w;
                S->book_base_row = base_row;
                S->book_step = step_now;
                S->book_room = room ? 1 : 0;
            } else {
                if (mx > -INFINITY) {
                    const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
                    S->Zs = S->Zs * scale + sums[0];
                    S->Hs = S->Hs * scale + sums[1];
                    S->M = m_new;
                }
                S->Z_lin += sums[2];
                S->H_lin += sums[3];
            }
            S->logX = logX0 + log((n_global - (double)k + 1.0) / (n_global + 1.0));
            S->n_dead = n_dead0 + k;
            if (room) S->n_rows = base_row + n_own;
            S->lstar = key_to_double(kstar);
            S->k_cur = n_own;
            S->do_walk = 1;
            if (end_walk_tc >= 0) {
                S->n_evals += (unsigned long long)((long long)kp * evals_per_walker);
                S->step = step_now;
            }
        }
    }
    cluster.sync();
}

// ---- end of refill: clear the dead flags, persist widths, advance the step ----------------
__global__ void __launch_bounds__(256) end_walk_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                       unsigned char *is_dead, int d, int tuning_cycles,
                                                       long long evals_per_walker) {
    const int k = S->k_cur;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
        is_dead[dead_list[i]] = 0;
    if (blockIdx.x == 0) {
        // GL_MAX_DIM <= blockDim: thread j owns component j
        const int j = threadIdx.x;
        const double wnew = j < d ? phase_width(S, j, tuning_cycles, k) : 0.0;
        __syncthreads();          // every width derived before any acc/width is overwritten
        if (j < d) {
            S->widths[j] = wnew;
            for (int c = 0; c < tuning_cycles; ++c) S->acc[c][j] = 0u;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S->n_evals += (unsigned long long)((long long)k * evals_per_walker);
        S->step += 1;
        S->do_walk = 0;
        S->k_cur = 0;
    }
}

// ---- inject: the stub-sampler path of the replay tests ----------------------------------
// Replacement rows arrive in DEATH order (ascending (logL, index)); dead_list is in index
// order, so each slot's rank in death order is computed on the fly (k is small in tests).
__global__ void __launch_bounds__(256) inject_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                     const double *__restrict__ old_logl_by_list,
                                                     const double *__restrict__ pts, const double *__restrict__ lls,
                                                     int d, double *__restrict__ live, double *__restrict__ logl) {
    const int k = S->k_cur;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        const unsigned long long ki = sort_key(old_logl_by_list[i]);
        int rank = 0;
        for (int q = 0; q < k; ++q) {
            const unsigned long long kq = sort_key(old_logl_by_list[q]);
            rank += (kq < ki) || (kq == ki && q < i);
        }
        const long long slot = dead_list[i];
        for (int j = 0; j < d; ++j) live[slot * d + j] = pts[(size_t)rank * d + j];
        logl[slot] = lls[rank];
    }
}
__global__ void __launch_bounds__(256) gather_logl_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                          const double *__restrict__ logl, double *__restrict__ out) {
    const int k = S->k_cur;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
        out[i] = logl[dead_list[i]];
}

// ---- initial proposal widths (samplers.py:70-81): 0.5 * range of 100 prior draws ---------
__global__ void __launch_bounds__(128) init_widths_kernel(EngineState *S, const PriorTable *pt, int d, uint64_t seed) {
    __shared__ double smin[4], smax[4];
    const int j = blockIdx.x;
    if (j >= d) return;
    double lo = INFINITY, hi = -INFINITY;
    if (threadIdx.x < 100) {
        const U2 u = philox_uniforms(seed, (uint32_t)threadIdx.x, 0u, (uint32_t)j, GL_TAG_WIDTH);
        const double v = prior_ppf(pt->kind[j], pt->p0[j], pt->p1[j], u.a);
        lo = v;
        hi = v;
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w) { lo = fmin(lo, smin[w]); hi = fmax(hi, smax[w]); }
        S->widths[j] = 0.5 * (hi - lo);
    }
}

// ---- finalize (nested_sampling.py:183-203) ----------------------------------------------
struct LsePair {
    double m, s;
};
__device__ __forceinline__ LsePair lse_combine(LsePair a, LsePair b) {
    if (a.m == -INFINITY) return b;
    if (b.m == -INFINITY) return a;
    LsePair r;
    if (a.m >= b.m) { r.m = a.m; r.s = a.s + b.s * exp(b.m - a.m); }
    else { r.m = b.m; r.s = b.s + a.s * exp(a.m - b.m); }
    return r;
}

constexpr int FIN_THREADS = 256;

// partials[b] = (m, s, linear sum of exp(L), count) over the survivors of block b's chunk
__global__ void __launch_bounds__(FIN_THREADS) final_partial_kernel(const double *__restrict__ logl,
                                                                    const unsigned char *__restrict__ is_dead,
                                                                    long long N, long long chunk, double *partials) {
    __shared__ double sm_m[FIN_THREADS], sm_s[FIN_THREADS], sm_l[FIN_THREADS], sm_c[FIN_THREADS];
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > N) hi = N;
    LsePair p{-INFINITY, 0.0};
    double lin = 0.0, cnt = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += FIN_THREADS) {
        if (is_dead[i]) continue;
        const double L = logl[i];
        cnt += 1.0;
        lin += exp(L);
        if (isfinite(L)) p = lse_combine(p, LsePair{L, 1.0});
    }
    sm_m[threadIdx.x] = p.m; sm_s[threadIdx.x] = p.s; sm_l[threadIdx.x] = lin; sm_c[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            LsePair a{sm_m[threadIdx.x], sm_s[threadIdx.x]}, b{sm_m[threadIdx.x + o], sm_s[threadIdx.x + o]};
            a = lse_combine(a, b);
            sm_m[threadIdx.x] = a.m; sm_s[threadIdx.x] = a.s;
            sm_l[threadIdx.x] += sm_l[threadIdx.x + o];
            sm_c[threadIdx.x] += sm_c[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[4 * blockIdx.x + 0] = sm_m[0];
        partials[4 * blockIdx.x + 1] = sm_s[0];
        partials[4 * blockIdx.x + 2] = sm_l[0];
        partials[4 * blockIdx.x + 3] = sm_c[0];
    }
}

// partials [local_first, local_first + local_count) belong to this engine (all of them when unsharded)
__global__ void final_combine_kernel(EngineState *S, const double *partials, int nblocks, int local_first,
                                     int local_count) {
    LsePair p{-INFINITY, 0.0};
    double lin = 0.0, cnt = 0.0, cnt_local = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        p = lse_combine(p, LsePair{partials[4 * b], partials[4 * b + 1]});
        lin += partials[4 * b + 2];
        cnt += partials[4 * b + 3];
        if (b >= local_first && b < local_first + local_count) cnt_local += partials[4 * b + 3];
    }
    S->fin_n_local = (long long)cnt_local;
    S->n_rows_dead = S->n_rows;
    if (cnt <= 0.0) return;
    const double X = exp(S->logX);
    const double lm = lin / cnt;                                  // mean likelihood (:187)
    S->Z_lin += X * lm;
    const double dH = X * lm * log(lm);
    if (dH == dH) S->H_lin += dH;
    if (p.m > -INFINITY) {
        const double loglm = p.m + log(p.s) - log(cnt);
        const double lw = S->logX + loglm;
        const double m_old = S->M, m_new = fmax(m_old, lw);
        const double scale = isfinite(m_old) ? exp(m_old - m_new) : 0.0;
        const double t = exp(lw - m_new);
        S->Zs = S->Zs * scale + t;
        S->Hs = S->Hs * scale + t * loglm;
        S->M = m_new;
    }
    S->fin_n = (long long)cnt;
}

// survivors appended in index order with weight X/(#survivors) (:195-203); a survivor's row
// position is its index minus the number of dead slots before it (dead_list is ascending).
__global__ void __launch_bounds__(256) final_rows_kernel(EngineState *S, const long long *__restrict__ dead_list,
                                                         const unsigned char *__restrict__ is_dead,
                                                         const double *__restrict__ live,
                                                         const double *__restrict__ logl, long long N, int d,
                                                         long long row_offset, long long capacity,
                                                         double *__restrict__ dead_x, double *__restrict__ dead_logl,
                                                         double *__restrict__ dead_w, long long *__restrict__ dead_idx,
                                                         int *__restrict__ dead_step) {
    const int k = S->k_cur;                      // unreplaced slots of the last batch
    const long long base = S->n_rows;
    const long long n_surv = S->fin_n;
    if (n_surv <= 0 || base + S->fin_n_local > capacity) return;
    const double aw = exp(S->logX) / (double)n_surv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x) {
        if (is_dead[i]) continue;
        int lo = 0, hi = k;                      // lower_bound(dead_list, i)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (dead_list[mid] < i) lo = mid + 1;
            else hi = mid;
        }
        const long long pos = base + (i - lo);
        for (int j = 0; j < d; ++j) dead_x[pos * d + j] = live[i * d + j];
        dead_logl[pos] = logl[i];
        dead_w[pos] = aw;
        dead_idx[pos] = i + row_offset;
        dead_step[pos] = -1;
    }
}

__global__ void final_done_kernel(EngineState *S, long long capacity) {
    if (S->fin_n > 0 && S->n_rows + S->fin_n_local <= capacity) S->n_rows += S->fin_n_local;
    S->finalized = 1;
}

constexpr int FIN_MAX_BLOCKS = 1024;

// ---- export: dead table [rows, d+2] in death order ---------------------------------------------
// The log holds each batch in slot order; the reference's table is in death order (ascending
// (logL, slot) inside a batch, survivors last in slot order; nested_sampling.py:130-133,195-203).
// One CTA per batch sorts its (key, row) pairs with a bitonic network in shared memory and writes
// the rows straight into the output table, so the export is one kernel + one D2H copy.
constexpr int EXPORT_THREADS = 512;

__global__ void __launch_bounds__(EXPORT_THREADS) export_batches_kernel(
    const EngineState *S, const double *__restrict__ dead_x, const double *__restrict__ dead_logl,
    const double *__restrict__ dead_w, const long long *__restrict__ dead_idx, const int *__restrict__ dead_step,
    int d, int seg_cap /* power of two >= largest batch */, double *__restrict__ out) {
    extern __shared__ unsigned long long sk[];                 // keys[seg_cap] | rows[seg_cap] (as u64)
    unsigned long long *sr = sk + seg_cap;
    const long long n_dead = S->finalized ? S->n_rows_dead : S->n_rows;
    const long long n_rows = S->n_rows;
    const int step = blockIdx.x;
    // rows of this batch: [lo, hi) by binary search on the non-decreasing step column
    long long lo = 0, hi = n_dead;
    {
        long long a = 0, b = n_dead;
        while (a < b) { const long long m = (a + b) >> 1; if (dead_step[m] < step) a = m + 1; else b = m; }
        lo = a;
        b = n_dead;
        while (a < b) { const long long m = (a + b) >> 1; if (dead_step[m] < step + 1) a = m + 1; else b = m; }
        hi = a;
    }
    const int n = (int)(hi - lo);
    if (n > 0 && n <= seg_cap) {
        for (int i = threadIdx.x; i < seg_cap; i += EXPORT_THREADS) {
            sk[i] = i < n ? sort_key(dead_logl[lo + i]) : ~0ull;
            sr[i] = i < n ? (unsigned long long)dead_idx[lo + i] << 24 | (unsigned long long)i : ~0ull;
        }
        __syncthreads();
        // bitonic sort on (key, slot) -- rows inside a batch are already in slot order, and the
        // slot is carried in sr's high bits, so ties break by lowest slot
        for (int size = 2; size <= seg_cap; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = threadIdx.x; t < seg_cap / 2; t += EXPORT_THREADS) {
                    const int i = 2 * t - (t & (stride - 1));
                    const int j = i + stride;
                    const bool up = (i & size) == 0;
                    const unsigned long long ki = sk[i], kj = sk[j], ri = sr[i], rj = sr[j];
                    const bool gt = ki > kj || (ki == kj && ri > rj);
                    if (gt == up) { sk[i] = kj; sk[j] = ki; sr[i] = rj; sr[j] = ri; }
                }
                __syncthreads();
            }
        }
        for (long long e = threadIdx.x; e < (long long)n * (d + 2); e += EXPORT_THREADS) {
            const int q = (int)(e / (d + 2)), c = (int)(e - (long long)q * (d + 2));
            const long long src = lo + (long long)(sr[q] & 0xffffffull);
            out[(lo + q) * (d + 2) + c] = c == 0 ? dead_logl[src] : c == 1 ? dead_w[src] : dead_x[src * d + (c - 2)];
        }
    } else if (n > seg_cap) {
        // oversized batch: keep slot order (the host sorts such batches)
        for (long long e = threadIdx.x; e < (long long)n * (d + 2); e += EXPORT_THREADS) {
            const long long q = e / (d + 2);
            const int c = (int)(e - q * (d + 2));
            const long long src = lo + q;
            out[src * (d + 2) + c] = c == 0 ? dead_logl[src] : c == 1 ? dead_w[src] : dead_x[src * d + (c - 2)];
        }
    }
    // the last block also copies the survivor rows (already in slot order)
    if (blockIdx.x == gridDim.x - 1) {
        for (long long e = threadIdx.x; e < (n_rows - n_dead) * (d + 2); e += EXPORT_THREADS) {
            const long long q = e / (d + 2);
            const int c = (int)(e - q * (d + 2));
            const long long src = n_dead + q;
            out[src * (d + 2) + c] = c == 0 ? dead_logl[src] : c == 1 ? dead_w[src] : dead_x[src * d + (c - 2)];
        }
    }
}

// Sharded run: every rank's log, all-gathered as packed rows [world][mx][d+4] = (logL, weight, x[d],
// global slot, step) with counts[r] valid rows in rank r's block (dead rows in step order, then the
// survivors with step -1), assembled into ONE table in death order: CTA s merges the segments of
// step s from all ranks and sorts them by (key, slot); CTAs n_steps.. copy the survivors rank by
// rank (= global slot order).  Every batch but the last holds exactly batch_k rows.
__device__ __forceinline__ long long packed_step_lower_bound(const double *__restrict__ rows, long long n, int w,
                                                             int d, int step) {
    long long a = 0, b = n;                        // first row whose step (survivors: +inf) is >= `step`
    while (a < b) {
        const long long m = (a + b) >> 1;
        const int sm = (int)rows[m * w + d + 3];
        if ((sm < 0 ? 0x7fffffff : sm) < step) a = m + 1; else b = m;
    }
    return a;
}

__global__ void __launch_bounds__(EXPORT_THREADS) export_gathered_kernel(
    const double *__restrict__ packed, const long long *__restrict__ counts, int world, long long mx, int d, int batch_k,
    int n_steps, long long n_dead_total, int seg_cap, double *__restrict__ out) {
    extern __shared__ unsigned long long sk[];                 // keys[seg_cap] | tags[seg_cap] | src[seg_cap] (u32)
    unsigned long long *sr = sk + seg_cap;
    unsigned int *ssrc = reinterpret_cast<unsigned int *>(sr + seg_cap);
    __shared__ long long s_lo[64], s_n[64];
    const int w = d + 4;
    if ((int)blockIdx.x < n_steps) {
        const int step = blockIdx.x;
        if (threadIdx.x < world) {
            const double *rows = packed + (long long)threadIdx.x * mx * w;
            const long long lo = packed_step_lower_bound(rows, counts[threadIdx.x], w, d, step);
            const long long hi = packed_step_lower_bound(rows, counts[threadIdx.x], w, d, step + 1);
            s_lo[threadIdx.x] = lo;
            s_n[threadIdx.x] = hi - lo;
        }
        __syncthreads();
        int n = 0;
        for (int r = 0; r < world; ++r) n += (int)s_n[r];
        if (n <= 0 || n > seg_cap) return;          // (host checks batch_k <= seg_cap)
        for (int i = threadIdx.x; i < seg_cap; i += EXPORT_THREADS) { sk[i] = ~0ull; sr[i] = ~0ull; }
        __syncthreads();
        int off = 0;
        for (int r = 0; r < world; ++r) {
            const int nr = (int)s_n[r];
            const long long row0 = (long long)r * mx + s_lo[r];
            for (int i = threadIdx.x; i < nr; i += EXPORT_THREADS) {
                const double *row = packed + (row0 + i) * w;
                sk[off + i] = sort_key(row[0]);
                sr[off + i] = (unsigned long long)(long long)row[d + 2] << 24 | (unsigned long long)(off + i);
                ssrc[off + i] = (unsigned int)(row0 + i);
            }
            off += nr;
        }
        __syncthreads();
        for (int size = 2; size <= seg_cap; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = threadIdx.x; t < seg_cap / 2; t += EXPORT_THREADS) {
                    const int i = 2 * t - (t & (stride - 1));
                    const int j = i + stride;
                    const bool up = (i & size) == 0;
                    const unsigned long long ki = sk[i], kj = sk[j], ri = sr[i], rj = sr[j];
                    const bool gt = ki > kj || (ki == kj && ri > rj);
                    if (gt == up) { sk[i] = kj; sk[j] = ki; sr[i] = rj; sr[j] = ri; }
                }
                __syncthreads();
            }
        }
        const long long out0 = (long long)step * batch_k;
        for (long long e = threadIdx.x; e < (long long)n * (d + 2); e += EXPORT_THREADS) {
            const int q = (int)(e / (d + 2)), c = (int)(e - (long long)q * (d + 2));
            const long long src = ssrc[sr[q] & 0xffffffull];
            out[(out0 + q) * (d + 2) + c] = packed[src * w + c];
        }
    } else {
        const int r = (int)blockIdx.x - n_steps;    // survivors of rank r
        long long out0 = n_dead_total;
        for (int q = 0; q < r; ++q) {
            const long long nd = packed_step_lower_bound(packed + (long long)q * mx * w, counts[q], w, d, 0x7fffffff);
            out0 += counts[q] - nd;
        }
        const long long nd = packed_step_lower_bound(packed + (long long)r * mx * w, counts[r], w, d, 0x7fffffff);
        const long long ns = counts[r] - nd;
        for (long long e = threadIdx.x; e < ns * (d + 2); e += EXPORT_THREADS) {
            const long long q = e / (d + 2);
            const int c = (int)(e - q * (d + 2));
            out[(out0 + q) * (d + 2) + c] = packed[((long long)r * mx + nd + q) * w + c];
        }
    }
}

}  // namespace gl

// ============================== host side / C ABI ==========================================
struct gl_ns {
    cudaGraphExec_t graph_exec = nullptr;         // gl_ns_advance_graph: kept until the next capture / destroy
    gl_ns_config cfg;
    int d;
    gl::ModelDev model;
    const gl_model *model_handle;
    gl_prior priors[GL_MAX_DIM];
    gl_ns_buffers buf;
    // carved out of buf.workspace
    gl::EngineState *S;
    gl::SelectState *sel;
    gl::PriorTable *pt;
    long long *dead_list;
    unsigned char *is_dead;
    double *scratch_x, *scratch_l, *tmp_k, *fin_partials;
    long long *cand_list, *positions;             // sharded run scratch ([k] each)
    unsigned long long *sk0, *sk1;                // sorted live set: keys (two buffers)
    unsigned int *si0, *si1;                      //                  slots
    int select_path;                              // 0 auto, 1 multipass, 2 fused radix, 3 fused sorted
    bool h_sorted_ready;
    gl::SelectState *sel2;                        // selection state for the gathered candidates
    // host mirrors (deterministic arithmetic; no device reads needed to drive the loop)
    long long h_n_dead, h_n_evals, h_launches;
    int h_step, h_k_last;
    bool initialized, pending_refill, finalized;
    // optional CUDA-event timing of the walker / select launches (bench roofline)
    bool prof;
    bool force_multipass;                         // tests: exercise the multi-launch select path
    bool end_walk_deferred;                       // last refill's epilogue is still to run
    bool defer_ok;                                // inside gl_ns_advance(_sharded): the next op is a kill / candidates
    bool sharded_loop;                            // inside gl_ns_advance_sharded
    bool commit_refills;                          // ... and the next commit is followed by a refill
    bool force_smem_walker;                       // tests: generic shared-memory walker even for small d
    bool book_deferred;                           // the last kill / commit left record + accumulate to the next refill
    int walk_k_max;                               // upper bound of the batch the next refill serves (0: batch_k)
    int book_k;                                   // values of that batch: L(i) = book_vals[book_idx[i]]
    const double *book_vals;
    const long long *book_idx;
    std::vector<cudaEvent_t> ev_walk, ev_sel;     // (start, stop) pairs on the launching stream
    long long prof_walk_launches, prof_sel_launches;
};

namespace gl {

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct WorkspaceLayout {
    size_t off_state, off_sel, off_pt, off_list, off_isdead, off_sx, off_sl, off_tmp, off_fin, off_cand, off_pos,
        off_sel2, off_sk0, off_sk1, off_si0, off_si1, total;
};

static WorkspaceLayout workspace_layout(const gl_ns_config &c, int d) {
    WorkspaceLayout L;
    size_t o = 0;
    const size_t k = (size_t)(c.batch_k > 0 ? c.batch_k : 1);
    L.off_state = o;  o += align256(sizeof(EngineState));
    L.off_sel = o;    o += align256(sizeof(SelectState));
    L.off_pt = o;     o += align256(sizeof(PriorTable));
    L.off_list = o;   o += align256(sizeof(long long) * k);
    L.off_isdead = o; o += align256((size_t)c.population_size);
    L.off_sx = o;     o += align256(sizeof(double) * k * (size_t)d);
    L.off_sl = o;     o += align256(sizeof(double) * k);
    L.off_tmp = o;    o += align256(sizeof(double) * k);
    L.off_fin = o;    o += align256(sizeof(double) * 4 * FIN_MAX_BLOCKS);
    L.off_cand = o;   o += align256(sizeof(long long) * k);
    L.off_pos = o;    o += align256(sizeof(long long) * k);
    L.off_sel2 = o;   o += align256(sizeof(SelectState));
    const size_t ns_ = (size_t)c.population_size <= (size_t)SORTED_MAX_N ? (size_t)c.population_size : 0;
    L.off_sk0 = o;    o += align256(sizeof(unsigned long long) * ns_);
    L.off_sk1 = o;    o += align256(sizeof(unsigned long long) * ns_);
    L.off_si0 = o;    o += align256(sizeof(unsigned int) * ns_);
    L.off_si1 = o;    o += align256(sizeof(unsigned int) * ns_);
    L.total = o;
    return L;
}

static int check_config(const gl_ns_config *cfg, int d) {
    GL_REQUIRE(cfg != nullptr, "cfg is NULL");
    GL_REQUIRE(d >= 1 && d <= GL_MAX_DIM, "d=%d outside [1, %d]", d, GL_MAX_DIM);
    GL_REQUIRE(cfg->population_size >= 2, "population_size must be >= 2");
    GL_REQUIRE(cfg->batch_k >= 1 && cfg->batch_k < cfg->population_size, "batch_k=%d outside [1, N)", cfg->batch_k);
    GL_REQUIRE(cfg->iterations >= 0 && cfg->burn_in >= 0, "iterations/burn_in must be >= 0");
    GL_REQUIRE(cfg->tuning_cycles >= 0 && cfg->tuning_cycles <= GL_MAX_TUNING_CYCLES, "tuning_cycles outside [0, %d]",
               GL_MAX_TUNING_CYCLES);
    GL_REQUIRE(cfg->proposal == GL_PROPOSAL_UNIFORM || cfg->proposal == GL_PROPOSAL_NORMAL, "unknown proposal");
    GL_REQUIRE(cfg->dead_capacity >= 1, "dead_capacity must be >= 1");
    GL_REQUIRE(cfg->walker_lanes >= 0 && cfg->walker_lanes <= 32, "walker_lanes outside [0, 32]");
    GL_REQUIRE(cfg->row_offset >= 0 && cfg->global_population >= cfg->population_size + cfg->row_offset,
               "shard [row_offset, row_offset + population_size) must lie inside the global population");
    return GL_OK;
}

static WalkParams walk_params(const gl_ns *ns) {
    WalkParams P;
    P.S = ns->S;
    P.dead_list = ns->dead_list;
    P.is_dead = ns->is_dead;
    P.live = ns->buf.live;
    P.logl = ns->buf.logl;
    P.priors = ns->pt;
    P.scratch_x = ns->scratch_x;
    P.scratch_l = ns->scratch_l;
    P.N = ns->cfg.population_size;
    P.row_offset = ns->cfg.row_offset;
    P.seed = ns->cfg.seed;
    P.d = ns->d;
    P.iterations = ns->cfg.iterations;
    P.burn_in = ns->cfg.burn_in;
    P.tuning_cycles = ns->cfg.tuning_cycles;
    P.proposal = ns->cfg.proposal;
    // sharded: a rank never refills more slots than the candidates it sent (m <= batch_k), so the walker
    // grid and the lanes per walker are sized for m, not for the global batch
    P.k_max = ns->walk_k_max > 0 && ns->walk_k_max < ns->cfg.batch_k ? ns->walk_k_max : ns->cfg.batch_k;
    P.lanes = ns->cfg.walker_lanes;
    P.force_smem_state = ns->force_smem_walker ? 1 : 0;
    P.book = ns->book_deferred ? ns->book_k : 0;
    P.book_vals = ns->book_vals;
    P.book_idx = ns->book_idx;
    P.n_global = (double)ns->cfg.global_population;
    P.dead_x = ns->buf.dead_x;
    P.dead_logl = ns->buf.dead_logl;
    P.dead_w = ns->buf.dead_w;
    P.dead_idx = (long long *)ns->buf.dead_idx;
    P.dead_step = ns->buf.dead_step;
    return P;
}

static int blocks_for(long long items, int threads, int cap_per_sm) {
    long long b = (items + threads - 1) / threads;
    const long long cap = (long long)sm_count() * cap_per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// opt a kernel into > 48 KB of dynamic shared memory (idempotent; raises the limit to the device max once)
static int ensure_smem(const void *fn, int bytes) {
    if (bytes <= 40 * 1024) return GL_OK;
    GL_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    return GL_OK;
}

static void prof_mark(gl_ns *ns, std::vector<cudaEvent_t> &v, cudaStream_t st) {
    if (!ns->prof || v.size() >= 16384) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    v.push_back(e);
}

static int end_walk_launch(gl_ns *ns, cudaStream_t st, long long evals_per_walker);

// a refill whose epilogue was left for the next fused kill: run it now (anything else is next)
static int flush_end_walk(gl_ns *ns, cudaStream_t st) {
    if (!ns->end_walk_deferred) return GL_OK;
    ns->end_walk_deferred = false;
    ns->h_step -= 1;                       // end_walk_launch counts the step again
    return end_walk_launch(ns, st, (20LL * ns->cfg.tuning_cycles + ns->cfg.iterations + ns->cfg.burn_in) * ns->d);
}

// Which kill the next gl_ns_kill uses: 1 multi-launch, 2 fused single-CTA radix, 3 fused sorted, 4 cluster.
// auto = cluster: at N = 10^4, k = 10^3 a kill measures 24.5 us on the 8-CTA cluster against 36 us for
// the single-CTA radix kernel and 52 us for the sorted-merge variant (a 1024-entry bitonic sort alone
// is ~43 k warp-instructions on the one SM that runs it); the cluster also serves N up to 196 608.
static int kill_variant(const gl_ns *ns) {
    const long long N = ns->cfg.population_size;
    const int path = ns->force_multipass ? 1 : ns->select_path;
    if (path == 3 && N <= SORTED_MAX_N) return 3;
    if ((path == 0 || path == 4) && N <= KC_MAX_N) return 4;
    if (path != 1 && N <= KILL_SMALL_MAX_N) return 2;
    return 1;
}
// the op that follows a refill inside the native loops folds end_walk_kernel into its prologue
static bool next_op_folds_end_walk(const gl_ns *ns) {
    if (ns->sharded_loop) return ns->cfg.population_size <= KC_MAX_N && !ns->force_multipass;
    return kill_variant(ns) != 1;
}

static int kill_launch(gl_ns *ns, int k_req, long long n_target, cudaStream_t st) {
    const long long N = ns->cfg.population_size;
    const double n_global = (double)ns->cfg.global_population;
    const int variant = kill_variant(ns);
    const bool use_sorted = variant == 3, use_cluster = variant == 4, use_radix = variant == 2;
    if (use_sorted || use_radix || use_cluster) {
        const int tc_arg = ns->end_walk_deferred ? ns->cfg.tuning_cycles : -1;
        ns->end_walk_deferred = false;
        const long long epw = (20LL * ns->cfg.tuning_cycles + ns->cfg.iterations + ns->cfg.burn_in) * ns->d;
        prof_mark(ns, ns->ev_sel, st);
        if (use_sorted) {
            // bitonic scratch: 12 B per new entry, padded to a power of two (all N on the first call)
            int nb = ns->h_sorted_ready ? ns->cfg.batch_k : (int)N, n2 = KILL_SMALL_THREADS;
            while (n2 < nb) n2 <<= 1;
            const int smem = 12 * n2 + (int)N + 16;
            int rc0 = ensure_smem((const void *)kill_small_kernel<true>, smem);
            if (rc0) return rc0;
            kill_small_kernel<true><<<1, KILL_SMALL_THREADS, smem, st>>>(
                ns->S, tc_arg, epw, k_req, n_target, (int)N, ns->d, ns->cfg.row_offset, n_global,
                ns->cfg.dead_capacity, ns->buf.live, ns->buf.logl, ns->dead_list, ns->is_dead, ns->buf.dead_x,
                ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx, ns->buf.dead_step, ns->sk0, ns->sk1,
                ns->si0, ns->si1, n2);
            ns->h_sorted_ready = true;
        } else if (use_cluster) {
            // inside gl_ns_advance a batch that will be refilled by the register walker (>= 1 tuning launch
            // in front of the production launch that overwrites the dead slots) defers its bookkeeping
            const bool defer_book = ns->defer_ok && !ns->sharded_loop && ns->cfg.tuning_cycles >= 1 &&
                                    !ns->force_smem_walker && ns->d <= model_reg_max_d(ns->model.kind);
            const int smem = (int)(sizeof(unsigned long long) * ((N + KC_CTAS - 1) / KC_CTAS));
            int rc0 = ensure_smem((const void *)kill_cluster_kernel, smem);
            if (rc0) return rc0;
            GL_CUDA(launch_pdl(kill_cluster_kernel, dim3(KC_CTAS), dim3(KC_THREADS), (size_t)smem, st, ns->S, tc_arg, epw,
                               k_req, n_target, (int)N, ns->d, ns->cfg.row_offset, n_global, ns->cfg.dead_capacity,
                               ns->buf.live, ns->buf.logl, ns->dead_list, ns->is_dead, ns->buf.dead_x, ns->buf.dead_logl,
                               ns->buf.dead_w, (long long *)ns->buf.dead_idx, ns->buf.dead_step, defer_book ? 1 : 0));
            long long kh = n_target - ns->h_n_dead;
            if (kh > k_req) kh = k_req;
            if (kh > N - 1) kh = N - 1;
            ns->book_deferred = defer_book && kh > 0 && ns->h_n_dead + kh < n_target;
            ns->book_k = (int)kh;
            ns->book_vals = ns->buf.logl;
            ns->book_idx = ns->dead_list;
        } else {
            const int smem = (int)(sizeof(unsigned long long) * N);
            int rc0 = ensure_smem((const void *)kill_small_kernel<false>, smem);
            if (rc0) return rc0;
            kill_small_kernel<false><<<1, KILL_SMALL_THREADS, smem, st>>>(
                ns->S, tc_arg, epw, k_req, n_target, (int)N, ns->d, ns->cfg.row_offset, n_global,
                ns->cfg.dead_capacity, ns->buf.live, ns->buf.logl, ns->dead_list, ns->is_dead, ns->buf.dead_x,
                ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx, ns->buf.dead_step, nullptr, nullptr,
                nullptr, nullptr, 0);
        }
        GL_CHECK_LAUNCH();
        prof_mark(ns, ns->ev_sel, st);
        if (ns->prof) ns->prof_sel_launches += 1;
        ns->h_launches += 1;
    } else {
        int rcf = flush_end_walk(ns, st);
        if (rcf) return rcf;
        begin_kernel<<<1, 1, 0, st>>>(ns->S, k_req, n_target, N);
        prof_mark(ns, ns->ev_sel, st);
        int rc = select_launch(ns->buf.logl, N, &ns->S->k_cur, 0, 0, ns->dead_list, nullptr, ns->is_dead, ns->sel, st);
        if (rc) return rc;
        prof_mark(ns, ns->ev_sel, st);
        if (ns->prof) ns->prof_sel_launches += SEL_PASSES + 2;
        ns->h_launches += 1 + SEL_PASSES + 2 + 2;     // begin + select + record + accumulate
        record_kernel<<<blocks_for((long long)k_req * ns->d, 256, 8), 256, 0, st>>>(
            ns->S, ns->dead_list, ns->buf.live, ns->buf.logl, ns->d, ns->cfg.row_offset, log(n_global + 1.0),
            ns->cfg.dead_capacity, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx,
            ns->buf.dead_step);
        accumulate_kernel<<<1, 1024, 0, st>>>(ns->S, ns->dead_list, ns->buf.logl, n_global, n_target,
                                              ns->cfg.dead_capacity, -1);
        GL_CHECK_LAUNCH();
    }
    long long k = n_target - ns->h_n_dead;
    if (k > k_req) k = k_req;
    if (k > N - 1) k = N - 1;
    if (k < 0) k = 0;
    ns->h_k_last = (int)k;
    ns->h_n_dead += k;
    ns->pending_refill = k > 0;
    return GL_OK;
}

static int end_walk_launch(gl_ns *ns, cudaStream_t st, long long evals_per_walker) {
    end_walk_kernel<<<blocks_for(ns->cfg.batch_k, 256, 4), 256, 0, st>>>(ns->S, ns->dead_list, ns->is_dead, ns->d,
                                                                          ns->cfg.tuning_cycles, evals_per_walker);
    GL_CHECK_LAUNCH();
    ns->h_launches += 1;
    ns->h_step += 1;
    ns->pending_refill = false;
    return GL_OK;
}

static int refill_launch(gl_ns *ns, cudaStream_t st) {
    prof_mark(ns, ns->ev_walk, st);
    int rc = walk_launch(walk_params(ns), ns->model, st);
    ns->book_deferred = false;
    if (rc) return rc;
    prof_mark(ns, ns->ev_walk, st);
    if (ns->prof) ns->prof_walk_launches += ns->cfg.tuning_cycles + 1;
    ns->h_launches += ns->cfg.tuning_cycles + 1;
    if (ns->defer_ok && next_op_folds_end_walk(ns)) {
        ns->end_walk_deferred = true;      // the next fused kill / candidates kernel finishes this refill
        ns->h_step += 1;
        ns->pending_refill = false;
        return GL_OK;
    }
    return end_walk_launch(ns, st, (20LL * ns->cfg.tuning_cycles + ns->cfg.iterations + ns->cfg.burn_in) * ns->d);
}

}  // namespace gl

extern "C" {

int64_t gl_ns_workspace_bytes(const gl_ns_config *cfg, int32_t d) {
    if (gl::check_config(cfg, d)) return GL_EINVAL;
    return (int64_t)gl::workspace_layout(*cfg, d).total;
}

int gl_ns_create(const gl_ns_config *cfg, const gl_prior *priors_host, int32_t d, const gl_model *model,
                 const gl_ns_buffers *b, gl_ns **out) {
    int rc = gl::check_config(cfg, d);
    if (rc) return rc;
    GL_REQUIRE(out != nullptr, "out is NULL");
    GL_REQUIRE(priors_host != nullptr && model != nullptr && b != nullptr, "NULL argument");
    GL_REQUIRE(model->dev.d == d, "model dimension %d != number of sampled parameters %d", model->dev.d, d);
    GL_REQUIRE(b->live && b->logl && b->dead_x && b->dead_logl && b->dead_w && b->dead_idx && b->dead_step &&
                   b->workspace,
               "NULL device buffer");
    for (int j = 0; j < d; ++j) {
        GL_REQUIRE(priors_host[j].kind >= GL_PRIOR_UNIFORM && priors_host[j].kind <= GL_PRIOR_LOGUNIFORM,
                   "prior %d: unknown kind %d", j, priors_host[j].kind);
    }
    int dev = 0;
    GL_CUDA(cudaGetDevice(&dev));          // fail loudly when no device is usable
    gl_ns *ns = new (std::nothrow) gl_ns();
    if (!ns) return GL_ENOMEM;
    ns->cfg = *cfg;
    ns->d = d;
    ns->model = model->dev;
    ns->model_handle = model;
    memcpy(ns->priors, priors_host, sizeof(gl_prior) * d);
    ns->buf = *b;
    const gl::WorkspaceLayout L = gl::workspace_layout(*cfg, d);
    char *w = (char *)b->workspace;
    ns->S = (gl::EngineState *)(w + L.off_state);
    ns->sel = (gl::SelectState *)(w + L.off_sel);
    ns->pt = (gl::PriorTable *)(w + L.off_pt);
    ns->dead_list = (long long *)(w + L.off_list);
    ns->is_dead = (unsigned char *)(w + L.off_isdead);
    ns->scratch_x = (double *)(w + L.off_sx);
    ns->scratch_l = (double *)(w + L.off_sl);
    ns->tmp_k = (double *)(w + L.off_tmp);
    ns->fin_partials = (double *)(w + L.off_fin);
    ns->cand_list = (long long *)(w + L.off_cand);
    ns->positions = (long long *)(w + L.off_pos);
    ns->sel2 = (gl::SelectState *)(w + L.off_sel2);
    ns->sk0 = (unsigned long long *)(w + L.off_sk0);
    ns->sk1 = (unsigned long long *)(w + L.off_sk1);
    ns->si0 = (unsigned int *)(w + L.off_si0);
    ns->si1 = (unsigned int *)(w + L.off_si1);
    ns->select_path = 0;
    ns->h_sorted_ready = false;
    ns->h_n_dead = ns->h_n_evals = ns->h_launches = 0;
    ns->h_step = ns->h_k_last = 0;
    ns->initialized = ns->pending_refill = ns->finalized = false;
    ns->prof = false;
    ns->force_multipass = false;
    ns->end_walk_deferred = false;
    ns->defer_ok = false;
    ns->sharded_loop = false;
    ns->commit_refills = false;
    ns->force_smem_walker = false;
    ns->book_deferred = false;
    ns->walk_k_max = 0;
    ns->prof_walk_launches = ns->prof_sel_launches = 0;
    *out = ns;
    return GL_OK;
}

int gl_ns_destroy(gl_ns *ns) {
    if (ns) {
        for (cudaEvent_t e : ns->ev_walk) cudaEventDestroy(e);
        for (cudaEvent_t e : ns->ev_sel) cudaEventDestroy(e);
        if (ns->graph_exec) cudaGraphExecDestroy(ns->graph_exec);
    }
    delete ns;
    return GL_OK;
}

// gl_ns_advance recorded into a CUDA graph and launched as one unit: the launch sequence of a run with a closed-form
// death count involves no host decision, so the ~4 launches per step (programmatic-dependent-launch edges included)
// can be replayed by the device front end instead of being issued one by one from a host thread.  Used by
// gl_ns_advance_many, where the host's launch rate is what limits concurrent runs.
int gl_ns_advance_graph(gl_ns *ns, int64_t n_deaths_total, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    GL_REQUIRE(st != nullptr, "graph capture needs a non-default stream");
    GL_REQUIRE(!ns->prof, "per-kernel profiling events cannot be captured");
    if (ns->graph_exec) { cudaGraphExecDestroy(ns->graph_exec); ns->graph_exec = nullptr; }
    GL_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = gl_ns_advance(ns, n_deaths_total, stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != GL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return gl::cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    const cudaError_t e2 = cudaGraphInstantiate(&ns->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) return gl::cuda_fail(e2, "cudaGraphInstantiate", __FILE__, __LINE__);
    GL_CUDA(cudaGraphLaunch(ns->graph_exec, st));
    return GL_OK;
}

int gl_ns_init(gl_ns *ns, int32_t skip_population, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const gl::WorkspaceLayout L = gl::workspace_layout(ns->cfg, ns->d);
    GL_CUDA(cudaMemsetAsync(ns->buf.workspace, 0, L.total, st));
    gl::EngineState h;
    memset(&h, 0, sizeof(h));
    h.M = -INFINITY;
    h.lstar = -INFINITY;
    h.n_evals = skip_population ? 0ull : (unsigned long long)ns->cfg.population_size;
    GL_CUDA(cudaMemcpyAsync(ns->S, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    gl::PriorTable pt;
    memset(&pt, 0, sizeof(pt));
    for (int j = 0; j < ns->d; ++j) {
        pt.kind[j] = ns->priors[j].kind;
        pt.p0[j] = ns->priors[j].p0;
        pt.p1[j] = ns->priors[j].p1;
    }
    GL_CUDA(cudaMemcpyAsync(ns->pt, &pt, sizeof(pt), cudaMemcpyHostToDevice, st));
    GL_CUDA(cudaStreamSynchronize(st));    // h / pt are stack objects
    const long long N = ns->cfg.population_size;
    if (!skip_population) {
        int rc = gl::prior_draw_launch(ns->priors, ns->d, ns->cfg.seed, gl::GL_TAG_INIT, ns->cfg.row_offset, N,
                                       ns->buf.live, st);
        if (rc) return rc;
        rc = gl::loglike_eval_launch(ns->model_handle, ns->buf.live, N, ns->buf.logl, st);
        if (rc) return rc;
        ns->h_n_evals = N;
        ns->h_launches += 2;
    } else {
        ns->h_n_evals = 0;
    }
    ns->h_launches += 1;
    gl::init_widths_kernel<<<ns->d, 128, 0, st>>>(ns->S, ns->pt, ns->d, ns->cfg.seed);
    GL_CHECK_LAUNCH();
    ns->h_n_dead = 0;
    ns->h_step = ns->h_k_last = 0;
    ns->h_sorted_ready = false;
    ns->initialized = true;
    ns->pending_refill = ns->finalized = false;
    return GL_OK;
}

int gl_ns_kill(gl_ns *ns, int32_t k, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (!ns->initialized || ns->finalized || ns->pending_refill) {
        gl::set_error("gl_ns_kill: call gl_ns_init first / refill the previous batch / not after finalize");
        return GL_ESTATE;
    }
    GL_REQUIRE(k >= 1 && k <= ns->cfg.batch_k, "k=%d outside [1, batch_k=%d]", k, ns->cfg.batch_k);
    return gl::kill_launch(ns, k, INT64_MAX, (cudaStream_t)stream);
}

int gl_ns_refill(gl_ns *ns, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (!ns->pending_refill || ns->finalized) {
        gl::set_error("gl_ns_refill: no batch is waiting for replacement");
        return GL_ESTATE;
    }
    return gl::refill_launch(ns, (cudaStream_t)stream);
}

int gl_ns_inject(gl_ns *ns, const double *points_dev, const double *logl_dev, int32_t k, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (!ns->pending_refill || ns->finalized) {
        gl::set_error("gl_ns_inject: no batch is waiting for replacement");
        return GL_ESTATE;
    }
    GL_REQUIRE(k == ns->h_k_last, "k=%d but the waiting batch has %d slots", k, ns->h_k_last);
    GL_REQUIRE(points_dev && logl_dev, "NULL device pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = gl::blocks_for(k, 256, 4);
    gl::gather_logl_kernel<<<blocks, 256, 0, st>>>(ns->S, ns->dead_list, ns->buf.logl, ns->tmp_k);
    gl::inject_kernel<<<blocks, 256, 0, st>>>(ns->S, ns->dead_list, ns->tmp_k, points_dev, logl_dev, ns->d,
                                              ns->buf.live, ns->buf.logl);
    GL_CHECK_LAUNCH();
    ns->h_launches += 2;
    return gl::end_walk_launch(ns, st, 0);
}

int gl_ns_advance(gl_ns *ns, int64_t n_deaths_total, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (!ns->initialized || ns->finalized) {
        gl::set_error("gl_ns_advance: engine not initialised or already finalized");
        return GL_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ns->defer_ok = true;                   // every refill below is followed by a kill
    int rc = GL_OK;
    if (ns->pending_refill && ns->h_n_dead < n_deaths_total) rc = gl::refill_launch(ns, st);
    while (rc == GL_OK && ns->h_n_dead < n_deaths_total) {
        rc = gl::kill_launch(ns, ns->cfg.batch_k, n_deaths_total, st);
        if (rc) break;
        if (ns->h_k_last <= 0) break;
        if (ns->h_n_dead < n_deaths_total) rc = gl::refill_launch(ns, st);
    }
    ns->defer_ok = false;
    if (rc == GL_OK) rc = gl::flush_end_walk(ns, st);
    return rc;
}

// R independent runs (the seeds of one problem, the models of a selection sweep) advanced concurrently: one host
// thread per engine enqueues that engine's whole run on its own stream, so the launches of the runs interleave and
// their kernels overlap on the device -- a single small run offers a few tens of thousands of threads to a machine
// with 3*10^5 of them (HypSelector.run_nested_sampling runs its models one after another, hyp_selector.py:236-248).
static int advance_many_impl(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams, bool graphs);

int gl_ns_advance_many(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams) {
    return advance_many_impl(engines, n, n_deaths_total, streams, false);
}

int gl_ns_advance_many_graphs(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams) {
    return advance_many_impl(engines, n, n_deaths_total, streams, true);
}

}  // extern "C"

static int advance_many_impl(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams, bool graphs) {
    GL_REQUIRE(engines != nullptr && streams != nullptr && n_deaths_total != nullptr && n >= 1, "bad argument");
    for (int i = 0; i < n; ++i) GL_REQUIRE(engines[i] != nullptr, "engine %d is NULL", i);
    int dev = 0;
    GL_CUDA(cudaGetDevice(&dev));
    std::vector<int> rcs((size_t)n, GL_OK);
    std::vector<std::string> msgs((size_t)n);
    std::vector<std::thread> threads;
    threads.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        threads.emplace_back([&, i]() {
            cudaSetDevice(dev);
            rcs[(size_t)i] = graphs ? gl_ns_advance_graph(engines[i], n_deaths_total[i], streams[i])
                                    : gl_ns_advance(engines[i], n_deaths_total[i], streams[i]);
            if (rcs[(size_t)i] != GL_OK) msgs[(size_t)i] = gl::error_buffer();
        });
    for (auto &t : threads) t.join();
    for (int i = 0; i < n; ++i)
        if (rcs[(size_t)i] != GL_OK) {
            gl::set_error("engine %d: %s", i, msgs[(size_t)i].c_str());
            return rcs[(size_t)i];
        }
    return GL_OK;
}

extern "C" {

int gl_ns_finalize(gl_ns *ns, void *stream) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (!ns->initialized || ns->finalized) {
        gl::set_error("gl_ns_finalize: engine not initialised or already finalized");
        return GL_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = ns->cfg.population_size;
    long long blocks = (N + 4095) / 4096;
    if (blocks > gl::FIN_MAX_BLOCKS) blocks = gl::FIN_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    long long chunk = (N + blocks - 1) / blocks;
    blocks = (N + chunk - 1) / chunk;
    gl::final_partial_kernel<<<(int)blocks, gl::FIN_THREADS, 0, st>>>(ns->buf.logl, ns->is_dead, N, chunk,
                                                                      ns->fin_partials);
    gl::final_combine_kernel<<<1, 1, 0, st>>>(ns->S, ns->fin_partials, (int)blocks, 0, (int)blocks);
    gl::final_rows_kernel<<<gl::blocks_for(N, 256, 8), 256, 0, st>>>(
        ns->S, ns->dead_list, ns->is_dead, ns->buf.live, ns->buf.logl, N, ns->d, ns->cfg.row_offset,
        ns->cfg.dead_capacity, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx,
        ns->buf.dead_step);
    gl::final_done_kernel<<<1, 1, 0, st>>>(ns->S, ns->cfg.dead_capacity);
    GL_CHECK_LAUNCH();
    ns->h_launches += 4;
    ns->finalized = true;
    return GL_OK;
}

int gl_ns_export_dead(gl_ns *ns, double *table_dev, void *stream) {
    GL_REQUIRE(ns != nullptr && table_dev != nullptr, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    int seg = 2;
    while (seg < ns->cfg.batch_k) seg <<= 1;
    if (seg > 8192) seg = 8192;                    // 128 KB of (key, row) pairs
    const int smem = seg * 16;
    int rc = gl::ensure_smem((const void *)gl::export_batches_kernel, smem);
    if (rc) return rc;
    const int n_steps = ns->h_step + 1;            // batches recorded (the last one is never refilled)
    gl::export_batches_kernel<<<n_steps, gl::EXPORT_THREADS, smem, st>>>(
        ns->S, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w, (const long long *)ns->buf.dead_idx,
        ns->buf.dead_step, ns->d, seg, table_dev);
    GL_CHECK_LAUNCH();
    ns->h_launches += 1;
    return GL_OK;
}

int gl_export_gathered_table(const double *packed_dev, const int64_t *counts_dev, int32_t world, int64_t max_rows,
                             int32_t d, int32_t batch_k, int32_t n_steps, int64_t n_dead_total, double *table_dev,
                             void *stream) {
    GL_REQUIRE(packed_dev != nullptr && counts_dev != nullptr && table_dev != nullptr, "NULL argument");
    GL_REQUIRE(world >= 1 && world <= 64 && d >= 1 && batch_k >= 1 && n_steps >= 0, "bad argument");
    GL_REQUIRE(batch_k <= 8192, "batch_k=%d > 8192: order such tables on the host", batch_k);
    GL_REQUIRE((long long)world * max_rows < (1LL << 32), "gathered log too large");
    int seg = 2;
    while (seg < batch_k) seg <<= 1;
    const int smem = seg * 20;
    int rc = gl::ensure_smem((const void *)gl::export_gathered_kernel, smem);
    if (rc) return rc;
    gl::export_gathered_kernel<<<n_steps + world, gl::EXPORT_THREADS, smem, (cudaStream_t)stream>>>(
        packed_dev, (const long long *)counts_dev, world, max_rows, d, batch_k, n_steps, n_dead_total, seg, table_dev);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

int gl_ns_read_state(gl_ns *ns, gl_ns_state *out, void *stream) {
    GL_REQUIRE(ns != nullptr && out != nullptr, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    gl::EngineState h;
    GL_CUDA(cudaMemcpyAsync(&h, ns->S, sizeof(h), cudaMemcpyDeviceToHost, st));
    GL_CUDA(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    out->n_dead = h.n_dead;
    out->n_rows = h.n_rows;
    out->n_evals = (int64_t)h.n_evals;
    out->n_launches = ns->h_launches;
    out->step = h.step;
    out->finalized = h.finalized;
    out->shard_overflow = h.shard_overflow;
    out->reserved = 0;
    out->log_x = h.logX;
    out->lstar = h.lstar;
    out->z_linear = h.Z_lin;
    out->h_linear = h.H_lin;
    out->lse_m = h.M;
    out->lse_z = h.Zs;
    out->lse_h = h.Hs;
    if (h.Zs > 0.0) {
        out->log_z = h.M + log(h.Zs);
        out->information = -out->log_z + h.Hs / h.Zs;
    } else {
        out->log_z = -INFINITY;
        out->information = 0.0;
    }
    memcpy(out->widths, h.widths, sizeof(double) * GL_MAX_DIM);
    return GL_OK;
}

}  // extern "C"

extern "C" {

// ---- sharded run (one engine per rank; the caller all-gathers between the two calls) ------------
int gl_ns_shard_candidates(gl_ns *ns, int32_t k, double *cand_logl_dev, void *stream) {
    GL_REQUIRE(ns != nullptr && cand_logl_dev != nullptr, "NULL argument");
    if (!ns->initialized || ns->finalized || ns->pending_refill) {
        gl::set_error("gl_ns_shard_candidates: engine not ready (init first / refill the previous batch)");
        return GL_ESTATE;
    }
    GL_REQUIRE(k >= 1 && k <= ns->cfg.batch_k && k < ns->cfg.population_size, "k=%d outside [1, min(batch_k, N_local-1)]", k);
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = ns->cfg.population_size;
    gl::prof_mark(ns, ns->ev_sel, st);
    if (N <= gl::KC_MAX_N && !ns->force_multipass) {
        const int smem = (int)(8 * ((N + gl::KC_CTAS - 1) / gl::KC_CTAS));
        int rc = gl::ensure_smem((const void *)gl::shard_candidates_cluster_kernel, smem);
        if (rc) return rc;
        const int tc_arg = ns->end_walk_deferred ? ns->cfg.tuning_cycles : -1;
        ns->end_walk_deferred = false;
        const long long epw = (20LL * ns->cfg.tuning_cycles + ns->cfg.iterations + ns->cfg.burn_in) * ns->d;
        GL_CUDA(gl::launch_pdl(gl::shard_candidates_cluster_kernel, dim3(gl::KC_CTAS), dim3(gl::KC_THREADS), (size_t)smem,
                               st, ns->S, tc_arg, epw, ns->d, ns->dead_list, ns->is_dead, ns->buf.logl, (int)N, k,
                               ns->cand_list, cand_logl_dev));
        GL_CHECK_LAUNCH();
        gl::prof_mark(ns, ns->ev_sel, st);
        if (ns->prof) ns->prof_sel_launches += 1;
        ns->h_launches += 1;
        return GL_OK;
    }
    int rc = gl::flush_end_walk(ns, st);
    if (rc) return rc;
    rc = gl::select_launch(ns->buf.logl, N, nullptr, k, 0, ns->cand_list, nullptr, nullptr, ns->sel, st);
    if (rc) return rc;
    gl::prof_mark(ns, ns->ev_sel, st);
    if (ns->prof) ns->prof_sel_launches += gl::SEL_PASSES + 2;
    gl::shard_gather_kernel<<<gl::blocks_for(k, 256, 4), 256, 0, st>>>(ns->cand_list, ns->buf.logl, k, cand_logl_dev);
    GL_CHECK_LAUNCH();
    ns->h_launches += gl::SEL_PASSES + 2 + 1;
    return GL_OK;
}

int gl_ns_shard_commit_truncated(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank, int32_t k,
                                 int32_t m, void *keys_scratch, void *stream);

int gl_ns_shard_commit(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank, int32_t k,
                       void *stream) {
    return gl_ns_shard_commit_truncated(ns, gathered_logl_dev, world, rank, k, k, nullptr, stream);
}

int gl_ns_shard_commit_scratch(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank, int32_t k,
                               void *keys_scratch, void *stream) {
    return gl_ns_shard_commit_truncated(ns, gathered_logl_dev, world, rank, k, k, keys_scratch, stream);
}

int64_t gl_ns_local_population(const gl_ns *ns) { return ns ? ns->cfg.population_size : 0; }

// comm.cu's loop: every refill it issues is followed by gl_ns_shard_candidates (not part of the ABI)
void gl_ns_internal_sharded_loop(gl_ns *ns, int on) {
    ns->sharded_loop = on != 0;
    ns->defer_ok = on != 0;
    ns->commit_refills = false;
}
// ... and whether the commit it is about to issue will be followed by a refill
void gl_ns_internal_commit_refills(gl_ns *ns, int yes) { ns->commit_refills = yes != 0; }

int32_t gl_ns_shard_candidate_count(int32_t k, int32_t world, int64_t n_local, double sigmas) {
    if (k < 1 || world < 1 || n_local < 2) return 0;
    long long m = k;
    if (world > 1 && sigmas > 0.0) {
        // deaths per rank ~ Binomial(k, 1/world): mean + `sigmas` standard deviations + a small floor
        const double p = 1.0 / (double)world;
        m = (long long)ceil((double)k * p + sigmas * sqrt((double)k * p * (1.0 - p)) + 8.0);
    }
    if (m > k) m = k;
    if (m > n_local - 1) m = n_local - 1;
    if (m * world < k) m = (k + world - 1) / world;       // the union must hold k keys
    return (int32_t)m;
}

int gl_ns_shard_commit_truncated(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank, int32_t k,
                                 int32_t m, void *keys_scratch, void *stream) {
    GL_REQUIRE(ns != nullptr && gathered_logl_dev != nullptr, "NULL argument");
    GL_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad world/rank");
    GL_REQUIRE(k >= 1 && k <= ns->cfg.batch_k, "k=%d outside [1, batch_k]", k);
    GL_REQUIRE(m >= 1 && m <= k && (long long)m * world >= k, "candidates per rank m=%d outside [ceil(k/world), k]", m);
    cudaStream_t st = (cudaStream_t)stream;
    const double n_global = (double)ns->cfg.global_population;
    (void)keys_scratch;                    // (kept in the ABI; the cluster commit holds its keys in shared memory)
    if ((long long)world * m <= gl::KC_MAX_N && world <= gl::GL_MAX_PEERS && !ns->force_multipass) {
        const int smem = (int)(8 * (((long long)world * m + gl::KC_CTAS - 1) / gl::KC_CTAS));
        int rc0 = gl::ensure_smem((const void *)gl::shard_commit_cluster_kernel, smem);
        if (rc0) return rc0;
        // native loop, and a refill by the register walker follows this commit: it does the bookkeeping
        const bool defer_book = ns->sharded_loop && ns->commit_refills && ns->cfg.tuning_cycles >= 1 &&
                                !ns->force_smem_walker && ns->d <= gl::model_reg_max_d(ns->model.kind);
        gl::prof_mark(ns, ns->ev_sel, st);
        GL_CUDA(gl::launch_pdl(gl::shard_commit_cluster_kernel, dim3(gl::KC_CTAS), dim3(gl::KC_THREADS), (size_t)smem, st,
                               ns->S, gathered_logl_dev, world, rank, k, m, ns->d, ns->cfg.row_offset, n_global,
                               ns->cfg.dead_capacity, ns->cand_list, ns->positions, ns->buf.live, ns->buf.logl,
                               ns->dead_list, ns->is_dead, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w,
                               (long long *)ns->buf.dead_idx, ns->buf.dead_step, defer_book ? 1 : 0));
        ns->book_deferred = defer_book;
        ns->book_k = k;
        ns->book_vals = gathered_logl_dev;
        ns->book_idx = ns->positions;
        ns->walk_k_max = m;
        GL_CHECK_LAUNCH();
        gl::prof_mark(ns, ns->ev_sel, st);
        if (ns->prof) ns->prof_sel_launches += 1;
        ns->h_launches += 1;
        ns->h_n_dead += k;
        ns->h_k_last = -1;
        ns->pending_refill = true;
        return GL_OK;
    }
    GL_REQUIRE(world <= 256, "world=%d > 256", world);
    int rc = gl::select_launch(gathered_logl_dev, (long long)world * m, nullptr, k, 0, ns->positions, nullptr, nullptr,
                               ns->sel2, st);
    if (rc) return rc;
    gl::shard_local_kernel<<<1, 256, 0, st>>>(ns->S, ns->positions, ns->cand_list, k, m, world, rank, ns->dead_list,
                                              ns->is_dead);
    gl::record_kernel<<<gl::blocks_for((long long)k * ns->d, 256, 8), 256, 0, st>>>(
        ns->S, ns->dead_list, ns->buf.live, ns->buf.logl, ns->d, ns->cfg.row_offset, log(n_global + 1.0),
        ns->cfg.dead_capacity, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx,
        ns->buf.dead_step);
    gl::accumulate_kernel<<<1, 1024, 0, st>>>(ns->S, ns->positions, gathered_logl_dev, n_global, INT64_MAX,
                                              ns->cfg.dead_capacity, k);
    GL_CHECK_LAUNCH();
    ns->h_launches += gl::SEL_PASSES + 2 + 3;
    ns->h_n_dead += k;
    ns->h_k_last = -1;                     // local batch size is only known on the device
    ns->pending_refill = true;
    ns->walk_k_max = m;
    return GL_OK;
}

// comm.cu's fused step (not part of the ABI): candidates + peer-memory exchange + commit in one launch
int gl_ns_internal_shard_step_p2p(gl_ns *ns, const gl::PeerTables *tables, int32_t world, int32_t rank, int32_t k,
                                  int32_t m, uint32_t epoch, void *stream) {
    GL_REQUIRE(ns != nullptr && tables != nullptr, "NULL argument");
    if (!ns->initialized || ns->finalized || ns->pending_refill) {
        gl::set_error("shard step: engine not ready (init first / refill the previous batch)");
        return GL_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = ns->cfg.population_size;
    const double n_global = (double)ns->cfg.global_population;
    GL_REQUIRE(N <= gl::KC_MAX_N && (long long)world * m <= gl::KC_MAX_N && world <= gl::GL_MAX_PEERS && m <= tables->cap,
               "shard step: sizes outside the fused kernel's range");
    const long long nmax = N > (long long)world * m ? N : (long long)world * m;
    const int smem = (int)(8 * ((nmax + gl::KC_CTAS - 1) / gl::KC_CTAS));
    int rc = gl::ensure_smem((const void *)gl::shard_step_p2p_kernel, smem);
    if (rc) return rc;
    const int tc_arg = ns->end_walk_deferred ? ns->cfg.tuning_cycles : -1;
    ns->end_walk_deferred = false;
    const long long epw = (20LL * ns->cfg.tuning_cycles + ns->cfg.iterations + ns->cfg.burn_in) * ns->d;
    const bool defer_book = ns->sharded_loop && ns->commit_refills && ns->cfg.tuning_cycles >= 1 &&
                            !ns->force_smem_walker && ns->d <= gl::model_reg_max_d(ns->model.kind);
    gl::prof_mark(ns, ns->ev_sel, st);
    GL_CUDA(gl::launch_pdl(gl::shard_step_p2p_kernel, dim3(gl::KC_CTAS), dim3(gl::KC_THREADS), (size_t)smem, st, ns->S,
                           tc_arg, epw, (int)N, m, world, rank, k, ns->d, ns->cfg.row_offset, n_global,
                           ns->cfg.dead_capacity, *tables, epoch, ns->cand_list, ns->positions, ns->buf.live,
                           ns->buf.logl, ns->dead_list, ns->is_dead, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w,
                           (long long *)ns->buf.dead_idx, ns->buf.dead_step, defer_book ? 1 : 0,
                           gl::peer_timeout_ns_host()));
    gl::prof_mark(ns, ns->ev_sel, st);
    if (ns->prof) ns->prof_sel_launches += 1;
    ns->h_launches += 1;
    ns->h_n_dead += k;
    ns->h_k_last = -1;
    ns->pending_refill = true;
    ns->walk_k_max = m;
    ns->book_deferred = defer_book;
    ns->book_k = k;
    ns->book_vals = tables->gather[rank] + (long long)(epoch & 1u) * (long long)world * tables->cap;
    ns->book_idx = ns->positions;
    return GL_OK;
}

// partial[4] = (m, s, linear sum, count) of this rank's survivors (nested_sampling.py:184-187)
int gl_ns_shard_final_partial(gl_ns *ns, double *partial_dev, void *stream) {
    GL_REQUIRE(ns != nullptr && partial_dev != nullptr, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = ns->cfg.population_size;
    gl::final_partial_kernel<<<1, gl::FIN_THREADS, 0, st>>>(ns->buf.logl, ns->is_dead, N, N, partial_dev);
    GL_CHECK_LAUNCH();
    ns->h_launches += 1;
    return GL_OK;
}

int gl_ns_shard_finalize(gl_ns *ns, const double *gathered_partials_dev, int32_t world, int32_t rank, void *stream) {
    GL_REQUIRE(ns != nullptr && gathered_partials_dev != nullptr && world >= 1 && rank >= 0 && rank < world,
               "bad argument");
    if (!ns->initialized || ns->finalized) {
        gl::set_error("gl_ns_shard_finalize: engine not initialised or already finalized");
        return GL_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = ns->cfg.population_size;
    gl::final_combine_kernel<<<1, 1, 0, st>>>(ns->S, gathered_partials_dev, world, rank, 1);
    gl::final_rows_kernel<<<gl::blocks_for(N, 256, 8), 256, 0, st>>>(
        ns->S, ns->dead_list, ns->is_dead, ns->buf.live, ns->buf.logl, N, ns->d, ns->cfg.row_offset,
        ns->cfg.dead_capacity, ns->buf.dead_x, ns->buf.dead_logl, ns->buf.dead_w, (long long *)ns->buf.dead_idx,
        ns->buf.dead_step);
    gl::final_done_kernel<<<1, 1, 0, st>>>(ns->S, ns->cfg.dead_capacity);
    GL_CHECK_LAUNCH();
    ns->h_launches += 3;
    ns->finalized = true;
    return GL_OK;
}

// ---- the reference's per-call sampler protocol (samplers.py:56): one walker -------------------------
namespace gl {
__global__ void single_walker_setup_kernel(EngineState *S, long long *dead_list, unsigned char *is_dead, double lstar,
                                           int step, const double *widths_in, int d) {
    S->k_cur = 1;
    S->do_walk = 1;
    S->lstar = lstar;
    S->step = step;
    dead_list[0] = 0;                 // slot 0 is the one to refill, slot 1 (the start point) the only survivor
    is_dead[0] = 1;
    is_dead[1] = 0;
    if (widths_in)
        for (int j = 0; j < d; ++j) S->widths[j] = widths_in[j];
}
}  // namespace gl

int gl_sampler_call_host(const gl_prior *priors_host, int32_t d, const gl_model *model, int32_t iterations,
                         int32_t burn_in, int32_t tuning_cycles, int32_t proposal, uint64_t seed, int32_t call_index,
                         const double *start_host, double ns_boundary, double *widths_inout_host, int32_t have_widths,
                         double *point_out_host, double *logl_out_host) {
    GL_REQUIRE(priors_host && model && start_host && widths_inout_host && point_out_host && logl_out_host,
               "NULL argument");
    gl_ns_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.population_size = 2;
    cfg.batch_k = 1;
    cfg.iterations = iterations;
    cfg.burn_in = burn_in;
    cfg.tuning_cycles = tuning_cycles;
    cfg.proposal = proposal;
    cfg.seed = seed;
    cfg.dead_capacity = 4;
    cfg.global_population = 2;
    const int64_t ws = gl_ns_workspace_bytes(&cfg, d);
    if (ws < 0) return (int)ws;
    const size_t D = (size_t)d;
    const size_t n_dbl = 2 * D + 2 + 4 * D + 4 + 4 + D;      // live, logl, dead_x, dead_logl, dead_w, widths
    char *slab = nullptr;
    GL_CUDA(cudaMalloc(&slab, sizeof(double) * n_dbl + 4 * 8 + 4 * 4 + 256 + (size_t)ws));
    double *p = (double *)slab;
    gl_ns_buffers b;
    b.live = p; p += 2 * D;
    b.logl = p; p += 2;
    b.dead_x = p; p += 4 * D;
    b.dead_logl = p; p += 4;
    b.dead_w = p; p += 4;
    double *widths_dev = p; p += D;
    b.dead_idx = (int64_t *)p; p += 4;
    b.dead_step = (int32_t *)p; p += 2;
    b.workspace = (void *)(((uintptr_t)p + 255) / 256 * 256);
    gl_ns *ns = nullptr;
    int rc = gl_ns_create(&cfg, priors_host, d, model, &b, &ns);
    cudaError_t e = cudaSuccess;
    if (rc == GL_OK) {
        e = cudaMemcpy(b.live, start_host, sizeof(double) * D, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(b.live + D, start_host, sizeof(double) * D, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && have_widths)
            e = cudaMemcpy(widths_dev, widths_inout_host, sizeof(double) * D, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = gl::cuda_fail(e, "sampler call upload", __FILE__, __LINE__);
    }
    if (rc == GL_OK) rc = gl::loglike_eval_launch(model, b.live, 2, b.logl, 0);          // samplers.py:83
    if (rc == GL_OK) rc = gl_ns_init(ns, 1, 0);            // state + widths from 100 prior draws (:70-81)
    if (rc == GL_OK) {
        gl::single_walker_setup_kernel<<<1, 1>>>(ns->S, ns->dead_list, ns->is_dead, ns_boundary, call_index,
                                                 have_widths ? widths_dev : nullptr, d);
        ns->pending_refill = true;
        ns->h_k_last = 1;
        rc = gl_ns_refill(ns, 0);
    }
    if (rc == GL_OK) {
        gl_ns_state st;
        rc = gl_ns_read_state(ns, &st, 0);
        if (rc == GL_OK) {
            memcpy(widths_inout_host, st.widths, sizeof(double) * D);
            e = cudaMemcpy(point_out_host, b.live, sizeof(double) * D, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(logl_out_host, b.logl, sizeof(double), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = gl::cuda_fail(e, "sampler call download", __FILE__, __LINE__);
        }
    }
    if (ns) gl_ns_destroy(ns);
    cudaFree(slab);
    return rc;
}

int gl_ns_force_multipass_select(gl_ns *ns, int32_t enable) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    ns->force_multipass = enable != 0;
    return GL_OK;
}

int gl_ns_set_select_path(gl_ns *ns, int32_t path) {
    GL_REQUIRE(ns != nullptr && path >= 0 && path <= 4,
               "path must be 0 (auto), 1 (multi-launch), 2 (fused radix), 3 (fused sorted), 4 (cluster radix)");
    GL_REQUIRE(!ns->h_sorted_ready || path == 0 || path == 3, "cannot leave the sorted path once the run has started");
    ns->select_path = path;
    return GL_OK;
}

int gl_ns_force_smem_walker(gl_ns *ns, int32_t enable) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    ns->force_smem_walker = enable != 0;
    return GL_OK;
}

int gl_ns_profile_enable(gl_ns *ns, int32_t enable) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    if (enable && !ns->prof) {                       // a new measurement window: drop what an earlier one left
        for (cudaEvent_t e : ns->ev_walk) cudaEventDestroy(e);
        for (cudaEvent_t e : ns->ev_sel) cudaEventDestroy(e);
        ns->ev_walk.clear();
        ns->ev_sel.clear();
        ns->prof_walk_launches = ns->prof_sel_launches = 0;
    }
    ns->prof = enable != 0;
    return GL_OK;
}

int gl_ns_profile_read(gl_ns *ns, double *walk_ms, int64_t *walk_launches, double *select_ms,
                       int64_t *select_launches) {
    GL_REQUIRE(ns != nullptr, "ns is NULL");
    auto total = [](std::vector<cudaEvent_t> &v, double *out) -> int {
        double acc = 0.0;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            float ms = 0.f;
            GL_CUDA(cudaEventSynchronize(v[i + 1]));
            GL_CUDA(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
            acc += ms;
        }
        if (out) *out = acc;
        return GL_OK;
    };
    int rc = total(ns->ev_walk, walk_ms);
    if (rc) return rc;
    rc = total(ns->ev_sel, select_ms);
    if (rc) return rc;
    if (walk_launches) *walk_launches = ns->prof_walk_launches;
    if (select_launches) *select_launches = ns->prof_sel_launches;
    return GL_OK;
}

}  // extern "C"
