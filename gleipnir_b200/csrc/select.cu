// k-worst selection (nested_sampling.py:115,156 generalised from argmin to k deaths).
//
// Order: (sort_key(logL), index) ascending -- NaN first, -0.0 == +0.0, ties by lowest
// index, i.e. exactly what repeated np.argmin would remove.  Algorithm:
//   6 histogram passes (MSB radix-select, 11-bit digits) find T = key of the k-th worst
//   and r = how many key == T elements belong to the selection;
//   count + scatter do an index-ordered compaction (deterministic output order).
// HBM/L2-bound: each pass reads 8 B per live point (the logL array is L2-resident for
// every BASELINE config: 8 MB at N = 10^6).  Grid-wide hand-offs use the "last block
// finalises" pattern (atomic ticket + __threadfence), so a selection is 8 launches and
// needs no host round trip; k itself lives in device memory.
#include "select.cuh"

namespace gl {

__device__ __forceinline__ int pass_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
__device__ __forceinline__ int pass_bits(int pass) { return pass < 5 ? 11 : 9; }

__global__ void __launch_bounds__(SEL_THREADS) select_hist_kernel(const double *__restrict__ logl, long long n,
                                                                  long long chunk, const int *k_ptr, int k_val,
                                                                  SelectState *S, int pass) {
    __shared__ unsigned int h[SEL_BINS];
    __shared__ unsigned int warp_tot[SEL_THREADS / 32];
    __shared__ int is_last;
    int k = k_ptr ? *k_ptr : k_val;
    if (k > n) k = (int)n;
    if (k <= 0) return;                       // uniform over the grid: nothing to select
    for (int b = threadIdx.x; b < SEL_BINS; b += SEL_THREADS) h[b] = 0u;
    __syncthreads();
    const int shift = pass_shift(pass), nbits = pass_bits(pass);
    const unsigned int mask = (1u << nbits) - 1u;
    const unsigned long long prefix = pass == 0 ? 0ull : S->prefix;
    const int hi_shift = shift + nbits;       // 64 on pass 0 (handled by the pass == 0 test)
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n) hi = n;
    for (long long i = lo + threadIdx.x; i < hi; i += SEL_THREADS) {
        const unsigned long long key = sort_key(logl[i]);
        if (pass == 0 || (key >> hi_shift) == prefix) atomicAdd(&h[(unsigned int)(key >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < SEL_BINS; b += SEL_THREADS)
        if (h[b]) atomicAdd(&S->hist[pass][b], h[b]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&S->ticket[pass], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last block: locate the digit where the running count crosses k_rem ----------
    const int k_rem = pass == 0 ? k : S->k_rem;
    constexpr int PER = SEL_BINS / SEL_THREADS;   // 8 consecutive bins per thread
    unsigned int c[PER], mine = 0u;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        c[q] = __ldcg(&S->hist[pass][threadIdx.x * PER + q]);
        mine += c[q];
    }
    // exclusive scan of `mine` over the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned int base = 0u;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    unsigned int run = base + incl - mine;        // elements in bins before this thread's first bin
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        if (run < (unsigned int)k_rem && run + c[q] >= (unsigned int)k_rem) {   // exactly one (thread, q)
            const unsigned long long digit = (unsigned long long)(threadIdx.x * PER + q);
            const unsigned long long np = (prefix << nbits) | digit;
            S->prefix = np;
            S->k_rem = k_rem - (int)run;
            if (pass == SEL_PASSES - 1) {
                S->T = np;
                S->r_eq = k_rem - (int)run;
                S->k_eff = k;
            }
        }
        run += c[q];
    }
    // self-clean for the next selection
#pragma unroll
    for (int q = 0; q < PER; ++q) S->hist[pass][threadIdx.x * PER + q] = 0u;
    if (threadIdx.x == 0) S->ticket[pass] = 0u;
}

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const double *__restrict__ logl, long long n,
                                                                   long long chunk, const int *k_ptr, int k_val,
                                                                   SelectState *S) {
    __shared__ int s_lt[SEL_THREADS / 32], s_eq[SEL_THREADS / 32];
    int k = k_ptr ? *k_ptr : k_val;
    if (k <= 0 || n <= 0) return;
    const unsigned long long T = S->T;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n) hi = n;
    int lt = 0, eq = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += SEL_THREADS) {
        const unsigned long long key = sort_key(logl[i]);
        lt += key < T;
        eq += key == T;
    }
    lt = warp_sum_i(lt);
    eq = warp_sum_i(eq);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_lt[wid] = lt; s_eq[wid] = eq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { a += s_lt[w]; b += s_eq[w]; }
        S->block_lt[blockIdx.x] = a;
        S->block_eq[blockIdx.x] = b;
    }
}

__global__ void __launch_bounds__(SEL_THREADS) select_scatter_kernel(const double *__restrict__ logl, long long n,
                                                                     long long chunk, const int *k_ptr, int k_val,
                                                                     long long idx_base, long long *__restrict__ idx_out,
                                                                     double *lstar_out, unsigned char *is_dead,
                                                                     SelectState *S) {
    __shared__ int s_a[SEL_THREADS / 32], s_b[SEL_THREADS / 32];
    __shared__ unsigned int s_scan[SEL_THREADS / 32];
    __shared__ int s_base_lt, s_base_eq;
    int k = k_ptr ? *k_ptr : k_val;
    if (k <= 0 || n <= 0) return;
    const unsigned long long T = S->T;
    const int r = S->r_eq;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // selected-before-this-block counters
    int a = 0, b = 0;
    for (int q = threadIdx.x; q < (int)blockIdx.x; q += SEL_THREADS) { a += S->block_lt[q]; b += S->block_eq[q]; }
    a = warp_sum_i(a);
    b = warp_sum_i(b);
    if (lane == 0) { s_a[wid] = a; s_b[wid] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int x = 0, y = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { x += s_a[w]; y += s_b[w]; }
        s_base_lt = x;
        s_base_eq = y;
        if (blockIdx.x == 0 && lstar_out) *lstar_out = key_to_double(T);
    }
    __syncthreads();
    int run_lt = s_base_lt, run_eq = s_base_eq;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n) hi = n;
    for (long long t0 = lo; t0 < hi; t0 += SEL_THREADS) {
        const long long i = t0 + threadIdx.x;
        unsigned int is_lt = 0u, is_eq = 0u;
        if (i < hi) {
            const unsigned long long key = sort_key(logl[i]);
            is_lt = key < T;
            is_eq = key == T;
        }
        // packed block scan: low 16 bits count lt, high 16 bits count eq
        const unsigned int mine = is_lt | (is_eq << 16);
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();                     // previous tile's s_scan fully consumed
        if (lane == 31) s_scan[wid] = incl;
        __syncthreads();
        unsigned int base = 0u, total = 0u;
        for (int w = 0; w < SEL_THREADS / 32; ++w) {
            const unsigned int v = s_scan[w];
            if (w < wid) base += v;
            total += v;
        }
        const unsigned int excl = base + incl - mine;
        const int lt_before = run_lt + (int)(excl & 0xffffu);
        const int eq_before = run_eq + (int)(excl >> 16);
        const bool sel = is_lt || (is_eq && eq_before < r);
        if (sel) {
            const int pos = lt_before + (eq_before < r ? eq_before : r);
            idx_out[pos] = idx_base + i;
            if (is_dead) is_dead[i] = 1;
        }
        run_lt += (int)(total & 0xffffu);
        run_eq += (int)(total >> 16);
    }
}

SelectPlan select_plan(long long n) {
    SelectPlan p;
    long long blocks = (n + 2047) / 2048;
    const long long cap = (long long)sm_count() * 4 < SEL_MAX_BLOCKS ? (long long)sm_count() * 4 : SEL_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    long long chunk = (n + blocks - 1) / blocks;
    chunk = ((chunk + SEL_THREADS - 1) / SEL_THREADS) * SEL_THREADS;
    if (chunk < SEL_THREADS) chunk = SEL_THREADS;
    blocks = (n + chunk - 1) / chunk;
    if (blocks < 1) blocks = 1;
    p.blocks = (int)blocks;
    p.chunk = chunk;
    return p;
}

int select_launch(const double *logl, long long n, const int *k_ptr, int k_val, long long idx_base,
                  long long *idx_out, double *lstar_out, unsigned char *is_dead, SelectState *S, cudaStream_t st) {
    GL_REQUIRE(n >= 1, "n must be >= 1");
    GL_REQUIRE(logl && idx_out && S, "NULL device pointer");
    const SelectPlan p = select_plan(n);
    for (int pass = 0; pass < SEL_PASSES; ++pass)
        select_hist_kernel<<<p.blocks, SEL_THREADS, 0, st>>>(logl, n, p.chunk, k_ptr, k_val, S, pass);
    select_count_kernel<<<p.blocks, SEL_THREADS, 0, st>>>(logl, n, p.chunk, k_ptr, k_val, S);
    select_scatter_kernel<<<p.blocks, SEL_THREADS, 0, st>>>(logl, n, p.chunk, k_ptr, k_val, idx_base, idx_out,
                                                           lstar_out, is_dead, S);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

}  // namespace gl

extern "C" {

int64_t gl_select_workspace_bytes(int64_t n) {
    (void)n;
    return (int64_t)((sizeof(gl::SelectState) + 255) / 256 * 256);
}

int gl_select_kworst(const double *logl_dev, int64_t n, int32_t k, int64_t *idx_out_dev, double *lstar_dev,
                     void *workspace_dev, void *stream) {
    GL_REQUIRE(k >= 1 && k <= n, "k=%d outside [1, n=%lld]", k, (long long)n);
    GL_REQUIRE(workspace_dev != nullptr, "workspace is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    GL_CUDA(cudaMemsetAsync(workspace_dev, 0, sizeof(gl::SelectState), st));
    return gl::select_launch(logl_dev, n, nullptr, k, 0, (long long *)idx_out_dev, lstar_dev, nullptr,
                             (gl::SelectState *)workspace_dev, st);
}

}  // extern "C"
