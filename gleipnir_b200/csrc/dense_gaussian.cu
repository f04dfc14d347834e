// Dense correlated Gaussian, batched full quadratic form on the FP64 tensor cores (DMMA):
//     logL_i = -0.5 (x_i - mu)^T P (x_i - mu) + log_norm          
// tcgen05/UMMA has no FP64 type, so the sm_100a FP64 tensor path is mma.sync.m8n8k4.f64.
// One warp owns 8 rows at a time: the 8 x d tile of residuals R is staged in shared memory
// (rows are contiguous in HBM, so the load is fully coalesced), Y = R P is accumulated tile by
// tile with DMMA (A = R fragments held in registers, B = P pre-arranged in fragment order in
// shared memory, loaded once per CTA), and the row dot q_i = sum_j R_ij Y_ij is folded into the
// accumulator read-out, so Y never leaves registers.
// Algorithmic work: 2d^2 + 2d flop and 8d + 8 B per point (12.5 flop/B at d = 50).
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "likelihoods.cuh"

namespace gl {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int DG_WARPS = 8;

// Pf[(ks*NT + nt)*32 + lane] = P[ks*4 + (lane&3)][nt*8 + (lane>>2)]   (B fragment order)
__global__ void dense_gaussian_pack_kernel(const double *__restrict__ P, int d, int KT, int NT,
                                           double *__restrict__ Pf) {
    const int total = KT * NT * 32;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int lane = e & 31, tile = e >> 5;
        const int ks = tile / NT, nt = tile - ks * NT;
        const int r = ks * 4 + (lane & 3), c = nt * 8 + (lane >> 2);
        Pf[e] = (r < d && c < d) ? P[(size_t)r * d + c] : 0.0;
    }
}

// Round 2: a warp owns TWO 8-row tiles at a time and walks the n-tiles in pairs, so four independent accumulator
// pairs are in flight (a dependent chain of KT DMMAs per pair was the r1 limiter: 27 % of the DMMA peak) and every
// B fragment fetched from shared memory feeds two DMMAs.
template <int KTM>
__global__ void __launch_bounds__(DG_WARPS * 32, 2) dense_gaussian_dmma_kernel(
    const double *__restrict__ x, long long n, int d, int KT, int NT, const double *__restrict__ mu,
    const double *__restrict__ Pf, double log_norm, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int dp = KT * 4 > NT * 8 ? KT * 4 : NT * 8;          // padded row length
    double *Ps = sm;                                           // [KT*NT*32]
    double *Rs_all = Ps + (size_t)KT * NT * 32;                // [DG_WARPS][16][dp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int e = tid; e < KT * NT * 32; e += blockDim.x) Ps[e] = Pf[e];
    __syncthreads();
    double *Rs = Rs_all + (size_t)warp * 16 * dp;
    const int grp = lane >> 2, tig = lane & 3;
    const long long n_tiles = (n + 15) / 16;
    for (long long t = (long long)blockIdx.x * nw + warp; t < n_tiles; t += (long long)gridDim.x * nw) {
        const long long r0 = t * 16;
        const int rows = (int)((n - r0) < 16 ? (n - r0) : 16);
        // stage the residual tiles (coalesced: rows*d contiguous doubles), zero padding
        // (lane = column: no index division, rows coalesced).  All 16 row loads of a column pass are issued before
        // the first store: the staging phase is one memory latency per pass instead of four (it was as long as the
        // DMMA phase of the tile).
        for (int c = lane; c < dp; c += 32) {
            const double m = c < d ? __ldg(mu + c) : 0.0;
            double v[16];
            const double *src = x + r0 * d + c;
            if (rows == 16 && c < d) {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = __ldg(src + (size_t)r * d);
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = (r < rows && c < d) ? __ldg(src + (size_t)r * d) : m;
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) Rs[r * dp + c] = v[r] - m;
        }
        __syncwarp();
        double a0[KTM], a1[KTM];
#pragma unroll
        for (int ks = 0; ks < KTM; ++ks) {
            a0[ks] = ks < KT ? Rs[grp * dp + ks * 4 + tig] : 0.0;
            a1[ks] = ks < KT ? Rs[(8 + grp) * dp + ks * 4 + tig] : 0.0;
        }
        double q0 = 0.0, q1 = 0.0;
        int nt = 0;
        for (; nt + 1 < NT; nt += 2) {
            double c00 = 0.0, d00 = 0.0, c01 = 0.0, d01 = 0.0, c10 = 0.0, d10 = 0.0, c11 = 0.0, d11 = 0.0;
            const double *B0 = Ps + (size_t)nt * 32 + lane, *B1 = B0 + 32;
#pragma unroll
            for (int ks = 0; ks < KTM; ++ks)
                if (ks < KT) {
                    const double b0 = B0[(size_t)ks * NT * 32], b1 = B1[(size_t)ks * NT * 32];
                    dmma_m8n8k4(c00, d00, a0[ks], b0);
                    dmma_m8n8k4(c01, d01, a0[ks], b1);
                    dmma_m8n8k4(c10, d10, a1[ks], b0);
                    dmma_m8n8k4(c11, d11, a1[ks], b1);
                }
            // accumulator layout: row = grp, cols = nt*8 + tig*2 + {0,1}
            const int col = nt * 8 + tig * 2;
            const double *R0 = Rs + grp * dp + col, *R1 = Rs + (8 + grp) * dp + col;
            q0 = fma(c00, R0[0], q0); q0 = fma(d00, R0[1], q0); q0 = fma(c01, R0[8], q0); q0 = fma(d01, R0[9], q0);
            q1 = fma(c10, R1[0], q1); q1 = fma(d10, R1[1], q1); q1 = fma(c11, R1[8], q1); q1 = fma(d11, R1[9], q1);
        }
        if (nt < NT) {                                          // odd number of n-tiles: the last one alone
            double c00 = 0.0, d00 = 0.0, c10 = 0.0, d10 = 0.0;
            const double *B0 = Ps + (size_t)nt * 32 + lane;
#pragma unroll
            for (int ks = 0; ks < KTM; ++ks)
                if (ks < KT) {
                    const double b0 = B0[(size_t)ks * NT * 32];
                    dmma_m8n8k4(c00, d00, a0[ks], b0);
                    dmma_m8n8k4(c10, d10, a1[ks], b0);
                }
            const int col = nt * 8 + tig * 2;
            const double *R0 = Rs + grp * dp + col, *R1 = Rs + (8 + grp) * dp + col;
            q0 = fma(c00, R0[0], q0); q0 = fma(d00, R0[1], q0);
            q1 = fma(c10, R1[0], q1); q1 = fma(d10, R1[1], q1);
        }
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1);
        q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        if (tig == 0 && grp < rows) out[r0 + grp] = -0.5 * q0 + log_norm;
        if (tig == 0 && 8 + grp < rows) out[r0 + 8 + grp] = -0.5 * q1 + log_norm;
        __syncwarp();
    }
}

// Round 2, second form: the RAW rows of the next 16-row tile are copied into the warp's second shared-memory
// buffer with cp.async while the DMMAs of the current tile run, so the staging phase (one HBM latency per tile,
// as long as a quarter of the tile's DMMA time) disappears from the warp's critical path.  The tile is kept as
// the contiguous 16*d doubles it is in HBM (16-byte copies, no padding); mu is subtracted when the A fragments
// and the epilogue operands are read, and the padding columns (k >= d, n >= d) are masked there.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NTM>
__global__ void __launch_bounds__(DG_WARPS * 32, 2) dense_gaussian_dmma_pipe_kernel(
    const double *__restrict__ x, long long n, int d, int KT, const double *__restrict__ mu,
    const double *__restrict__ Pf, double log_norm, double *__restrict__ out) {
    // k-outer: per k-slice one A fragment per row tile and NTM B fragments feed 2*NTM DMMAs on 2*NTM INDEPENDENT
    // accumulator pairs (the n-outer form above has four chains of KT dependent DMMAs per warp)
    constexpr int NT = NTM;
    extern __shared__ double sm[];
    const int dp = KT * 4 > NT * 8 ? KT * 4 : NT * 8;
    const int tile_len = 16 * d;                               // doubles per raw tile (even: 16-byte copies)
    double *Ps = sm;                                           // [KT*NT*32]
    double *mus = Ps + (size_t)KT * NT * 32;                   // [dp + 16] mu, zero padded
    double *Xs_all = mus + dp + 16;                            // [warps][2][tile_len]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int e = tid; e < KT * NT * 32; e += blockDim.x) Ps[e] = Pf[e];
    for (int e = tid; e < dp + 16; e += blockDim.x) mus[e] = e < d ? mu[e] : 0.0;
    __syncthreads();
    double *Xw = Xs_all + (size_t)warp * 2 * tile_len;
    const int grp = lane >> 2, tig = lane & 3;
    const long long n_tiles = (n + 15) / 16;
    const long long stride = (long long)gridDim.x * nw;
    long long t = (long long)blockIdx.x * nw + warp;
    auto fetch = [&](long long tile, double *dst) {            // rows [16 tile, 16 tile + 16) clipped to n
        const long long r0 = tile * 16;
        const int len = (int)((n - r0 < 16 ? n - r0 : 16) * d);    // doubles
        const double *src = x + r0 * d;
        for (int e = lane * 2; e + 1 < len; e += 64) cp_async16(dst + e, src + e);
        if ((len & 1) && lane == 0) dst[len - 1] = src[len - 1];    // odd tail (last partial tile, odd d)
    };
    int buf = 0;
    if (t < n_tiles) fetch(t, Xw);
    cp_async_commit();
    for (; t < n_tiles; t += stride) {
        const long long tn = t + stride;
        if (tn < n_tiles) fetch(tn, Xw + (buf ^ 1) * tile_len);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const double *Rs = Xw + buf * tile_len;
        const long long r0 = t * 16;
        const int rows = (int)((n - r0) < 16 ? (n - r0) : 16);
        const bool ok0 = grp < rows, ok1 = 8 + grp < rows;
        const double *row0 = Rs + grp * d, *row1 = Rs + (8 + grp) * d;
        double c0[NT], e0[NT], c1[NT], e1[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) c0[nt] = e0[nt] = c1[nt] = e1[nt] = 0.0;
        const double *B = Ps + lane;
        // (a hand-made software pipeline over the k-slices -- fragments of slice ks+1 requested before the DMMAs of
        // slice ks -- measured slower: 18.9 vs 19.7 TFLOP/s at d = 50; the loads are not what the DMMAs wait for)
#pragma unroll 2
        for (int ks = 0; ks < KT; ++ks) {
            const int c = ks * 4 + tig;
            const bool in = c < d;
            const double m_ = mus[c];                          // zero beyond d
            const double a0 = (in && ok0) ? row0[c] - m_ : 0.0;
            const double a1 = (in && ok1) ? row1[c] - m_ : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b = B[(ks * NT + nt) * 32];
                dmma_m8n8k4(c0[nt], e0[nt], a0, b);
                dmma_m8n8k4(c1[nt], e1[nt], a1, b);
            }
        }
        // accumulator layout: row = grp, cols = nt*8 + tig*2 + {0,1}
        double q0a = 0.0, q0b = 0.0, q1a = 0.0, q1b = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 8 + tig * 2;
            const bool i0 = col < d, i1 = col + 1 < d;
            const double m0 = mus[col], m1 = mus[col + 1];
            q0a = fma(c0[nt], (ok0 && i0) ? row0[col] - m0 : 0.0, q0a);
            q0b = fma(e0[nt], (ok0 && i1) ? row0[col + 1] - m1 : 0.0, q0b);
            q1a = fma(c1[nt], (ok1 && i0) ? row1[col] - m0 : 0.0, q1a);
            q1b = fma(e1[nt], (ok1 && i1) ? row1[col + 1] - m1 : 0.0, q1b);
        }
        double q0 = q0a + q0b, q1 = q1a + q1b;
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1);
        q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        if (tig == 0 && ok0) out[r0 + grp] = -0.5 * q0 + log_norm;
        if (tig == 0 && ok1) out[r0 + 8 + grp] = -0.5 * q1 + log_norm;
        __syncwarp();                                          // everybody is done with this buffer
        buf ^= 1;
    }
    cp_async_wait<0>();
}

int dense_gaussian_eval_launch(const gl_model *m, const double *x, long long n, double *out, cudaStream_t st) {
    const int d = m->dev.d;
    const int KT = (d + 3) / 4, NT = (d + 7) / 8;
    const int dp = KT * 4 > NT * 8 ? KT * 4 : NT * 8;
    const size_t p_bytes = sizeof(double) * (size_t)KT * NT * 32;
    // pipelined form: two raw 16-row tiles per warp; as many warps per block (<= 8) as let two blocks share an SM
    {
        const size_t fixed = p_bytes + sizeof(double) * (dp + 16), per_warp = sizeof(double) * 2 * 16 * (size_t)d;
        int warps = DG_WARPS;
        while (warps > 2 && fixed + warps * per_warp > 112 * 1024) --warps;
        const size_t smem = fixed + warps * per_warp;
        const bool aligned = ((uintptr_t)x & 15) == 0;
        static const bool off = getenv("GLEIPNIR_DMMA_NO_PIPELINE") != nullptr;
        if (!off && aligned && smem <= 112 * 1024 && n >= 16 && NT <= 8) {
            long long blocks = ((n + 15) / 16 + warps - 1) / warps;
            const long long cap = (long long)sm_count() * 2;
            if (blocks > cap) blocks = cap;
#define GL_DG_PIPE(NTM)                                                                                          \
    do {                                                                                                         \
        GL_CUDA(cudaFuncSetAttribute(dense_gaussian_dmma_pipe_kernel<NTM>,                                       \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        dense_gaussian_dmma_pipe_kernel<NTM><<<(int)blocks, warps * 32, smem, st>>>(                          \
            x, n, d, KT, m->dev.aux0, m->dev.aux2, m->dev.s0, out);                                              \
    } while (0)
            switch (NT) {
                case 1: GL_DG_PIPE(1); break;
                case 2: GL_DG_PIPE(2); break;
                case 3: GL_DG_PIPE(3); break;
                case 4: GL_DG_PIPE(4); break;
                case 5: GL_DG_PIPE(5); break;
                case 6: GL_DG_PIPE(6); break;
                case 7: GL_DG_PIPE(7); break;
                default: GL_DG_PIPE(8); break;
            }
#undef GL_DG_PIPE
            GL_CHECK_LAUNCH();
            return GL_OK;
        }
    }
    int warps = DG_WARPS;                                      // fewer warps per block when P itself fills shared memory
    while (warps > 1 && sizeof(double) * ((size_t)KT * NT * 32 + (size_t)warps * 16 * dp) > 200 * 1024) warps >>= 1;
    const int smem = (int)(sizeof(double) * ((size_t)KT * NT * 32 + (size_t)warps * 16 * dp));
    GL_REQUIRE(smem <= 227 * 1024, "dense Gaussian: d=%d does not fit the tensor-core kernel's shared memory", d);
    long long blocks = ((n + 15) / 16 + warps - 1) / warps;
    const long long cap = (long long)sm_count() * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define GL_DG_LAUNCH(KTM)                                                                                        \
    do {                                                                                                         \
        if (smem > 48 * 1024)                                                                                    \
            GL_CUDA(cudaFuncSetAttribute(dense_gaussian_dmma_kernel<KTM>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
        dense_gaussian_dmma_kernel<KTM><<<(int)blocks, warps * 32, smem, st>>>(                               \
            x, n, d, KT, NT, m->dev.aux0, m->dev.aux2, m->dev.s0, out);                                          \
    } while (0)
    if (KT <= 4) GL_DG_LAUNCH(4);
    else if (KT <= 8) GL_DG_LAUNCH(8);
    else if (KT <= 16) GL_DG_LAUNCH(16);
    else GL_DG_LAUNCH(32);
#undef GL_DG_LAUNCH
    GL_CHECK_LAUNCH();
    return GL_OK;
}

int dense_gaussian_pack(const double *P_dev, int d, double *Pf_dev) {
    const int KT = (d + 3) / 4, NT = (d + 7) / 8;
    dense_gaussian_pack_kernel<<<(KT * NT * 32 + 255) / 256, 256>>>(P_dev, d, KT, NT, Pf_dev);
    GL_CHECK_LAUNCH();
    GL_CUDA(cudaDeviceSynchronize());
    return GL_OK;
}

// FP64 tensor-core peak: 4 independent accumulator pairs per warp, back-to-back DMMAs
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double a, double b) {
    double c[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int i = 0; i < iters; ++i) {
        dmma_m8n8k4(c[0], c[1], a, b);
        dmma_m8n8k4(c[2], c[3], a, b);
        dmma_m8n8k4(c[4], c[5], a, b);
        dmma_m8n8k4(c[6], c[7], a, b);
    }
    const double s = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace gl

extern "C" {

// FP64 tensor-core (DMMA m8n8k4) peak microbenchmark; flops = blocks*8 warps*iters*4*512.
int gl_microbench_dmma(int32_t blocks, int32_t iters, double *scratch_dev, int64_t *flops_out, void *stream) {
    GL_REQUIRE(blocks >= 1 && iters >= 1 && scratch_dev != nullptr, "bad argument");
    gl::dmma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch_dev, iters, 1.0000001, 0.9999999);
    GL_CHECK_LAUNCH();
    if (flops_out) *flops_out = (int64_t)blocks * 8 * (int64_t)iters * 4 * 512;
    return GL_OK;
}

}  // extern "C"
