// Roofline denominators that MEASURED_PEAKS.json does not carry: FP64 DFMA pipe peak
// ("the builder must measure DFMA ... peak in the same run").
#include "common.cuh"

namespace gl {

// 8 independent DFMA chains per thread; 2 flop per DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains live
}

// dependent-issue latency of DFMA: one warp, a chain of `iters` x 8 dependent fmas, clock64 around it
__global__ void dfma_latency_kernel(double *out, long long *cycles, int iters, double a, double b) {
    double x = out[threadIdx.x];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x = fma(x, a, b);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

}  // namespace gl

extern "C" {

// Cycles per dependent DFMA seen by a single warp (host-synchronous; scratch_dev: >= 32 doubles + 1 int64).
int gl_microbench_dfma_latency(int32_t iters, double *scratch_dev, double *cycles_per_dfma_out) {
    GL_REQUIRE(iters >= 1 && scratch_dev != nullptr && cycles_per_dfma_out != nullptr, "bad argument");
    long long *cyc = reinterpret_cast<long long *>(scratch_dev + 32);
    GL_CUDA(cudaMemset(scratch_dev, 0, 33 * sizeof(double)));
    gl::dfma_latency_kernel<<<1, 32>>>(scratch_dev, cyc, iters, 0.999999, 1e-9);
    GL_CHECK_LAUNCH();
    long long h = 0;
    GL_CUDA(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    *cycles_per_dfma_out = (double)h / ((double)iters * 8.0);
    return GL_OK;
}

// Enqueue the DFMA kernel; flops = blocks*256*iters*8*2.  The caller times it with events.
int gl_microbench_dfma(int32_t blocks, int32_t iters, double *scratch_dev, int64_t *flops_out, void *stream) {
    GL_REQUIRE(blocks >= 1 && iters >= 1 && scratch_dev != nullptr, "bad argument");
    gl::dfma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch_dev, iters, 0.999999, 1e-9);
    GL_CHECK_LAUNCH();
    if (flops_out) *flops_out = (int64_t)blocks * 256 * (int64_t)iters * 16;
    return GL_OK;
}

}  // extern "C"
