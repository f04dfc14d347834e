// Batched log-likelihood evaluation kernel template (replaces the N scalar
// `loglikelihood(row)` calls of nested_sampling.py:108); header-only so that plugin
// libraries instantiate it for their own functors.
#pragma once
#include "common.cuh"
#include "likelihoods.cuh"

namespace gl {

// One thread per parameter vector; rows are read straight from global memory (for the
// small-d analytic models a row is 16-40 B, the whole array streams through L1 once).
template <class Like>
__global__ void __launch_bounds__(256) loglike_rows_kernel(ModelDev M, const double *__restrict__ x, long long n,
                                                           double *__restrict__ out) {
    const Like like(M);
    const int d = M.d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = like.eval(x + i * d, 1);
}

template <class Like>
inline int loglike_rows_launch(const ModelDev &M, const double *x, long long n, double *out, cudaStream_t st,
                               int threads = 256) {
    long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    loglike_rows_kernel<Like><<<(int)blocks, threads, 0, st>>>(M, x, n, out);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

}  // namespace gl
