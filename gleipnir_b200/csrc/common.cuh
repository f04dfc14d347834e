// Shared helpers for the gleipnir_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdarg.h>

#include "../../include/gleipnir_b200.h"

namespace gl {

// ---- error plumbing (thread-local last-error string behind gl_last_error) --------
// Header-only so that code-generated plugin libraries (ODE networks) are self-contained; each
// shared object keeps its own buffer (the core library reads a plugin's through its
// gl_plugin_last_error entry point).
inline char *error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}
inline void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
}
inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return GL_ECUDA;
}

#define GL_CUDA(call)                                                       \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return gl::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define GL_CHECK_LAUNCH()                                                   \
    do {                                                                    \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return gl::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define GL_REQUIRE(cond, ...)                                               \
    do {                                                                    \
        if (!(cond)) {                                                      \
            gl::set_error(__VA_ARGS__);                                     \
            return GL_EINVAL;                                               \
        }                                                                   \
    } while (0)

inline int sm_count() {   // of the current device, cached per device; 148 on B200
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cached[dev & 63] > 0) return cached[dev & 63];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev & 63] = n;
    return n;
}

// ---- ordering key: np.argmin order (NaN first, -0.0 == +0.0), ascending ----------
__host__ __device__ __forceinline__ unsigned long long sort_key(double x) {
    if (x != x) return 0ull;
    x += 0.0;   // -0.0 -> +0.0
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b;
    memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__host__ __device__ __forceinline__ double key_to_double(unsigned long long k) {
    unsigned long long b;
    if (k == 0ull) b = 0x7ff8000000000000ull;   // NaN
    else b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

// ---- warp / block reductions ------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree); result valid in every thread.
// `scratch` must hold 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < nw ? scratch[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A kernel launched with launch_pdl() may start while its predecessor in the stream is still
// running; it must call pdl_wait() before touching anything the predecessor (or anything still in
// flight before it) writes.  Protocol used throughout: pdl_wait() THEN pdl_trigger(), so a kernel
// only lets its successor start once its own predecessor has completed -- the successor's
// pre-wait code may therefore read data written two or more launches back, never newer.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace gl
