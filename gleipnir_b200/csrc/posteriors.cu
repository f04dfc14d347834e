// posteriors(): weighted marginal histograms of the dead points (nested_sampling.py:266-293)
//   norm_w_i = w_i * exp(L_i) / Z  (kept as exp(log w + L - logZ): no overflow);  mask norm_w > 0;
//   nbins = 2*int(cbrt(#masked)) (Rice rule, done by the caller from the count);
//   per parameter: np.histogram(x[mask], weights=norm_w[mask], density=True, bins=nbins).
// The bin assignment follows numpy's uniform-bin path exactly (scaled index, then the two edge
// corrections against edges = linspace(min, max, nbins+1)).  Sums are deterministic: each warp
// adds a contiguous slice of rows in row order, partial histograms are then added in warp order.
#include "common.cuh"

namespace gl {

struct PosteriorStats {                 // device
    unsigned long long n_masked;
    unsigned long long min_key[GL_MAX_DIM], max_key[GL_MAX_DIM];     // sort_key() of the extremes
};

__global__ void __launch_bounds__(256) posterior_weights_kernel(const double *__restrict__ dead_logl,
                                                                const double *__restrict__ dead_w,
                                                                const double *__restrict__ dead_x, long long rows,
                                                                int d, double log_z, double *__restrict__ nw,
                                                                PosteriorStats *st) {
    unsigned long long cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x) {
        double v = exp(log(dead_w[i]) + dead_logl[i] - log_z);
        if (v != v) v = 0.0;
        nw[i] = v;
        if (v > 0.0) {
            ++cnt;
            for (int j = 0; j < d; ++j) {
                const unsigned long long key = sort_key(dead_x[i * d + j]);
                atomicMin(&st->min_key[j], key);
                atomicMax(&st->max_key[j], key);
            }
        }
    }
    cnt = (unsigned long long)warp_sum_i((int)cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&st->n_masked, cnt);
}

__device__ __forceinline__ int numpy_bin(double x, double first, double last, double step, int nbins) {
    int idx = (int)(((x - first) / (last - first)) * (double)nbins);          // numpy: divide, then scale
    if (idx == nbins) idx -= 1;
    // edges[i] = i*step + first (np.linspace), edges[nbins] = last
    const double e_lo = idx == nbins ? last : __dadd_rn(__dmul_rn((double)idx, step), first);
    if (x < e_lo) idx -= 1;
    const double e_hi = idx + 1 == nbins ? last : __dadd_rn(__dmul_rn((double)(idx + 1), step), first);
    if (x >= e_hi && idx != nbins - 1) idx += 1;
    return idx;
}

constexpr int POST_WARPS = 8;

// grid = n_slices blocks... each WARP owns rows [slice*chunk, ...); partial[(slice*d + j)*nbins + b]
__global__ void __launch_bounds__(POST_WARPS * 32) posterior_hist_kernel(const double *__restrict__ dead_x,
                                                                         const double *__restrict__ nw,
                                                                         long long rows, int d, int nbins,
                                                                         const PosteriorStats *st, long long chunk,
                                                                         double *__restrict__ partial) {
    extern __shared__ double sh[];                 // [POST_WARPS][nbins]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long slice = (long long)blockIdx.x * POST_WARPS + warp;
    const long long lo = slice * chunk;
    long long hi = lo + chunk;
    if (hi > rows) hi = rows;
    double *h = sh + (size_t)warp * nbins;
    for (int j = 0; j < d; ++j) {
        double first = key_to_double(st->min_key[j]), last = key_to_double(st->max_key[j]);
        if (first == last) { first -= 0.5; last += 0.5; }      // numpy widens a degenerate range
        const double step = (last - first) / (double)nbins;
        for (int b = lane; b < nbins; b += 32) h[b] = 0.0;
        __syncwarp();
        for (long long r0 = lo; r0 < hi; r0 += 32) {
            const long long i = r0 + lane;
            double wv = 0.0;
            int idx = 0;
            if (i < hi) {
                wv = nw[i];
                if (wv > 0.0) idx = numpy_bin(dead_x[i * d + j], first, last, step, nbins);
            }
            // rows are added in row order: lane l's contribution goes in at turn l
            for (int l = 0; l < 32; ++l) {
                const double wl = __shfl_sync(0xffffffffu, wv, l);
                const int il = __shfl_sync(0xffffffffu, idx, l);
                if (lane == 0 && wl > 0.0) h[il] += wl;
            }
        }
        __syncwarp();
        if (lo < hi)
            for (int b = lane; b < nbins; b += 32) partial[((size_t)slice * d + j) * nbins + b] = h[b];
        __syncwarp();
    }
}

// marginal[j][b] = (sum over slices, in slice order) / (bin width) / (total weight); edges/centers as numpy
__global__ void posterior_finish_kernel(const double *__restrict__ partial, int n_slices, int d, int nbins,
                                        const PosteriorStats *st, double *__restrict__ marginal,
                                        double *__restrict__ edges, double *__restrict__ centers) {
    const int j = blockIdx.x;
    extern __shared__ double tot[];                // [nbins]
    double first = key_to_double(st->min_key[j]), last = key_to_double(st->max_key[j]);
    if (first == last) { first -= 0.5; last += 0.5; }          // numpy widens a degenerate range
    const double step = (last - first) / (double)nbins;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        double s = 0.0;
        for (int sl = 0; sl < n_slices; ++sl) s += partial[((size_t)sl * d + j) * nbins + b];
        tot[b] = s;
    }
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x)
        edges[(size_t)j * (nbins + 1) + b] = b == nbins ? last : __dadd_rn(__dmul_rn((double)b, step), first);
    __syncthreads();
    __shared__ double total;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < nbins; ++b) s += tot[b];
        total = s;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        const double e0 = edges[(size_t)j * (nbins + 1) + b], e1 = edges[(size_t)j * (nbins + 1) + b + 1];
        marginal[(size_t)j * nbins + b] = tot[b] / (e1 - e0) / total;
        centers[(size_t)j * nbins + b] = (e0 + e1) / 2.0;
    }
}

}  // namespace gl

extern "C" {

int64_t gl_posterior_stats_bytes(void) { return (int64_t)sizeof(gl::PosteriorStats); }

// Stage 1: normalised weights nw_dev[rows] and the mask count / per-parameter range in stats_dev.
int gl_posterior_weights(const double *dead_logl_dev, const double *dead_w_dev, const double *dead_x_dev, int64_t rows,
                         int32_t d, double log_z, double *nw_dev, void *stats_dev, int64_t *n_masked_host,
                         void *stream) {
    GL_REQUIRE(dead_logl_dev && dead_w_dev && dead_x_dev && nw_dev && stats_dev && rows >= 1 && d >= 1 &&
                   d <= GL_MAX_DIM, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    gl::PosteriorStats h;
    memset(&h, 0, sizeof(h));
    for (int j = 0; j < GL_MAX_DIM; ++j) h.min_key[j] = ~0ull;
    GL_CUDA(cudaMemcpyAsync(stats_dev, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    GL_CUDA(cudaStreamSynchronize(st));
    long long blocks = (rows + 255) / 256;
    const long long cap = (long long)gl::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    gl::posterior_weights_kernel<<<(int)blocks, 256, 0, st>>>(dead_logl_dev, dead_w_dev, dead_x_dev, rows, d, log_z,
                                                              nw_dev, (gl::PosteriorStats *)stats_dev);
    GL_CHECK_LAUNCH();
    if (n_masked_host) {
        GL_CUDA(cudaMemcpyAsync(&h, stats_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
        GL_CUDA(cudaStreamSynchronize(st));
        *n_masked_host = (int64_t)h.n_masked;
    }
    return GL_OK;
}

int64_t gl_posterior_partial_doubles(int64_t rows, int32_t d, int32_t nbins) {
    long long slices = (rows + 4095) / 4096;
    if (slices > 512) slices = 512;
    if (slices < 1) slices = 1;
    slices = (slices + gl::POST_WARPS - 1) / gl::POST_WARPS * gl::POST_WARPS;
    return slices * (int64_t)d * nbins;
}

// Stage 2: marginal[d, nbins], edges[d, nbins+1], centers[d, nbins] (np.histogram(..., density=True)).
int gl_posterior_histograms(const double *dead_x_dev, const double *nw_dev, int64_t rows, int32_t d, int32_t nbins,
                            const void *stats_dev, double *partial_dev, double *marginal_dev, double *edges_dev,
                            double *centers_dev, void *stream) {
    GL_REQUIRE(dead_x_dev && nw_dev && stats_dev && partial_dev && marginal_dev && edges_dev && centers_dev,
               "NULL argument");
    GL_REQUIRE(nbins >= 1 && nbins <= 4096, "nbins outside [1, 4096]");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_slices = gl_posterior_partial_doubles(rows, d, nbins) / ((long long)d * nbins);
    const long long chunk = (rows + n_slices - 1) / n_slices;
    GL_CUDA(cudaMemsetAsync(partial_dev, 0, sizeof(double) * n_slices * d * nbins, st));
    const int smem = (int)(sizeof(double) * gl::POST_WARPS * nbins);
    if (smem > 48 * 1024)
        GL_CUDA(cudaFuncSetAttribute(gl::posterior_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gl::posterior_hist_kernel<<<(int)(n_slices / gl::POST_WARPS), gl::POST_WARPS * 32, smem, st>>>(
        dead_x_dev, nw_dev, rows, d, nbins, (const gl::PosteriorStats *)stats_dev, chunk, partial_dev);
    gl::posterior_finish_kernel<<<d, 256, sizeof(double) * nbins, st>>>(
        partial_dev, (int)n_slices, d, nbins, (const gl::PosteriorStats *)stats_dev, marginal_dev, edges_dev,
        centers_dev);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

}  // extern "C"
