// Prior draw / ppf / pdf / logpdf kernels: SampledParameter as device transforms
// (gleipnir/sampled_parameter.py:40-80; replaces the N*d scalar rvs(1) calls of
// nested_sampling.py:93-101).  HBM-bound elementwise kernels: 8 B/element written by
// the draw, 16 B/element moved by ppf/pdf.  Grid-stride, grid = multiple of SM count.
#include "common.cuh"
#include "philox.cuh"
#include "priors.cuh"

namespace gl {

struct PriorParams {   // passed by value (<= 4 KB kernel parameter space)
    int32_t kind[GL_MAX_DIM];
    double p0[GL_MAX_DIM];
    double p1[GL_MAX_DIM];
};

// x[r, j] = ppf_j(u(seed; row, j)).  One thread per element; consecutive threads write
// consecutive addresses (row-major), so stores coalesce for any d.
__global__ void __launch_bounds__(256) prior_draw_kernel(PriorParams P, int d, uint64_t seed, uint32_t tag,
                                                         long long first_row, long long n_rows,
                                                         double *__restrict__ x) {
    const long long total = n_rows * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / d;
        const int j = (int)(e - r * d);
        const unsigned long long g = (unsigned long long)(first_row + r);
        U2 u = philox_uniforms(seed, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)j, tag);
        x[e] = prior_ppf(P.kind[j], P.p0[j], P.p1[j], u.a);
    }
}

enum { OP_PPF = 0, OP_PDF = 1, OP_LOGPDF = 2 };

template <int OP>
__global__ void __launch_bounds__(256) prior_map_kernel(int kind, double p0, double p1,
                                                        const double *__restrict__ in, double *__restrict__ out,
                                                        long long n) {
    // two elements per thread per trip through 16-byte accesses when aligned
    const long long n2 = n >> 1;
    const bool aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (aligned) {
        const double2 *in2 = reinterpret_cast<const double2 *>(in);
        double2 *out2 = reinterpret_cast<double2 *>(out);
        for (long long i = tid; i < n2; i += stride) {
            double2 v = __ldg(in2 + i), r;
            if (OP == OP_PPF) { r.x = prior_ppf(kind, p0, p1, v.x); r.y = prior_ppf(kind, p0, p1, v.y); }
            else if (OP == OP_PDF) { r.x = prior_pdf(kind, p0, p1, v.x); r.y = prior_pdf(kind, p0, p1, v.y); }
            else { r.x = log(prior_pdf(kind, p0, p1, v.x)); r.y = log(prior_pdf(kind, p0, p1, v.y)); }
            out2[i] = r;
        }
        if (tid == 0 && (n & 1)) {
            double v = in[n - 1];
            out[n - 1] = OP == OP_PPF ? prior_ppf(kind, p0, p1, v)
                         : OP == OP_PDF ? prior_pdf(kind, p0, p1, v) : log(prior_pdf(kind, p0, p1, v));
        }
    } else {
        for (long long i = tid; i < n; i += stride) {
            double v = in[i];
            out[i] = OP == OP_PPF ? prior_ppf(kind, p0, p1, v)
                     : OP == OP_PDF ? prior_pdf(kind, p0, p1, v) : log(prior_pdf(kind, p0, p1, v));
        }
    }
}

__global__ void __launch_bounds__(256) prior_transform_kernel(PriorParams P, int d, const double *__restrict__ u,
                                                              double *__restrict__ x, long long n_rows) {
    const long long total = n_rows * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e % d);
        x[e] = prior_ppf(P.kind[j], P.p0[j], P.p1[j], u[e]);
    }
}

static int fill_params(const gl_prior *priors, int d, PriorParams &P) {
    GL_REQUIRE(priors != nullptr, "priors is NULL");
    GL_REQUIRE(d >= 1 && d <= GL_MAX_DIM, "d=%d outside [1, %d]", d, GL_MAX_DIM);
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < d; ++j) {
        GL_REQUIRE(priors[j].kind >= GL_PRIOR_UNIFORM && priors[j].kind <= GL_PRIOR_LOGUNIFORM,
                   "prior %d: unknown kind %d", j, priors[j].kind);
        if (priors[j].kind == GL_PRIOR_LOGUNIFORM)
            GL_REQUIRE(priors[j].p0 > 0.0 && priors[j].p1 > priors[j].p0, "prior %d: loguniform needs 0 < a < b", j);
        else
            GL_REQUIRE(priors[j].p1 > 0.0, "prior %d: scale must be > 0", j);
        P.kind[j] = priors[j].kind;
        P.p0[j] = priors[j].p0;
        P.p1[j] = priors[j].p1;
    }
    return GL_OK;
}

static int grid_for(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)sm_count() * 8;   // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int prior_draw_launch(const gl_prior *priors, int d, uint64_t seed, uint32_t tag, long long first_row,
                      long long n_rows, double *x_dev, cudaStream_t st) {
    PriorParams P;
    int rc = fill_params(priors, d, P);
    if (rc) return rc;
    if (n_rows <= 0) return GL_OK;
    GL_REQUIRE(x_dev != nullptr, "x_dev is NULL");
    prior_draw_kernel<<<grid_for(n_rows * d, 256), 256, 0, st>>>(P, d, seed, tag, first_row, n_rows, x_dev);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

template <int OP>
static int map_launch(const gl_prior *prior, const double *in, double *out, long long n, cudaStream_t st) {
    PriorParams P;
    int rc = fill_params(prior, 1, P);
    if (rc) return rc;
    if (n <= 0) return GL_OK;
    GL_REQUIRE(in != nullptr && out != nullptr, "NULL device pointer");
    prior_map_kernel<OP><<<grid_for((n + 1) / 2, 256), 256, 0, st>>>(P.kind[0], P.p0[0], P.p1[0], in, out, n);
    GL_CHECK_LAUNCH();
    return GL_OK;
}

template <int OP>
static int map_host(const gl_prior *prior, const double *in_h, double *out_h, long long n) {
    if (n <= 0) return GL_OK;
    GL_REQUIRE(in_h != nullptr && out_h != nullptr, "NULL host pointer");
    double *buf = nullptr;
    GL_CUDA(cudaMalloc(&buf, sizeof(double) * n * 2));
    cudaError_t e = cudaMemcpy(buf, in_h, sizeof(double) * n, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? map_launch<OP>(prior, buf, buf + n, n, 0) : cuda_fail(e, "H2D", __FILE__, __LINE__);
    if (rc == GL_OK) {
        e = cudaMemcpy(out_h, buf + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H", __FILE__, __LINE__);
    }
    cudaFree(buf);
    return rc;
}

}  // namespace gl

extern "C" {

int gl_prior_draw(const gl_prior *priors_host, int32_t d, uint64_t seed, uint32_t tag, int64_t first_row,
                  int64_t n_rows, double *x_dev, void *stream) {
    return gl::prior_draw_launch(priors_host, d, seed, tag, first_row, n_rows, x_dev, (cudaStream_t)stream);
}
int gl_prior_ppf(const gl_prior *p, const double *u_dev, double *out_dev, int64_t n, void *stream) {
    return gl::map_launch<gl::OP_PPF>(p, u_dev, out_dev, n, (cudaStream_t)stream);
}
int gl_prior_pdf(const gl_prior *p, const double *x_dev, double *out_dev, int64_t n, void *stream) {
    return gl::map_launch<gl::OP_PDF>(p, x_dev, out_dev, n, (cudaStream_t)stream);
}
int gl_prior_logpdf(const gl_prior *p, const double *x_dev, double *out_dev, int64_t n, void *stream) {
    return gl::map_launch<gl::OP_LOGPDF>(p, x_dev, out_dev, n, (cudaStream_t)stream);
}
int gl_prior_transform(const gl_prior *priors_host, int32_t d, const double *u_dev, double *x_dev, int64_t n_rows,
                       void *stream) {
    gl::PriorParams P;
    int rc = gl::fill_params(priors_host, d, P);
    if (rc) return rc;
    if (n_rows <= 0) return GL_OK;
    GL_REQUIRE(u_dev != nullptr && x_dev != nullptr, "NULL device pointer");
    gl::prior_transform_kernel<<<gl::grid_for(n_rows * d, 256), 256, 0, (cudaStream_t)stream>>>(P, d, u_dev, x_dev,
                                                                                               n_rows);
    GL_CHECK_LAUNCH();
    return GL_OK;
}
int gl_prior_ppf_host(const gl_prior *p, const double *u, double *out, int64_t n) {
    return gl::map_host<gl::OP_PPF>(p, u, out, n);
}
int gl_prior_pdf_host(const gl_prior *p, const double *x, double *out, int64_t n) {
    return gl::map_host<gl::OP_PDF>(p, x, out, n);
}
int gl_prior_logpdf_host(const gl_prior *p, const double *x, double *out, int64_t n) {
    return gl::map_host<gl::OP_LOGPDF>(p, x, out, n);
}

}  // extern "C"
