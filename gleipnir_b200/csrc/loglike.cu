// Model handles + the batched log-likelihood evaluation kernel
// (replaces the N scalar `loglikelihood(row)` calls of nested_sampling.py:108).
#include "common.cuh"
#include "likelihoods.cuh"
#include "loglike.cuh"

namespace gl {

int dense_gaussian_eval_launch(const gl_model *m, const double *x, long long n, double *out, cudaStream_t st);
int dense_gaussian_pack(const double *P_dev, int d, double *Pf_dev);

struct EvalLauncher {
    const gl_model *m;
    const double *x;
    long long n;
    double *out;
    cudaStream_t st;
    template <class Like>
    int operator()() {
        return loglike_rows_launch<Like>(m->dev, x, n, out, st);
    }
};

int loglike_eval_launch(const gl_model *m, const double *x_dev, long long n, double *logl_dev, cudaStream_t st) {
    GL_REQUIRE(m != nullptr, "model is NULL");
    if (n <= 0) return GL_OK;
    GL_REQUIRE(x_dev != nullptr && logl_dev != nullptr, "NULL device pointer");
    if (m->dev.kind == GL_MODEL_EXTERNAL) {
        GL_REQUIRE(m->dev.ext_eval != nullptr, "external model without an eval entry point");
        return m->dev.ext_eval(&m->dev, x_dev, n, logl_dev, (void *)st);
    }
    if (m->dev.kind == GL_MODEL_DENSE_GAUSSIAN && m->dev.aux2 != nullptr && !m->no_tensor)
        return dense_gaussian_eval_launch(m, x_dev, n, logl_dev, st);     // FP64 tensor cores (DMMA)
    EvalLauncher L{m, x_dev, n, logl_dev, st};
    return dispatch_model(m->dev.kind, L);
}

static int new_model(int kind, int d, gl_model **out) {
    GL_REQUIRE(out != nullptr, "out is NULL");
    GL_REQUIRE(d >= 1 && d <= GL_MAX_DIM, "d=%d outside [1, %d]", d, GL_MAX_DIM);
    gl_model *m = new gl_model();
    memset(m, 0, sizeof(*m));
    m->dev.kind = kind;
    m->dev.d = d;
    *out = m;
    return GL_OK;
}

}  // namespace gl

extern "C" {

int gl_model_create_iid_gaussian(int32_t d, double mu, double sigma, gl_model **out) {
    GL_REQUIRE(sigma > 0.0, "sigma must be > 0");
    int rc = gl::new_model(GL_MODEL_IID_GAUSSIAN, d, out);
    if (rc) return rc;
    (*out)->dev.s0 = mu;
    (*out)->dev.s1 = sigma;
    (*out)->dev.s2 = (double)d * (-GL_HALF_LOG_2PI - log(sigma));
    return GL_OK;
}

int gl_model_create_eggcarton(int32_t d, gl_model **out) { return gl::new_model(GL_MODEL_EGGCARTON, d, out); }

int gl_model_create_mixture1d(const double *positions_host, int32_t n_positions, double width, gl_model **out) {
    GL_REQUIRE(positions_host != nullptr && n_positions >= 1, "need >= 1 position");
    GL_REQUIRE(width > 0.0, "width must be > 0");
    int rc = gl::new_model(GL_MODEL_MIXTURE1D, 1, out);
    if (rc) return rc;
    gl_model *m = *out;
    cudaError_t e = cudaMalloc(&m->aux_alloc, sizeof(double) * n_positions);
    if (e == cudaSuccess)
        e = cudaMemcpy(m->aux_alloc, positions_host, sizeof(double) * n_positions, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(m->aux_alloc);
        delete m;
        *out = nullptr;
        return gl::cuda_fail(e, "mixture1d upload", __FILE__, __LINE__);
    }
    m->dev.s0 = width;
    m->dev.s1 = 1.0 / sqrt(2.0 * M_PI * width * width);
    m->dev.n_aux = n_positions;
    m->dev.aux0 = m->aux_alloc;
    return GL_OK;
}

int gl_model_create_dense_gaussian(int32_t d, const double *mu_host, const double *P_host, double log_norm,
                                   gl_model **out) {
    GL_REQUIRE(mu_host != nullptr && P_host != nullptr, "mu/P is NULL");
    int rc = gl::new_model(GL_MODEL_DENSE_GAUSSIAN, d, out);
    if (rc) return rc;
    gl_model *m = *out;
    const size_t nd_ = (size_t)d;
    const size_t KT = (nd_ + 3) / 4, NT = (nd_ + 7) / 8;
    const size_t nd = (size_t)d, bytes = sizeof(double) * (nd + nd * nd + KT * NT * 32);
    cudaError_t e = cudaMalloc(&m->aux_alloc, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(m->aux_alloc, mu_host, sizeof(double) * nd, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(m->aux_alloc + nd, P_host, sizeof(double) * nd * nd, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(m->aux_alloc);
        delete m;
        *out = nullptr;
        return gl::cuda_fail(e, "dense gaussian upload", __FILE__, __LINE__);
    }
    m->dev.s0 = log_norm;
    m->dev.aux0 = m->aux_alloc;
    m->dev.aux1 = m->aux_alloc + nd;
    m->dev.aux2 = m->aux_alloc + nd + nd * nd;
    rc = gl::dense_gaussian_pack(m->dev.aux1, d, m->aux_alloc + nd + nd * nd);
    if (rc) {
        cudaFree(m->aux_alloc);
        delete m;
        *out = nullptr;
        return rc;
    }
    return GL_OK;
}

int gl_model_set_tensor_path(gl_model *m, int32_t enable) {
    GL_REQUIRE(m != nullptr, "model is NULL");
    m->no_tensor = enable == 0;
    return GL_OK;
}

int gl_model_create_external(int32_t d, void *walk_fn, void *eval_fn, const void *blob_host, int64_t blob_bytes,
                             gl_model **out) {
    GL_REQUIRE(walk_fn != nullptr && eval_fn != nullptr, "plugin entry points are NULL");
    GL_REQUIRE(blob_bytes >= 0 && (blob_bytes == 0 || blob_host != nullptr), "bad blob");
    int rc = gl::new_model(GL_MODEL_EXTERNAL, d, out);
    if (rc) return rc;
    gl_model *m = *out;
    if (blob_bytes > 0) {
        cudaError_t e = cudaMalloc(&m->aux_alloc, (size_t)blob_bytes);
        if (e == cudaSuccess) e = cudaMemcpy(m->aux_alloc, blob_host, (size_t)blob_bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(m->aux_alloc);
            delete m;
            *out = nullptr;
            return gl::cuda_fail(e, "external model blob upload", __FILE__, __LINE__);
        }
    }
    m->dev.aux0 = m->aux_alloc;
    m->dev.n_aux = (int32_t)(blob_bytes / 8);
    // host-visible launch hints: header words 8..10 of the blob (ode_coop.cuh: number of conditions, ...)
    if (blob_bytes >= 11 * 8) {
        const double *hb = (const double *)blob_host;
        m->dev.s0 = hb[8];
        m->dev.s1 = hb[9];
        m->dev.s2 = hb[10];
    }
    m->dev.ext_walk = (int (*)(const void *, const void *, void *))walk_fn;
    m->dev.ext_eval = (int (*)(const void *, const double *, long long, double *, void *))eval_fn;
    return GL_OK;
}

int gl_model_dim(const gl_model *m) { return m ? m->dev.d : GL_EINVAL; }
int gl_model_kind_of(const gl_model *m) { return m ? m->dev.kind : GL_EINVAL; }

int gl_model_destroy(gl_model *m) {
    if (!m) return GL_OK;
    if (m->aux_alloc) cudaFree(m->aux_alloc);
    delete m;
    return GL_OK;
}

int gl_loglike_eval(const gl_model *m, const double *x_dev, int64_t n, double *logl_dev, void *stream) {
    return gl::loglike_eval_launch(m, x_dev, n, logl_dev, (cudaStream_t)stream);
}

int gl_loglike_eval_host(const gl_model *m, const double *x_host, int64_t n, double *logl_host) {
    GL_REQUIRE(m != nullptr, "model is NULL");
    if (n <= 0) return GL_OK;
    GL_REQUIRE(x_host != nullptr && logl_host != nullptr, "NULL host pointer");
    const size_t d = (size_t)m->dev.d;
    double *buf = nullptr;
    GL_CUDA(cudaMalloc(&buf, sizeof(double) * (n * d + n)));
    cudaError_t e = cudaMemcpy(buf, x_host, sizeof(double) * n * d, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? gl::loglike_eval_launch(m, buf, n, buf + n * d, 0)
                              : gl::cuda_fail(e, "H2D", __FILE__, __LINE__);
    if (rc == GL_OK) {
        e = cudaMemcpy(logl_host, buf + n * d, sizeof(double) * n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = gl::cuda_fail(e, "D2H", __FILE__, __LINE__);
    }
    cudaFree(buf);
    return rc;
}

}  // extern "C"
