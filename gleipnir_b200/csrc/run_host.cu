// gl_ns_run_host: NestedSampling(...).run() (nested_sampling.py:83-213) through HOST
// buffers -- the end-to-end entry point a non-Python caller binds.  Allocates the device
// buffers, runs init/advance/finalize on a private stream, downloads the dead table and
// returns it in death order.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

struct DeviceSlab {
    void *p = nullptr;
    ~DeviceSlab() {
        if (p) cudaFree(p);
    }
};

inline size_t a256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" int gl_ns_run_host(const gl_ns_config *cfg, const gl_prior *priors_host, int32_t d, const gl_model *model,
                              int64_t n_deaths_total, double *dead_host, gl_ns_state *state_out) {
    GL_REQUIRE(cfg && priors_host && model && dead_host && state_out, "NULL argument");
    GL_REQUIRE(n_deaths_total >= 1, "n_deaths_total must be >= 1");
    const int64_t ws_bytes = gl_ns_workspace_bytes(cfg, d);
    if (ws_bytes < 0) return (int)ws_bytes;
    const size_t N = (size_t)cfg->population_size, cap = (size_t)cfg->dead_capacity, D = (size_t)d;
    GL_REQUIRE(cap >= (size_t)n_deaths_total + N, "dead_capacity must be >= n_deaths_total + population_size");

    size_t off = 0;
    const size_t o_live = off;  off += a256(sizeof(double) * N * D);
    const size_t o_logl = off;  off += a256(sizeof(double) * N);
    const size_t o_dx = off;    off += a256(sizeof(double) * cap * D);
    const size_t o_dl = off;    off += a256(sizeof(double) * cap);
    const size_t o_dw = off;    off += a256(sizeof(double) * cap);
    const size_t o_di = off;    off += a256(sizeof(int64_t) * cap);
    const size_t o_ds = off;    off += a256(sizeof(int32_t) * cap);
    const size_t o_ws = off;    off += a256((size_t)ws_bytes);
    DeviceSlab slab;
    GL_CUDA(cudaMalloc(&slab.p, off));
    char *base = (char *)slab.p;
    gl_ns_buffers b;
    b.live = (double *)(base + o_live);
    b.logl = (double *)(base + o_logl);
    b.dead_x = (double *)(base + o_dx);
    b.dead_logl = (double *)(base + o_dl);
    b.dead_w = (double *)(base + o_dw);
    b.dead_idx = (int64_t *)(base + o_di);
    b.dead_step = (int32_t *)(base + o_ds);
    b.workspace = base + o_ws;

    cudaStream_t st;
    GL_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    gl_ns *ns = nullptr;
    int rc = gl_ns_create(cfg, priors_host, d, model, &b, &ns);
    if (rc == GL_OK) rc = gl_ns_init(ns, 0, st);
    if (rc == GL_OK) rc = gl_ns_advance(ns, n_deaths_total, st);
    if (rc == GL_OK) rc = gl_ns_finalize(ns, st);
    if (rc == GL_OK) rc = gl_ns_read_state(ns, state_out, st);
    if (rc == GL_OK) {
        const size_t rows = (size_t)state_out->n_rows;
        std::vector<double> hx(rows * D), hl(rows), hw(rows);
        std::vector<int64_t> hi(rows);
        std::vector<int32_t> hs(rows);
        cudaError_t e = cudaMemcpyAsync(hx.data(), b.dead_x, sizeof(double) * rows * D, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hl.data(), b.dead_logl, sizeof(double) * rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hw.data(), b.dead_w, sizeof(double) * rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hi.data(), b.dead_idx, sizeof(int64_t) * rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hs.data(), b.dead_step, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = gl::cuda_fail(e, "D2H dead table", __FILE__, __LINE__);
        if (rc == GL_OK) {
            // death order: batches in sequence, (logL key, slot) ascending inside a batch;
            // survivors (step == -1) keep their index order at the end.
            std::vector<size_t> order(rows);
            for (size_t i = 0; i < rows; ++i) order[i] = i;
            size_t lo = 0;
            while (lo < rows) {
                size_t hi_ = lo;
                while (hi_ < rows && hs[hi_] == hs[lo]) ++hi_;
                if (hs[lo] >= 0)
                    std::sort(order.begin() + lo, order.begin() + hi_, [&](size_t a, size_t c) {
                        const unsigned long long ka = gl::sort_key(hl[a]), kc = gl::sort_key(hl[c]);
                        return ka != kc ? ka < kc : hi[a] < hi[c];
                    });
                lo = hi_;
            }
            for (size_t r = 0; r < rows; ++r) {
                const size_t s = order[r];
                double *row = dead_host + r * (D + 2);
                row[0] = hl[s];
                row[1] = hw[s];
                for (size_t j = 0; j < D; ++j) row[2 + j] = hx[s * D + j];
            }
        }
    }
    if (ns) gl_ns_destroy(ns);
    cudaStreamDestroy(st);
    return rc;
}
