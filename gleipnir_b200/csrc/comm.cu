// NCCL plumbing for the sharded run: one rank per process/GPU, the only data-path
// collective is an all-gather of k doubles per rank per step (global likelihood threshold L*)
// plus one of 4 doubles per rank for the final live-point term.  The collective is enqueued on
// the SAME stream as the engine's kernels, from the same C++ loop, so a whole sharded run is
// enqueued without host round trips.
//
// libnccl is resolved lazily with dlopen/dlsym: the library carries no link-time NCCL dependency
// and single-GPU use never touches it.  Inside a PyTorch process "libnccl.so.2" resolves to the
// copy torch already loaded.
#include <dlfcn.h>
#include <nccl.h>
#include "engine.cuh"

struct gl_comm {
    ncclComm_t comm;
    int world, rank;
    double *send;        // [cap]        this rank's candidates / partials
    double *recv;        // [world*cap]  gathered
    double *keys;        // [world*cap]  ordering-key scratch of the fused commit kernel
    size_t cap;
    // peer-memory exchange of the fused step (gl_comm_p2p_*): one allocation per rank,
    // [64 arrival flags (u32) | 2 x world*p2p_cap doubles], opened on every peer through CUDA IPC
    void *p2p_local;
    void *p2p_peer[gl::GL_MAX_PEERS];
    gl::PeerTables tables;
    bool p2p_ready;
    unsigned int epoch;  // steps exchanged so far (identical on every rank)
};

namespace gl {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl() {
    static NcclApi api;
    return api;
}

static int nccl_load() {
    NcclApi &a = nccl();
    if (a.handle) return GL_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) {
        set_error("cannot load libnccl (%s)", dlerror());
        return GL_ECUDA;
    }
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        set_error("libnccl lacks a required symbol");
        a.handle = nullptr;
        return GL_ECUDA;
    }
    return GL_OK;
}

static int nccl_fail(ncclResult_t r, const char *what) {
    set_error("NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    return GL_ECUDA;
}

#define GL_NCCL(call)                                              \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != ncclSuccess) return gl::nccl_fail(r__, #call);  \
    } while (0)

int comm_reserve(gl_comm *c, size_t count) {
    if (count <= c->cap) return GL_OK;
    if (c->send) cudaFree(c->send);
    if (c->recv) cudaFree(c->recv);
    if (c->keys) cudaFree(c->keys);
    c->send = c->recv = c->keys = nullptr;
    GL_CUDA(cudaMalloc(&c->send, sizeof(double) * count));
    GL_CUDA(cudaMalloc(&c->recv, sizeof(double) * count * (size_t)c->world));
    GL_CUDA(cudaMalloc(&c->keys, sizeof(double) * count * (size_t)c->world));
    c->cap = count;
    return GL_OK;
}

int comm_all_gather(gl_comm *c, const double *send, double *recv, size_t count, cudaStream_t st) {
    GL_NCCL(nccl().AllGather(send, recv, count, ncclDouble, c->comm, st));
    return GL_OK;
}

}  // namespace gl

extern "C" int gl_ns_shard_commit_truncated(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank,
                                            int32_t k, int32_t m, void *keys_scratch, void *stream);
extern "C" int32_t gl_ns_shard_candidate_count(int32_t k, int32_t world, int64_t n_local, double sigmas);
extern "C" int64_t gl_ns_local_population(const gl_ns *ns);
extern "C" void gl_ns_internal_sharded_loop(gl_ns *ns, int on);
extern "C" void gl_ns_internal_commit_refills(gl_ns *ns, int yes);
extern "C" int gl_ns_internal_shard_step_p2p(gl_ns *ns, const gl::PeerTables *tables, int32_t world, int32_t rank,
                                             int32_t k, int32_t m, uint32_t epoch, void *stream);

extern "C" {

int gl_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

int gl_comm_unique_id(uint8_t *out_bytes) {
    GL_REQUIRE(out_bytes != nullptr, "out is NULL");
    int rc = gl::nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    GL_NCCL(gl::nccl().GetUniqueId(&id));
    memcpy(out_bytes, &id, sizeof(id));
    return GL_OK;
}

int gl_comm_create(int32_t world, int32_t rank, const uint8_t *id_bytes, gl_comm **out) {
    GL_REQUIRE(out != nullptr && id_bytes != nullptr, "NULL argument");
    GL_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad world/rank");
    int rc = gl::nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    gl_comm *c = new gl_comm();
    c->world = world;
    c->rank = rank;
    c->send = c->recv = c->keys = nullptr;
    c->cap = 0;
    c->p2p_local = nullptr;
    for (int p = 0; p < gl::GL_MAX_PEERS; ++p) c->p2p_peer[p] = nullptr;
    c->p2p_ready = false;
    c->epoch = 0;
    ncclResult_t r = gl::nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return gl::nccl_fail(r, "ncclCommInitRank");
    }
    *out = c;
    return GL_OK;
}

int gl_comm_destroy(gl_comm *c) {
    if (!c) return GL_OK;
    if (c->send) cudaFree(c->send);
    if (c->recv) cudaFree(c->recv);
    if (c->keys) cudaFree(c->keys);
    for (int p = 0; p < c->world && p < gl::GL_MAX_PEERS; ++p)
        if (p != c->rank && c->p2p_peer[p]) cudaIpcCloseMemHandle(c->p2p_peer[p]);
    if (c->p2p_local) cudaFree(c->p2p_local);
    if (c->comm) gl::nccl().CommDestroy(c->comm);
    delete c;
    return GL_OK;
}

// ---- peer-memory exchange (fused sharded step) ------------------------------------------------------
// 1. every rank: gl_comm_p2p_alloc -> its CUDA IPC handle;  2. the host all-gathers the handles (any
// transport);  3. every rank: gl_comm_p2p_open maps the peers' buffers (NVLink peer access).  A barrier
// between 3 and the first step keeps a fast rank from writing into a buffer that is still being zeroed.
int gl_comm_p2p_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

static size_t p2p_bytes(int world, size_t cap) { return 256 + 2 * sizeof(double) * (size_t)world * cap; }

int gl_comm_p2p_alloc(gl_comm *c, int64_t cap, uint8_t *handle_out) {
    GL_REQUIRE(c != nullptr && handle_out != nullptr && cap >= 1, "bad argument");
    GL_REQUIRE(c->world <= gl::GL_MAX_PEERS, "world=%d > %d peers", c->world, gl::GL_MAX_PEERS);
    GL_REQUIRE(c->p2p_local == nullptr, "peer buffers already allocated");
    const size_t bytes = p2p_bytes(c->world, (size_t)cap);
    GL_CUDA(cudaMalloc(&c->p2p_local, bytes));
    GL_CUDA(cudaMemset(c->p2p_local, 0, bytes));
    GL_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    GL_CUDA(cudaIpcGetMemHandle(&h, c->p2p_local));
    memcpy(handle_out, &h, sizeof(h));
    c->tables.cap = cap;
    return GL_OK;
}

int gl_comm_p2p_open(gl_comm *c, const uint8_t *handles_all) {
    GL_REQUIRE(c != nullptr && handles_all != nullptr && c->p2p_local != nullptr, "bad argument / alloc first");
    for (int p = 0; p < c->world; ++p) {
        void *base = c->p2p_local;
        if (p != c->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles_all + (size_t)p * sizeof(h), sizeof(h));
            GL_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            c->p2p_peer[p] = base;
        }
        c->tables.flags[p] = reinterpret_cast<unsigned int *>(base);
        c->tables.gather[p] = reinterpret_cast<double *>(reinterpret_cast<char *>(base) + 256);
    }
    c->p2p_ready = true;
    return GL_OK;
}

int gl_comm_p2p_ready(const gl_comm *c) { return c != nullptr && c->p2p_ready ? 1 : 0; }

// every rank must take the same path: when mapping failed anywhere, everybody falls back to NCCL
int gl_comm_p2p_disable(gl_comm *c) {
    GL_REQUIRE(c != nullptr, "comm is NULL");
    c->p2p_ready = false;
    return GL_OK;
}

int gl_comm_all_gather_f64(gl_comm *c, const double *send_dev, double *recv_dev, int64_t count, void *stream) {
    GL_REQUIRE(c != nullptr && send_dev != nullptr && recv_dev != nullptr && count >= 1, "bad argument");
    return gl::comm_all_gather(c, send_dev, recv_dev, (size_t)count, (cudaStream_t)stream);
}

// Enqueue the sharded run until `n_deaths_total` GLOBAL deaths: per step candidates -> NCCL
// all-gather -> commit -> refill, all on `stream`.
}  // extern "C"

// `open_ended`: the call's last batch may be refilled by a later call (stepping a long run batch by batch);
// the device-side protocol is the same, the flag only documents the caller's intent.
static int advance_sharded_impl(gl_ns *ns, gl_comm *c, int64_t n_deaths_total, int32_t batch_k, int64_t n_dead_so_far,
                                double candidate_sigmas, bool open_ended, void *stream);

extern "C" {

int gl_ns_advance_sharded(gl_ns *ns, gl_comm *c, int64_t n_deaths_total, int32_t batch_k, int64_t n_dead_so_far,
                          double candidate_sigmas, void *stream) {
    return advance_sharded_impl(ns, c, n_deaths_total, batch_k, n_dead_so_far, candidate_sigmas, false, stream);
}

int gl_ns_step_sharded(gl_ns *ns, gl_comm *c, int32_t n_batches, int32_t batch_k, int64_t n_dead_so_far,
                       double candidate_sigmas, void *stream) {
    GL_REQUIRE(n_batches >= 1, "n_batches must be >= 1");
    return advance_sharded_impl(ns, c, n_dead_so_far + (int64_t)n_batches * batch_k, batch_k, n_dead_so_far,
                                candidate_sigmas, true, stream);
}

}  // extern "C"

static int advance_sharded_impl(gl_ns *ns, gl_comm *c, int64_t n_deaths_total, int32_t batch_k, int64_t n_dead_so_far,
                                double candidate_sigmas, bool open_ended, void *stream) {
    GL_REQUIRE(ns != nullptr && c != nullptr, "NULL argument");
    GL_REQUIRE(batch_k >= 1, "batch_k must be >= 1");
    int rc = gl::comm_reserve(c, (size_t)(batch_k > 4 ? batch_k : 4));
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_local = gl_ns_local_population(ns);
    int64_t done = n_dead_so_far;
    bool first = true;
    struct LoopGuard {                       // refills inside this loop leave their epilogue to the candidates kernel
        gl_ns *ns;
        explicit LoopGuard(gl_ns *p) : ns(p) { gl_ns_internal_sharded_loop(ns, 1); }
        ~LoopGuard() { gl_ns_internal_sharded_loop(ns, 0); }
    } guard(ns);
    while (done < n_deaths_total) {
        const int32_t k = (int32_t)((n_deaths_total - done) < batch_k ? (n_deaths_total - done) : batch_k);
        if (!first || n_dead_so_far > 0) {
            rc = gl_ns_refill(ns, stream);
            if (rc) return rc;
        }
        first = false;
        const int32_t m = gl_ns_shard_candidate_count(k, c->world, n_local, candidate_sigmas);
        // "a refill follows inside this call": only then may the commit leave its bookkeeping to the walkers.  The
        // last batch of an open-ended call is refilled by a LATER call, or finalised instead: it records its own rows.
        gl_ns_internal_commit_refills(ns, done + k < n_deaths_total ? 1 : 0);
        if (c->p2p_ready && m <= c->tables.cap && n_local <= 196608 && (int64_t)c->world * m <= 196608) {
            // fused step: candidates, exchange over NVLink peer memory and commit in ONE launch
            c->epoch += 1;
            rc = gl_ns_internal_shard_step_p2p(ns, &c->tables, c->world, c->rank, k, m, c->epoch, stream);
            if (rc) return rc;
            done += k;
            continue;
        }
        rc = gl_ns_shard_candidates(ns, m, c->send, stream);
        if (rc) return rc;
        rc = gl::comm_all_gather(c, c->send, c->recv, (size_t)m, st);
        if (rc) return rc;
        rc = gl_ns_shard_commit_truncated(ns, c->recv, c->world, c->rank, k, m, c->keys, stream);
        if (rc) return rc;
        done += k;
    }
    return GL_OK;
}

extern "C" {

int gl_ns_finalize_sharded(gl_ns *ns, gl_comm *c, void *stream) {
    GL_REQUIRE(ns != nullptr && c != nullptr, "NULL argument");
    int rc = gl::comm_reserve(c, 4);
    if (rc) return rc;
    rc = gl_ns_shard_final_partial(ns, c->send, stream);
    if (rc) return rc;
    rc = gl::comm_all_gather(c, c->send, c->recv, 4, (cudaStream_t)stream);
    if (rc) return rc;
    return gl_ns_shard_finalize(ns, c->recv, c->world, c->rank, stream);
}

}  // extern "C"
