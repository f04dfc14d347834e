/* gleipnir_b200 -- C ABI of the B200-native (sm_100a) Nested Sampling hot path.
 *
 * Drop-in boundary for the built-in classic Nested Sampling of LoLab-MSM/Gleipnir
 * (reference: pure Python, no FFI of its own -- SURVEY 8b).  Every entry point
 * names the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C, `extern "C"`, pointers + sizes only; no torch / C++ types;
 *   - `*_dev` pointers are CUDA device pointers owned by the caller (in the Python
 *     host they are torch tensors); `*_host` are ordinary host pointers;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *     all device work is enqueued on it and NOT synchronised unless stated;
 *   - every function returns GL_OK (0) or a negative GL_E* code; the message of the
 *     last failure on the calling thread is returned by gl_last_error();
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with GL_ECUDA.
 *   - all floating-point data is IEEE fp64; indices are int64 unless noted.
 */
#ifndef GLEIPNIR_B200_H
#define GLEIPNIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_ABI_VERSION 1

#if defined(__GNUC__)
#define GL_API __attribute__((visibility("default")))
#else
#define GL_API
#endif

#define GL_OK 0
#define GL_EINVAL (-1)   /* bad argument */
#define GL_ECUDA (-2)    /* CUDA runtime error / no device */
#define GL_ESTATE (-3)   /* call out of order */
#define GL_ENOMEM (-4)

#define GL_MAX_DIM 128   /* sampled parameters per model */
#define GL_MAX_TUNING_CYCLES 8

/* ---- library ------------------------------------------------------------------ */
GL_API int gl_abi_version(void);
GL_API const char *gl_last_error(void);
/* sm count / compute capability of the current device (GL_ECUDA when none). */
GL_API int gl_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- priors: SampledParameter (gleipnir/sampled_parameter.py:21-80) ------------ */
enum gl_prior_kind {
    GL_PRIOR_UNIFORM = 0,    /* scipy.stats.uniform(loc=p0, scale=p1), closed support */
    GL_PRIOR_NORM = 1,       /* scipy.stats.norm(loc=p0, scale=p1)                     */
    GL_PRIOR_LOGUNIFORM = 2  /* scipy.stats.loguniform(a=p0, b=p1) (= reciprocal)      */
};
typedef struct gl_prior {
    int32_t kind;
    int32_t reserved;
    double p0, p1;
} gl_prior;

/* SampledParameter.rvs (sampled_parameter.py:40-49) for a whole population:
 * x[r, j] = ppf_j(u(seed, first_row + r, j)), row-major [n_rows, d]; Philox4x32-10,
 * stream tag `tag` (GL_TAG_USER for user draws).  Replaces the N*d scalar rvs(1)
 * calls of nested_sampling.py:93-101. */
GL_API int gl_prior_draw(const gl_prior *priors_host, int32_t d, uint64_t seed, uint32_t tag,
                  int64_t first_row, int64_t n_rows, double *x_dev, void *stream);
/* SampledParameter.invcdf (sampled_parameter.py:72-80): out[i] = ppf(u[i]). */
GL_API int gl_prior_ppf(const gl_prior *prior_host, const double *u_dev, double *out_dev, int64_t n, void *stream);
/* SampledParameter.prior (sampled_parameter.py:62-70): out[i] = pdf(x[i]). */
GL_API int gl_prior_pdf(const gl_prior *prior_host, const double *x_dev, double *out_dev, int64_t n, void *stream);
/* SampledParameter.logprior (sampled_parameter.py:52-60): out[i] = log(pdf(x[i])). */
GL_API int gl_prior_logpdf(const gl_prior *prior_host, const double *x_dev, double *out_dev, int64_t n, void *stream);
/* Hypercube transform used by the third-party wrappers (multinest.py:131-132 etc.):
 * x[r, j] = ppf_j(u[r, j]) for a whole [n, d] block. */
GL_API int gl_prior_transform(const gl_prior *priors_host, int32_t d, const double *u_dev, double *x_dev, int64_t n_rows,
                       void *stream);
/* Host-buffer variants (copies inside; synchronous). */
GL_API int gl_prior_ppf_host(const gl_prior *prior_host, const double *u_host, double *out_host, int64_t n);
GL_API int gl_prior_pdf_host(const gl_prior *prior_host, const double *x_host, double *out_host, int64_t n);
GL_API int gl_prior_logpdf_host(const gl_prior *prior_host, const double *x_host, double *out_host, int64_t n);

/* ---- log-likelihood models: the batched device protocol ----------------------- */
/* The reference's protocol is `loglikelihood(ndarray[d]) -> float`
 * (nested_sampling.py:108, samplers.py:83,108,144).  A device model is an opaque
 * handle; gl_loglike_eval is the batched form of that call. */
typedef struct gl_model gl_model;

enum gl_model_kind {
    GL_MODEL_IID_GAUSSIAN = 0,  /* tests/test_nested_sampling.py:34-36                */
    GL_MODEL_EGGCARTON = 1,     /* examples/eggcarton/NS_eggcarton.py:25-27          */
    GL_MODEL_MIXTURE1D = 2,     /* examples/1d_multimodal/NS_1d_multimodal.py:20-32  */
    GL_MODEL_DENSE_GAUSSIAN = 3, /* synthetic C3 (not in reference) */
    GL_MODEL_EXTERNAL = 100      /* code-generated plugin (ODE networks)               */
};

/* -0.5*sum(((x-mu)/sigma)^2) - d*(0.5*log(2*pi) + log(sigma)) */
GL_API int gl_model_create_iid_gaussian(int32_t d, double mu, double sigma, gl_model **out);
/* (2 + prod_j cos(x_j))^5 */
GL_API int gl_model_create_eggcarton(int32_t d, gl_model **out);
/* log(mean_m N(x_0; positions[m], width)); NaN -> -inf.  d = 1. */
GL_API int gl_model_create_mixture1d(const double *positions_host, int32_t n_positions, double width, gl_model **out);
/* -0.5*(x-mu)^T P (x-mu) + log_norm; P row-major [d, d] symmetric. */
GL_API int gl_model_create_dense_gaussian(int32_t d, const double *mu_host, const double *P_host, double log_norm,
                                   gl_model **out);
/* A code-generated model (PySB-style mass-action ODE network, gleipnir_b200/odegen.py): the
 * plugin library compiled from the generated source exports gl_plugin_walk_launch /
 * gl_plugin_eval_launch, which instantiate this library's walker and batch-eval kernel
 * templates for the network's functor (stiff integrator + fused misfit).  `blob_host` is the
 * network's data block (nominal parameters, sampled-parameter map, time points, data, sd),
 * copied to the device.  Replaces the `loglikelihood` closures over ScipyOdeSimulator of
 * gleipnir/pysb_utilities/nestedsample_it.py:273-339. */
GL_API int gl_model_create_external(int32_t d, void *walk_fn, void *eval_fn, const void *blob_host,
                                    int64_t blob_bytes, gl_model **out);
/* Dense Gaussian: the batch evaluation uses the FP64 tensor cores (DMMA mma.sync.m8n8k4.f64)
 * by default; enable = 0 selects the scalar DFMA kernel (parity tests compare the two). */
GL_API int gl_model_set_tensor_path(gl_model *m, int32_t enable);
GL_API int gl_model_dim(const gl_model *m);
GL_API int gl_model_kind_of(const gl_model *m);
GL_API int gl_model_destroy(gl_model *m);

/* logl[i] = loglikelihood(x[i, :]), x row-major [n, d]. */
GL_API int gl_loglike_eval(const gl_model *m, const double *x_dev, int64_t n, double *logl_dev, void *stream);
GL_API int gl_loglike_eval_host(const gl_model *m, const double *x_host, int64_t n, double *logl_host);

/* ---- selection: np.argmin / k-worst (nested_sampling.py:115,156) -------------- */
/* The k smallest of logl[0..n) under the order (value, index): NaN first (np.argmin
 * semantics), -0.0 == +0.0, ties by lowest index.  Writes the k indices in ASCENDING
 * INDEX order to idx_out_dev[0..k) and the largest selected value to *lstar_dev.
 * `workspace_dev` must hold gl_select_workspace_bytes(n) bytes. */
GL_API int64_t gl_select_workspace_bytes(int64_t n);
GL_API int gl_select_kworst(const double *logl_dev, int64_t n, int32_t k, int64_t *idx_out_dev, double *lstar_dev,
                     void *workspace_dev, void *stream);

/* ---- the Nested Sampling engine: NestedSampling.run (nested_sampling.py:83-213) */
enum gl_proposal_kind { GL_PROPOSAL_UNIFORM = 0, GL_PROPOSAL_NORMAL = 1 };

typedef struct gl_ns_config {
    int64_t population_size; /* N (nested_sampling.py:60)                                   */
    int32_t batch_k;         /* deaths per step; 1 = the reference's sequencing              */
    int32_t iterations;      /* samplers.py:41  MetropolisComponentWiseHardNSRejection attrs */
    int32_t burn_in;         /* samplers.py:42                                               */
    int32_t tuning_cycles;   /* samplers.py:43 (each cycle = 20 sweeps)                      */
    int32_t proposal;        /* samplers.py:44  gl_proposal_kind                             */
    int32_t walker_lanes;    /* lanes cooperating on one walker (speculative prefetch): 0 = auto, else 1..32 */
    uint64_t seed;           /* Philox key                                                   */
    int64_t dead_capacity;   /* rows the dead buffers can hold (>= n_deaths + N)             */
    int64_t row_offset;      /* global index of local row 0 (multi-GPU shard), else 0        */
    int64_t global_population; /* N summed over ranks (== population_size on one GPU)       */
} gl_ns_config;

/* Caller-owned device buffers (torch tensors in the Python host). */
typedef struct gl_ns_buffers {
    double *live;      /* [N, d]   row-major live points  (`_live_points`, nested_sampling.py:103) */
    double *logl;      /* [N]      their log-likelihoods  (:108)                                   */
    double *dead_x;    /* [dead_capacity, d]  dead-point parameter rows (:130-133,170-173)         */
    double *dead_logl; /* [dead_capacity]     'log_l' column                                       */
    double *dead_w;    /* [dead_capacity]     'weight' column                                      */
    int64_t *dead_idx; /* [dead_capacity]     live slot each dead row came from                    */
    int32_t *dead_step;/* [dead_capacity]     batch index of each dead row                         */
    void *workspace;   /* gl_ns_workspace_bytes() bytes of scratch                                 */
} gl_ns_buffers;

/* Host-visible snapshot of the device-resident run state. */
typedef struct gl_ns_state {
    int64_t n_dead;       /* deaths recorded (`_n_iterations`)                          */
    int64_t n_rows;       /* rows valid in the dead buffers (n_dead + survivors after finalize) */
    int64_t n_evals;      /* log-likelihood evaluations performed on the device          */
    int64_t n_launches;   /* kernels of this library launched for the run so far        */
    int32_t step;         /* batches completed                                           */
    int32_t finalized;
    double log_x;         /* ln of remaining prior mass                                  */
    double lstar;         /* current likelihood wall                                     */
    double z_linear;      /* `_evidence` accumulated in linear space (nested_sampling.py:118-119) */
    double h_linear;      /* `_H` in linear space (:121-123)                             */
    double lse_m, lse_z, lse_h; /* scaled accumulators: Z = lse_z*exp(lse_m), H = lse_h*exp(lse_m) */
    double log_z;         /* ln Z from the scaled accumulators                           */
    double information;   /* -ln Z + H/Z (:125,167,194)                                  */
    double widths[GL_MAX_DIM]; /* current proposal widths (`_widths`, samplers.py:50,124) */
    int32_t shard_overflow; /* sharded run: a truncated candidate exchange could not prove a step exact
                             * (the results must be discarded and the run repeated with a full exchange) */
    int32_t reserved;
} gl_ns_state;

typedef struct gl_ns gl_ns;

GL_API int64_t gl_ns_workspace_bytes(const gl_ns_config *cfg, int32_t d);
GL_API int gl_ns_create(const gl_ns_config *cfg, const gl_prior *priors_host, int32_t d, const gl_model *model,
                 const gl_ns_buffers *buffers, gl_ns **out);
GL_API int gl_ns_destroy(gl_ns *ns);

/* nested_sampling.py:93-108 + samplers.py:70-81: draw the initial population with
 * Philox, evaluate its log-likelihoods, initialise the proposal widths.  With
 * `skip_population` != 0 the caller has already filled buffers.live / buffers.logl
 * (replay tests) and only the widths/state are initialised. */
GL_API int gl_ns_init(gl_ns *ns, int32_t skip_population, void *stream);

/* Enqueue batches (select k worst -> record -> accumulate -> refill by walkers) until
 * exactly `n_deaths_total` deaths have been recorded; the last batch is shortened as
 * needed and is not refilled.  nested_sampling.py:111-181. */
GL_API int gl_ns_advance(gl_ns *ns, int64_t n_deaths_total, void *stream);

/* Decomposed step, for replay tests and the multi-GPU host loop:
 *   gl_ns_kill   select + record + accumulate one batch of (at most) k deaths
 *   gl_ns_refill run the walkers for the last batch (nested_sampling.py:145-154)
 *   gl_ns_inject stand-in for the sampler: write caller-supplied replacement rows
 *                (death order = ascending (logL, index)) into the last batch's slots */
GL_API int gl_ns_kill(gl_ns *ns, int32_t k, void *stream);
GL_API int gl_ns_refill(gl_ns *ns, void *stream);
GL_API int gl_ns_inject(gl_ns *ns, const double *points_dev, const double *logl_dev, int32_t k, void *stream);

/* nested_sampling.py:183-203: final live-point term; appends the survivors to the
 * dead buffers in index order with weight X/(#survivors). */
/* `n` independent engines (seeds of one problem, models of a selection sweep -- HypSelector.run_nested_sampling,
 * pysb_utilities/hyp_selector.py:236-248, runs them one after another) advanced to n_deaths_total[i] deaths each,
 * concurrently: one host thread per engine enqueues on streams[i], so the runs' kernels overlap on the device.
 * Per-engine results are identical to n separate gl_ns_advance calls. */
GL_API int gl_ns_advance_many(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams);
/* The same with every run recorded into a CUDA graph first (gl_ns_advance_graph) and launched as one unit: the
 * device front end replays the launches, so the host's launch rate stops limiting concurrent runs.  Streams must
 * be non-default; engines must not have per-kernel profiling enabled. */
GL_API int gl_ns_advance_many_graphs(gl_ns **engines, int32_t n, const int64_t *n_deaths_total, void **streams);
GL_API int gl_ns_advance_graph(gl_ns *ns, int64_t n_deaths_total, void *stream);
GL_API int gl_ns_finalize(gl_ns *ns, void *stream);

/* `dead_points` (nested_sampling.py:211,338-344) assembled on the device: table_dev[n_rows, d+2]
 * (log_l, weight, x_0..x_{d-1}) in death order -- batches in sequence, ascending (logL, slot)
 * inside a batch (sorted per batch in shared memory; batches larger than 8192 keep slot order),
 * survivors last.  One kernel; the caller copies the table to the host. */
GL_API int gl_ns_export_dead(gl_ns *ns, double *table_dev, void *stream);

/* Copy the device state to the host (synchronises `stream`). */
GL_API int gl_ns_read_state(gl_ns *ns, gl_ns_state *out_host, void *stream);

/* ---- sharded run: one engine per GPU/rank ------------------------------
 * Rank r owns the contiguous global slots [row_offset, row_offset + population_size) of a
 * population of `global_population`.  Per step, with k = global batch size:
 *   1. gl_ns_shard_candidates   logL of this rank's k worst slots (ascending slot order);
 *   2. the caller all-gathers the k doubles of every rank, rank-major (NCCL / torch.distributed)
 *      -- the only data-path collective: the global likelihood threshold L*;
 *   3. gl_ns_shard_commit       every rank selects the global k worst from the gathered
 *      [world*k] array (position order == global-index order, so ties break as np.argmin
 *      would), records ITS dead rows, and updates evidence/information/prior mass identically
 *      and redundantly on every rank;
 *   4. gl_ns_refill             walkers refill this rank's dead slots, seeded from this rank's
 *      survivors.
 * At the end gl_ns_shard_final_partial / all-gather of 4 doubles per rank /
 * gl_ns_shard_finalize give the final live-point term (nested_sampling.py:183-203). */
GL_API int gl_ns_shard_candidates(gl_ns *ns, int32_t k, double *cand_logl_dev, void *stream);
GL_API int gl_ns_shard_commit(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank, int32_t k,
                              void *stream);
/* same, with a caller-provided scratch of world*k 8-byte words for gathered sets too large for
 * shared memory (keeps the commit a single fused launch up to world*k = 261 120) */
GL_API int gl_ns_shard_commit_scratch(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank,
                                      int32_t k, void *keys_scratch, void *stream);
/* Truncated exchange.  With the live points spread at random over the ranks, a rank holds
 * Binomial(k, 1/world) of a step's k deaths, so exchanging all k candidates per rank is mostly
 * wasted work (and the redundant selection over world*k keys is the scaling bottleneck).
 * gl_ns_shard_candidate_count returns m = mean + `sigmas` standard deviations (clamped to
 * [ceil(k/world), min(k, n_local-1)]; sigmas <= 0 or world == 1: m = k).  Steps 1-3 then run with
 * m candidates per rank: gl_ns_shard_candidates(ns, m, ...), all-gather of m doubles,
 * gl_ns_shard_commit_truncated(..., k, m, ...).  The commit PROVES the selection exact (no rank had
 * all m of its candidates selected); when it cannot, it raises the sticky gl_ns_state.shard_overflow
 * flag and the caller repeats the run with m = k -- results are never silently approximate. */
GL_API int64_t gl_ns_local_population(const gl_ns *ns);
GL_API int32_t gl_ns_shard_candidate_count(int32_t k, int32_t world, int64_t n_local, double sigmas);
GL_API int gl_ns_shard_commit_truncated(gl_ns *ns, const double *gathered_logl_dev, int32_t world, int32_t rank,
                                        int32_t k, int32_t m, void *keys_scratch, void *stream);
GL_API int gl_ns_shard_final_partial(gl_ns *ns, double *partial_dev, void *stream);
GL_API int gl_ns_shard_finalize(gl_ns *ns, const double *gathered_partials_dev, int32_t world, int32_t rank,
                                void *stream);

/* Sharded run, table assembly on the device: `packed_dev` = the ranks' logs all-gathered as rows
 * [world][max_rows][d+4] = (log_l, weight, x[d], global slot, step; survivors carry step -1 and
 * come last in a rank's block), `counts_dev[world]` valid rows per rank.  Writes the
 * [sum(counts), d+2] table in death order (nested_sampling.py:130-133,195-203,211): batches in
 * sequence, ascending (log_l, slot) inside a batch, survivors last in global slot order. */
GL_API int gl_export_gathered_table(const double *packed_dev, const int64_t *counts_dev, int32_t world,
                                    int64_t max_rows, int32_t d, int32_t batch_k, int32_t n_steps,
                                    int64_t n_dead_total, double *table_dev, void *stream);

/* Native sharded loop: NCCL all-gathers enqueued on the engine's stream from C++ (libnccl is
 * dlopen'ed on first use).  gl_comm_unique_id on rank 0 -> broadcast the bytes by any means
 * (torch.distributed in the Python host) -> gl_comm_create on every rank (collective). */
typedef struct gl_comm gl_comm;
GL_API int gl_comm_unique_id_bytes(void);
GL_API int gl_comm_unique_id(uint8_t *out_bytes);
GL_API int gl_comm_create(int32_t world, int32_t rank, const uint8_t *id_bytes, gl_comm **out);
GL_API int gl_comm_destroy(gl_comm *c);
GL_API int gl_comm_all_gather_f64(gl_comm *c, const double *send_dev, double *recv_dev, int64_t count, void *stream);
/* Peer-memory exchange: with it gl_ns_advance_sharded runs candidates + exchange + commit as ONE
 * kernel per step that stores the candidates straight into every peer's buffer over NVLink and
 * synchronises with system-scope flags -- no NCCL call on the step path.  Setup: every rank calls
 * gl_comm_p2p_alloc (cap >= batch_k candidates per rank) and gets its CUDA IPC handle
 * (gl_comm_p2p_handle_bytes bytes); the host all-gathers the handles rank-major by any means;
 * every rank calls gl_comm_p2p_open; one barrier.  Without it the NCCL all-gather path is used. */
GL_API int gl_comm_p2p_handle_bytes(void);
GL_API int gl_comm_p2p_alloc(gl_comm *c, int64_t cap, uint8_t *handle_out);
GL_API int gl_comm_p2p_open(gl_comm *c, const uint8_t *handles_all);
GL_API int gl_comm_p2p_ready(const gl_comm *c);
GL_API int gl_comm_p2p_disable(gl_comm *c);   /* collective decision of the host: all ranks use the NCCL path */
/* Enqueue steps (refill of the waiting batch, candidates, all-gather, commit) until
 * `n_deaths_total` global deaths; `n_dead_so_far` = global deaths already recorded.
 * `candidate_sigmas` > 0 selects the truncated exchange (check gl_ns_state.shard_overflow after
 * the run); 0 exchanges all k candidates per rank. */
GL_API int gl_ns_advance_sharded(gl_ns *ns, gl_comm *c, int64_t n_deaths_total, int32_t batch_k,
                                 int64_t n_dead_so_far, double candidate_sigmas, void *stream);
/* The same loop for a run that is driven batch by batch (monitoring, user-defined stopping criteria evaluated
 * between batches -- `criterion(nested_sampler) -> bool`, nested_sampling.py:141,215-217): `n_batches` steps, each
 * one refilling the waiting batch and killing the next; the last batch of the call stays waiting and IS refilled
 * by the next call (gl_ns_advance_sharded treats the batch that reaches n_deaths_total as the run's last). */
GL_API int gl_ns_step_sharded(gl_ns *ns, gl_comm *c, int32_t n_batches, int32_t batch_k, int64_t n_dead_so_far,
                              double candidate_sigmas, void *stream);
GL_API int gl_ns_finalize_sharded(gl_ns *ns, gl_comm *c, void *stream);

/* Populations whose ordering keys fit in shared memory (N <= 24576) use a single-CTA fused
 * kill kernel; this switch forces the general multi-launch selection path (used by tests so
 * that both paths are checked against each other and the oracle). */
GL_API int gl_ns_force_multipass_select(gl_ns *ns, int32_t enable);
/* Selection path: 0 auto (N <= 196608: fused kill on an 8-CTA thread-block cluster; else
 * multi-launch), 1 multi-launch, 2 fused single-CTA radix (N <= 24576), 3 fused kill over an
 * incrementally sorted live set (N <= 16384; measured slower than the radix select on one SM, kept
 * as a tested alternative), 4 cluster radix.
 * All paths produce identical results (tests compare them). */
/* Walkers keep their state in registers when the model offers that form and d <= 8; this switch
 * forces the generic shared-memory walker (tests compare the two bit for bit). */
GL_API int gl_ns_force_smem_walker(gl_ns *ns, int32_t enable);
GL_API int gl_ns_set_select_path(gl_ns *ns, int32_t path);

/* Optional CUDA-event timing of the two kernel groups that dominate a step, recorded on the
 * launching stream: walker launches (tuning_cycles+1 per step) and selection launches (8 per
 * step).  gl_ns_profile_read synchronises on the recorded events and returns the summed
 * device time (ms) and the number of launches covered. */
GL_API int gl_ns_profile_enable(gl_ns *ns, int32_t enable);
GL_API int gl_ns_profile_read(gl_ns *ns, double *walk_ms, int64_t *walk_launches, double *select_ms,
                              int64_t *select_launches);

/* ---- whole run with HOST buffers (the end-to-end entry point) ------------------ */
/* NestedSampling(...).run() through host memory: allocates device buffers, uploads
 * the prior table, runs init/advance/finalize, downloads the dead table
 * [n_rows, d+2] (columns log_l, weight, x_0..x_{d-1}, rows in death order) and the
 * state.  `dead_host` must hold (n_deaths_total + N) * (d + 2) doubles. */
GL_API int gl_ns_run_host(const gl_ns_config *cfg, const gl_prior *priors_host, int32_t d, const gl_model *model,
                   int64_t n_deaths_total, double *dead_host, gl_ns_state *state_out);

/* ---- posteriors(): weighted marginal histograms (nested_sampling.py:266-293) ---------------------
 * Stage 1 writes the normalised weights w_i*exp(L_i)/Z (as exp(log w + L - ln Z)) and returns the
 * number of rows with positive weight, from which the caller takes the Rice bin count
 * nbins = 2*int(cbrt(count)) (:286).  Stage 2 builds, per parameter, np.histogram(x[mask],
 * weights=.., density=True, bins=nbins) with numpy's own bin-assignment rule, deterministically.
 * The dead_* pointers are the engine's caller-owned buffers; rows = gl_ns_state.n_rows. */
GL_API int64_t gl_posterior_stats_bytes(void);
GL_API int gl_posterior_weights(const double *dead_logl_dev, const double *dead_w_dev, const double *dead_x_dev,
                                int64_t rows, int32_t d, double log_z, double *nw_dev, void *stats_dev,
                                int64_t *n_masked_host, void *stream);
GL_API int64_t gl_posterior_partial_doubles(int64_t rows, int32_t d, int32_t nbins);
GL_API int gl_posterior_histograms(const double *dead_x_dev, const double *nw_dev, int64_t rows, int32_t d,
                                   int32_t nbins, const void *stats_dev, double *partial_dev, double *marginal_dev,
                                   double *edges_dev, double *centers_dev, void *stream);

/* ---- the sampler's per-call protocol -------------------------------------------------------------
 * MetropolisComponentWiseHardNSRejection.__call__(sampled_parameters, loglikelihood, start_param_vec,
 * ns_boundary) -> (point, logL)  (samplers.py:56-152): ONE walker from `start_host` under the wall
 * `ns_boundary`, run by the same walker kernel as the batched engine.  `widths_inout_host[d]` carries
 * the sampler's persistent `_widths` (have_widths = 0 on the first call: initialised from 100 prior
 * draws as samplers.py:70-81); `call_index` separates the Philox streams of successive calls. */
GL_API int gl_sampler_call_host(const gl_prior *priors_host, int32_t d, const gl_model *model, int32_t iterations,
                                int32_t burn_in, int32_t tuning_cycles, int32_t proposal, uint64_t seed,
                                int32_t call_index, const double *start_host, double ns_boundary,
                                double *widths_inout_host, int32_t have_widths, double *point_out_host,
                                double *logl_out_host);

/* ---- reference-exact validation mode ----------------------------------------------------------
 * The reference AS SHIPPED (nested_sampling.py:83-213 + samplers.py:56-152 with its quirks,
 * numpy's legacy MT19937 stream seeded by `seed`, the reference's draw order), replayed on the
 * device for scipy `uniform` priors, the 'uniform' proposal and the iid-Gaussian likelihood
 * -0.5*sum(x*x) + cst -- the configurations whose arithmetic is IEEE-exact on both sides.
 * weights_host = [w_1..w_n, alpha**n] computed by the caller with the reference's expressions.
 * dead_host[(n_iterations + N - 1), d + 2] is bit-identical to the reference's `dead_points`;
 * scalars_host = {Z, H, n_evals, information} (device exp/log: <= 1e-15 relative). */
GL_API int gl_reference_exact_run_host(uint32_t seed, int32_t N, int32_t d, const double *loc_host,
                                       const double *scale_host, double cst, int32_t iterations,
                                       int32_t burn_in, int32_t tuning_cycles, int32_t n_iterations,
                                       const double *weights_host, double *dead_host, double *widths_host,
                                       double *scalars_host);

/* ---- measurement helpers ------------------------------------------------------ */
/* FP64 DFMA-pipe peak microbenchmark (roofline denominator for the walker / ODE kernels;
 * MEASURED_PEAKS.json has no FP64 entry).  Enqueues `blocks` CTAs of 256 threads running
 * `iters` x 8 independent DFMAs each; *flops_out = floating-point operations enqueued.
 * `scratch_dev` must hold blocks*256 doubles.  Time it with CUDA events on `stream`. */
/* cycles per DEPENDENT DFMA for one warp (the walker's chain is bound by this, not by the pipe's throughput) */
GL_API int gl_microbench_dfma_latency(int32_t iters, double *scratch_dev, double *cycles_per_dfma_out);
GL_API int gl_microbench_dfma(int32_t blocks, int32_t iters, double *scratch_dev, int64_t *flops_out, void *stream);
/* Same for the FP64 tensor cores: `blocks` CTAs of 8 warps, `iters` x 4 DMMA m8n8k4 per warp. */
GL_API int gl_microbench_dmma(int32_t blocks, int32_t iters, double *scratch_dev, int64_t *flops_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GLEIPNIR_B200_H */
