// Device-resident state of one Nested Sampling run (NestedSampling.run,
// nested_sampling.py:83-213) and the launch interfaces between the engine's parts.
#pragma once
#include "common.cuh"
#include "likelihoods.cuh"
#include "priors.cuh"
#include "select.cuh"

namespace gl {

struct EngineState {                 // one instance in device memory per engine
    long long n_dead;                // `_n_iterations`: deaths recorded
    long long n_rows;                // valid rows in the dead buffers
    int step;                        // batches whose refill has completed (Philox counter word)
    int k_cur;                       // size of the current batch (0 = nothing to do)
    int do_walk;                     // current batch is to be refilled
    int finalized;
    double logX;                     // ln remaining prior mass
    double lstar;                    // likelihood wall of the current batch
    double Z_lin, H_lin;             // reference's linear-space accumulators
    double M, Zs, Hs;                // scaled accumulators: Z = Zs*e^M, H = Hs*e^M
    double widths[GL_MAX_DIM];       // persisted proposal widths (`_widths`)
    unsigned int acc[GL_MAX_TUNING_CYCLES][GL_MAX_DIM];   // pooled acceptance counts of this step
    // finalize scratch
    double fin_m, fin_s, fin_lin;
    long long fin_n;                 // survivors over ALL ranks (weight denominator)
    long long fin_n_local;           // survivors of this engine (rows it appends)
    long long n_rows_dead;           // rows recorded by deaths (before the survivor rows)
    unsigned long long n_evals;      // log-likelihood evaluations performed by this engine
    // incrementally sorted live set (kill_small_kernel<true>)
    int sorted_ready;                // the sorted (key, slot) arrays are valid
    int sort_cur;                    // which of the two buffers is current
    int front_dead;                  // leading entries of the current buffer that were killed (now stale)
    int shard_overflow;              // sticky: a truncated candidate exchange could not prove its selection exact
    // bookkeeping of the current batch left to the first walker launch (kill_cluster_kernel, defer_book)
    double book_logw, book_w;        // ln weight / weight of every death of the batch
    long long book_base_row;         // first row of the batch in the dead log
    int book_step, book_room;        // batch id; the rows fit in the log
    double book_part[4][32];         // per-virtual-warp partial sums of the book blocks
    unsigned int book_count;         // book blocks finished (last one combines and resets)
};

// Peer-memory tables of the fused sharded step (comm.cu): base pointers of every rank's exchange
// buffer and arrival flags, mapped into this process through CUDA IPC over NVLink.
constexpr int GL_MAX_PEERS = 16;
struct PeerTables {
    double *gather[GL_MAX_PEERS];        // rank p's buffer: [2][world * cap] doubles (double-buffered by epoch parity)
    unsigned int *flags[GL_MAX_PEERS];   // rank p's arrival flags: [world] epochs
    long long cap;                       // candidates per rank the buffers were sized for
};

struct WalkParams {                  // by-value kernel parameter
    EngineState *S;
    const long long *dead_list;      // [k_cur] local slots of the current batch, ascending
    const unsigned char *is_dead;    // [N] 1 for slots in dead_list
    double *live;                    // [N, d]
    double *logl;                    // [N]
    const PriorTable *priors;        // device
    double *scratch_x;               // [k_max, d] walker positions between tuning cycles
    double *scratch_l;               // [k_max]
    long long N;                     // local population
    long long row_offset;
    uint64_t seed;
    int d, iterations, burn_in, tuning_cycles, proposal, k_max;
    int lanes;                       // lanes per walker (0 = choose from k_max)
    int force_smem_state;            // tests: use the generic shared-memory walker even for small d
    // deferred bookkeeping (walk_reg_kernel, phase 0): dead-row record by the walkers, evidence update by one extra block
    int book;                        // 0, or the number of values of the batch (global k)
    const double *book_vals;         // L(i) = book_vals[book_idx[i]], i < book: logl/dead_list, or gathered/positions
    const long long *book_idx;
    double n_global;
    double *dead_x, *dead_logl, *dead_w;
    long long *dead_idx;
    int *dead_step;
};

// walker.cu
int walk_launch(const WalkParams &P, const ModelDev &M, cudaStream_t st);

// loglike.cu / priors.cu
int loglike_eval_launch(const gl_model *m, const double *x_dev, long long n, double *logl_dev, cudaStream_t st);
int prior_draw_launch(const gl_prior *priors, int d, uint64_t seed, uint32_t tag, long long first_row,
                      long long n_rows, double *x_dev, cudaStream_t st);

// Proposal widths entering phase `phase` of the current step: the persisted widths adapted by
// the pooled acceptance of the tuning cycles already run (x0.66 below 0.2, x1.33 above 0.6;
// docstring semantics of samplers.py:26-30).
__device__ __forceinline__ double phase_width(const EngineState *S, int j, int phase, int n_walkers) {
    double w = S->widths[j];
    const double denom = 20.0 * (double)n_walkers;
    for (int c = 0; c < phase; ++c) {
        const double rate = (double)S->acc[c][j] / denom;
        if (rate < 0.2) w = w * 0.66;
        else if (rate > 0.6) w = w * 1.33;
    }
    return w;
}

}  // namespace gl
