// Constrained component-wise Metropolis walkers: the B200 form of
// MetropolisComponentWiseHardNSRejection.__call__ (samplers.py:56-152) applied to a whole
// batch of dead slots at once.  One walker per thread; each replaces one dead slot:
//   * seed  = uniformly chosen survivor, int(U*N) redrawn while it hits a dead slot
//             (nested_sampling.py:147-149);
//   * move  = the reference's *tuning-loop* move (samplers.py:95-114): propose component j
//             from a symmetric uniform/normal window of width w_j, accept iff
//             u < prior_j(new)/prior_j(cur) AND logL(new) > L*  (hard NS wall);
//   * phases: `tuning_cycles` launches of 20 sweeps whose positions are discarded and whose
//             pooled acceptance adapts w (see phase_width), then one launch of
//             iterations+burn_in sweeps from the seed again whose end point is written
//             into the dead slot (samplers.py:124-152 without quirk Q1).
// Randomness: Philox4x32-10, counter (global slot, step, sweep*d + component, TAG_MOVE): one
// block yields the proposal uniform r and the acceptance uniform u of a trial move, so
// results do not depend on grid shape or sharding.
// State: the walker's point lives in shared memory as a per-thread column x[j*T + tid]
// (conflict-free for any d, dynamic component index without local memory).
// Bound: FP64 pipe / latency (no HBM traffic beyond 8(2d+1) B per replacement).
#include "engine.cuh"
#include "philox.cuh"

namespace gl {

constexpr int WALK_THREADS = 64;

template <class Like>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(WalkParams P, ModelDev M, int phase) {
    extern __shared__ double sm[];
    EngineState *S = P.S;
    if (!S->do_walk) return;                       // grid-uniform
    const int k = S->k_cur;
    const int T = WALK_THREADS;
    const int d = P.d;
    double *x = sm;                                // [d][T]
    double *wj = sm + (size_t)d * T;               // [d] widths of this phase
    unsigned int *acc = reinterpret_cast<unsigned int *>(wj + d);   // [d]
    const int tid = threadIdx.x;
    const int w = blockIdx.x * T + tid;
    if (blockIdx.x * T >= k) return;               // block-uniform
    const bool active = w < k;
    const bool tuning = phase < P.tuning_cycles;

    for (int j = tid; j < d; j += T) {
        wj[j] = phase_width(S, j, phase, k);
        acc[j] = 0u;
    }
    __syncthreads();

    const Like like(M);
    const double lstar = S->lstar;
    const uint32_t step = (uint32_t)S->step;
    const PriorTable *pt = P.priors;
    double L = 0.0;
    uint32_t g = 0;
    long long slot = 0;
    if (active) {
        slot = P.dead_list[w];
        g = (uint32_t)(slot + P.row_offset);
        if (phase == 0 || !tuning) {
            // survivor pick
            long long pick;
            for (uint32_t a = 0;; ++a) {
                const U2 u = philox_uniforms(P.seed, g, step, a, GL_TAG_PICK);
                pick = (long long)(u.a * (double)P.N);
                if (pick >= P.N) pick = P.N - 1;
                if (!P.is_dead[pick]) break;
            }
            const double *src = P.live + pick * d;
            for (int j = 0; j < d; ++j) x[j * T + tid] = src[j];
            L = P.logl[pick];
        } else {
            const double *src = P.scratch_x + (size_t)w * d;
            for (int j = 0; j < d; ++j) x[j * T + tid] = src[j];
            L = P.scratch_l[w];
        }
    }

    const int n_sweeps = tuning ? 20 : P.iterations + P.burn_in;
    const int t0 = tuning ? phase * 20 : P.tuning_cycles * 20;
    const bool normal = P.proposal == GL_PROPOSAL_NORMAL;
    for (int t = t0; t < t0 + n_sweeps; ++t) {
        for (int j = 0; j < d; ++j) {
            bool ok = false;
            if (active) {
                const U2 ru = philox_uniforms(P.seed, g, step, (uint32_t)(t * d + j), GL_TAG_MOVE);
                const double width = wj[j];
                const double xj = x[j * T + tid];
                const double nj = normal ? normcdfinv(ru.a) * width + xj : ru.a * width + (xj - width / 2.0);
                const int kind = pt->kind[j];
                const double p0 = pt->p0[j], p1 = pt->p1[j];
                const double ratio = prior_pdf(kind, p0, p1, nj) / prior_pdf(kind, p0, p1, xj);
                x[j * T + tid] = nj;
                const double Lt = like.eval(x + tid, T);
                ok = (ru.b < ratio) && (Lt > lstar);
                if (ok) L = Lt;
                else x[j * T + tid] = xj;
            }
            if (tuning) {
                const unsigned int m = __ballot_sync(0xffffffffu, ok);
                if ((tid & 31) == 0 && m) atomicAdd(&acc[j], (unsigned int)__popc(m));
            }
        }
    }

    if (tuning) {
        if (active && phase + 1 < P.tuning_cycles) {
            double *dst = P.scratch_x + (size_t)w * d;
            for (int j = 0; j < d; ++j) dst[j] = x[j * T + tid];
            P.scratch_l[w] = L;
        }
        __syncthreads();
        for (int j = tid; j < d; j += T)
            if (acc[j]) atomicAdd(&S->acc[phase][j], acc[j]);
    } else if (active) {
        double *dst = P.live + slot * d;
        for (int j = 0; j < d; ++j) dst[j] = x[j * T + tid];
        P.logl[slot] = L;
    }
}

int walker_smem_bytes(int d, int threads) {
    return (int)(sizeof(double) * ((size_t)d * threads + d) + sizeof(unsigned int) * d + 16);
}

struct WalkLauncher {
    const WalkParams &P;
    const ModelDev &M;
    cudaStream_t st;
    template <class Like>
    int operator()() {
        const int smem = walker_smem_bytes(P.d, WALK_THREADS);
        static thread_local int configured_smem = 0;
        if (smem > 48 * 1024 && smem > configured_smem) {
            GL_CUDA(cudaFuncSetAttribute(walk_kernel<Like>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured_smem = smem;
        }
        const int blocks = (P.k_max + WALK_THREADS - 1) / WALK_THREADS;
        for (int phase = 0; phase <= P.tuning_cycles; ++phase)
            walk_kernel<Like><<<blocks, WALK_THREADS, smem, st>>>(P, M, phase);
        GL_CHECK_LAUNCH();
        return GL_OK;
    }
};

int walk_launch(const WalkParams &P, const ModelDev &M, cudaStream_t st) {
    WalkLauncher L{P, M, st};
    return dispatch_model(M.kind, L);
}

}  // namespace gl
