// k-worst selection over the live log-likelihoods (np.argmin order generalised to k):
// exact MSB radix-select on the 64-bit ordering key, then an index-ordered compaction.
#pragma once
#include "common.cuh"

namespace gl {

constexpr int SEL_PASSES = 6;        // digits of 11,11,11,11,11,9 bits (MSB first)
constexpr int SEL_BINS = 2048;
constexpr int SEL_MAX_BLOCKS = 1024;
constexpr int SEL_THREADS = 256;

struct SelectState {                  // lives in device memory, zero-initialised once
    unsigned long long prefix;        // digits fixed so far
    unsigned long long T;             // threshold key: k-th smallest
    int k_rem;                        // still to take inside the current bucket
    int r_eq;                         // how many key == T elements are selected (lowest indices)
    int k_eff;                        // k actually selected (clamped to n)
    int pad;
    unsigned int ticket[SEL_PASSES];
    unsigned int hist[SEL_PASSES][SEL_BINS];
    int block_lt[SEL_MAX_BLOCKS];
    int block_eq[SEL_MAX_BLOCKS];
};

struct SelectPlan {
    int blocks;
    long long chunk;                  // elements per block (multiple of SEL_THREADS)
};

SelectPlan select_plan(long long n);

// Enqueue the 8 selection kernels.  k is read on the device from *k_ptr when k_ptr != NULL
// (engine: batch size lives in device state), else k_val.  Outputs: idx_out[0..k) ascending
// index (+ idx_base), *lstar_out, optional is_dead[i] = 1 flags.
int select_launch(const double *logl, long long n, const int *k_ptr, int k_val, long long idx_base,
                  long long *idx_out, double *lstar_out, unsigned char *is_dead, SelectState *S, cudaStream_t st);

}  // namespace gl
