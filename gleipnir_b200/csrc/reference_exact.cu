// Reference-exact mode: the reference AS SHIPPED, bit for bit, on the device.
//
// Replays NestedSampling.run (nested_sampling.py:83-213) with the shipped sampler
// (samplers.py:56-152, quirks Q1/Q2 of SURVEY 3.3 INCLUDED) drawing from numpy's legacy global
// MT19937 in exactly the reference's order, for the configurations in which every arithmetic
// step is IEEE-754 exact on both sides:
//   priors     scipy.stats.uniform(loc, scale)      rvs = u*scale + loc; pdf = 1/scale on the closed support
//   proposal   'uniform'                            ppf(r, loc=x-w/2, scale=w) = r*w + (x - w/2)
//   likelihood iid Gaussian of tests/test_nested_sampling.py:34-36   -0.5*sum(x*x) + const
// (norm priors / normal proposals / cos() involve libm transcendentals whose last bits differ
// between numpy and CUDA, so they cannot be reproduced bit-exactly and are rejected.)
// Weights alpha**(n-1) - alpha**n come from the caller (computed with the reference's own Python
// expression, nested_sampling.py:112,143,184), so dead rows -- log_l, weight, parameters, order --
// are BIT-IDENTICAL to the reference's `dead_points`; Z and H use the device exp (<= 1e-15 rel).
// The algorithm is strictly sequential (one death per iteration, one global RNG stream): one
// thread runs it; this is a validation mode, not a fast path.
#include "common.cuh"

namespace gl {

struct MT19937 {
    uint32_t mt[624];
    int idx;
    __device__ void seed(uint32_t s) {               // init_genrand (numpy legacy seeding of an int seed)
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    __device__ uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    __device__ double random_sample() {              // numpy legacy: 53-bit double in [0, 1)
        const uint32_t a = next() >> 5, b = next() >> 6;
        return __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
    }
};

struct ExactArgs {
    uint32_t seed;
    int N, d, iterations, burn_in, tuning_cycles, n_iterations;
    const double *loc, *scale;        // [d] uniform priors
    double cst;                       // ndim*const of the reference likelihood
    const double *weights;            // [n_iterations + 1]: w_1..w_n, then alpha**n
    double *live, *logl;              // scratch [N, d], [N]
    double *work;                     // scratch [8 d]
    double *dead;                     // out [n_iterations + N - 1, d + 2]
    double *widths_out;               // out [d]
    double *scalars_out;              // out: Z, H, n_evals, information
};

__device__ __forceinline__ double exact_loglike(const double *x, int d, double cst) {
    double s = -0.0;                                  // numpy pairwise sum for n < 8 starts from -0.0
    for (int j = 0; j < d; ++j) s = __dadd_rn(s, __dmul_rn(x[j], x[j]));
    return __dadd_rn(__dmul_rn(-0.5, s), cst);
}

// scipy.stats.uniform(loc, scale).pdf(x) / cdf(inf)
__device__ __forceinline__ double exact_pdf(double x, double loc, double scale) {
    if (x != x) return x;
    const double y = __ddiv_rn(__dadd_rn(x, -loc), scale);
    return (y >= 0.0 && y <= 1.0) ? __ddiv_rn(__ddiv_rn(1.0, scale), 1.0) : 0.0;
}

__global__ void reference_exact_kernel(ExactArgs A) {
    __shared__ MT19937 rng;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int N = A.N, d = A.d;
    double *widths = A.work, *steps = A.work + d, *acc = A.work + 2 * d, *cur = A.work + 3 * d;
    double *trial = A.work + 4 * d, *rs = A.work + 5 * d, *us = A.work + 6 * d, *start = A.work + 7 * d;
    rng.seed(A.seed);
    // ---- initial population: rvs(1) per parameter, row-major (nested_sampling.py:93-101) -----------
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < d; ++j) A.live[i * d + j] = __dadd_rn(__dmul_rn(rng.random_sample(), A.scale[j]), A.loc[j]);
    for (int i = 0; i < N; ++i) A.logl[i] = exact_loglike(A.live + i * d, d, A.cst);
    double n_evals = N;
    double Z = 0.0, H = 0.0, info = 0.0;
    bool first_call = true;
    long long row = 0;
    const int W = d + 2;
    int n = 1;
    auto argmin = [&]() {
        int best = 0;
        for (int i = 1; i < N; ++i)
            if (A.logl[i] < A.logl[best]) best = i;
        return best;
    };
    int ndx = argmin();
    while (true) {
        const double ll = A.logl[ndx], w = A.weights[n - 1];
        const double dZ = w * exp(ll);
        Z += dZ;
        const double dH = dZ * ll;
        if (dH == dH) H += dH;
        if (Z > 0.0) info = -log(Z) + H / Z;
        A.dead[row * W] = ll;
        A.dead[row * W + 1] = w;
        for (int j = 0; j < d; ++j) A.dead[row * W + 2 + j] = A.live[ndx * d + j];
        ++row;
        if (n >= A.n_iterations) break;
        ++n;
        // ---- survivor pick (nested_sampling.py:147-149) ---------------------------------------------
        int pick = (int)__dmul_rn(rng.random_sample(), (double)N);
        while (pick == ndx) pick = (int)__dmul_rn(rng.random_sample(), (double)N);
        for (int j = 0; j < d; ++j) start[j] = A.live[pick * d + j];
        const double boundary = ll;
        // ---- the shipped sampler (samplers.py:56-152) ---------------------------------------------------
        if (first_call) {                             // :70-81: 0.5*(max-min) of 100 draws per parameter
            for (int j = 0; j < d; ++j) {
                double lo = INFINITY, hi = -INFINITY;
                for (int q = 0; q < 100; ++q) {
                    const double v = __dadd_rn(__dmul_rn(rng.random_sample(), A.scale[j]), A.loc[j]);
                    lo = v < lo ? v : lo;
                    hi = v > hi ? v : hi;
                }
                widths[j] = __dmul_rn(0.5, __dadd_rn(hi, -lo));
            }
            first_call = false;
        }
        double cur_l = exact_loglike(start, d, A.cst);   // :83
        n_evals += 1;
        for (int j = 0; j < d; ++j) { steps[j] = widths[j]; acc[j] = 0.0; cur[j] = start[j]; }
        for (int cyc = 0; cyc < A.tuning_cycles; ++cyc) {
            for (int kk = 0; kk < 20; ++kk) {
                for (int j = 0; j < d; ++j) rs[j] = rng.random_sample();
                for (int j = 0; j < d; ++j) us[j] = rng.random_sample();
                for (int j = 0; j < d; ++j) {
                    for (int q = 0; q < d; ++q) trial[q] = cur[q];
                    const double xj = cur[j], wj = widths[j];             // widths, not steps (:97)
                    const double nj = __dadd_rn(__dmul_rn(rs[j], wj), __dadd_rn(xj, -__ddiv_rn(wj, 2.0)));
                    trial[j] = nj;
                    const double ratio = __ddiv_rn(exact_pdf(nj, A.loc[j], A.scale[j]), exact_pdf(xj, A.loc[j], A.scale[j]));
                    const double tl = exact_loglike(trial, d, A.cst);
                    n_evals += 1;
                    if (us[j] < ratio && tl > boundary) { cur[j] = nj; cur_l = tl; acc[j] += 1.0; }
                }
                for (int j = 0; j < d; ++j) {                             // Q2: inside the k loop (:115-121)
                    const double rate = __ddiv_rn(acc[j], 20.0);
                    if (rate < 0.2) steps[j] = __dmul_rn(steps[j], 0.66);
                    if (rate > 0.6) steps[j] = __dmul_rn(steps[j], 1.33);
                    acc[j] = 0.0;
                }
            }
        }
        for (int j = 0; j < d; ++j) { widths[j] = steps[j]; cur[j] = start[j]; }   // :124-125
        for (int it = 0; it < A.iterations + A.burn_in; ++it) {
            for (int j = 0; j < d; ++j) rs[j] = rng.random_sample();
            for (int j = 0; j < d; ++j) us[j] = rng.random_sample();
            for (int j = 0; j < d; ++j) {
                for (int q = 0; q < d; ++q) trial[q] = cur[q];
                const double xj = cur[j], wj = widths[j];
                const double nj = __dadd_rn(__dmul_rn(rs[j], wj), __dadd_rn(xj, -__ddiv_rn(wj, 2.0)));
                // Q1: trial[j] is never set (:131-144): prior and likelihood see the current point
                const double ratio = __ddiv_rn(exact_pdf(trial[j], A.loc[j], A.scale[j]), exact_pdf(xj, A.loc[j], A.scale[j]));
                const double tl = exact_loglike(trial, d, A.cst);
                n_evals += 1;
                if (us[j] < ratio && tl > boundary) { cur[j] = nj; cur_l = tl; }
            }
        }
        A.logl[ndx] = cur_l;                                               // :153-154
        for (int j = 0; j < d; ++j) A.live[ndx * d + j] = cur[j];
        ndx = argmin();                                                    // :156
    }
    // ---- final live-point term and survivor rows (nested_sampling.py:184-203) ------------------------
    const double wf = A.weights[A.n_iterations];
    double lsum = -0.0;
    int n_surv = 0;
    for (int i = 0; i < N; ++i)
        if (i != ndx) { lsum += exp(A.logl[i]); ++n_surv; }
    const double lm = lsum / n_surv;
    Z += wf * lm;
    const double dH = wf * lm * log(lm);
    if (dH == dH) H += dH;
    if (Z > 0.0) info = -log(Z) + H / Z;
    const double aw = __ddiv_rn(wf, (double)n_surv);
    for (int i = 0; i < N; ++i) {
        if (i == ndx) continue;
        A.dead[row * W] = A.logl[i];
        A.dead[row * W + 1] = aw;
        for (int j = 0; j < d; ++j) A.dead[row * W + 2 + j] = A.live[i * d + j];
        ++row;
    }
    for (int j = 0; j < d; ++j) A.widths_out[j] = widths[j];
    A.scalars_out[0] = Z;
    A.scalars_out[1] = H;
    A.scalars_out[2] = n_evals;
    A.scalars_out[3] = info;
}

}  // namespace gl

extern "C" int gl_reference_exact_run_host(uint32_t seed, int32_t N, int32_t d, const double *loc_host,
                                           const double *scale_host, double cst, int32_t iterations, int32_t burn_in,
                                           int32_t tuning_cycles, int32_t n_iterations, const double *weights_host,
                                           double *dead_host, double *widths_host, double *scalars_host) {
    GL_REQUIRE(N >= 2 && d >= 1 && d <= GL_MAX_DIM && n_iterations >= 1, "bad sizes");
    GL_REQUIRE(loc_host && scale_host && weights_host && dead_host && widths_host && scalars_host, "NULL argument");
    const size_t rows = (size_t)n_iterations + N - 1, W = (size_t)d + 2;
    const size_t n_dbl = 2 * (size_t)d + (size_t)n_iterations + 1 + (size_t)N * d + N + 8 * (size_t)d + rows * W + d + 4;
    double *buf = nullptr;
    GL_CUDA(cudaMalloc(&buf, sizeof(double) * n_dbl));
    double *p = buf;
    gl::ExactArgs A;
    A.seed = seed; A.N = N; A.d = d; A.iterations = iterations; A.burn_in = burn_in;
    A.tuning_cycles = tuning_cycles; A.n_iterations = n_iterations; A.cst = cst;
    double *loc = p; p += d;
    double *scale = p; p += d;
    double *weights = p; p += n_iterations + 1;
    A.loc = loc; A.scale = scale; A.weights = weights;
    A.live = p; p += (size_t)N * d;
    A.logl = p; p += N;
    A.work = p; p += 8 * (size_t)d;
    A.dead = p; p += rows * W;
    A.widths_out = p; p += d;
    A.scalars_out = p; p += 4;
    cudaError_t e = cudaMemcpy(loc, loc_host, sizeof(double) * d, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(scale, scale_host, sizeof(double) * d, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(weights, weights_host, sizeof(double) * (n_iterations + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        gl::reference_exact_kernel<<<1, 32>>>(A);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(dead_host, A.dead, sizeof(double) * rows * W, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(widths_host, A.widths_out, sizeof(double) * d, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(scalars_host, A.scalars_out, sizeof(double) * 4, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (e != cudaSuccess) return gl::cuda_fail(e, "reference-exact run", __FILE__, __LINE__);
    return GL_OK;
}
