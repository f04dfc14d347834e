// ODE-network log-likelihood functor: per-parameter-set stiff integration with the data
// misfit fused in.  B200 form of the PySB likelihood closures
// (gleipnir/pysb_utilities/nestedsample_it.py:273-339;
//  examples/pysb_dimerization_model/run_NS_dimerization_model.py:56-65):
//     params[rate_mask] = 10**Y  ->  ScipyOdeSimulator.run  ->  sum over observables x masked
//     time points of norm(data, sd).logpdf(sim) | -mean((data-sim)^2) | -sum((data-sim)^2);  NaN -> -inf.
//
// `Net` is a code-generated struct (gleipnir_b200/odegen.py) with compile-time sizes and
// straight-line mass-action right-hand side / analytic Jacobian:
//     static constexpr int NS, NP, NOBS;
//     static __device__ void rhs(const double *y, const double *p, double *f);
//     static __device__ void jac(const double *y, const double *p, double *J);   // row-major NS x NS
//     static __device__ void initials(const double *p, double *y0);
//     static __device__ double observable(int o, const double *y);
//
// Integrator: Kaps-Rentrop 4(3) Rosenbrock method (Shampine's parameter set, autonomous form):
// linearly implicit (one LU of 1/(gamma*h) I - J and four back-substitutions per step),
// A-stable, embedded third-order error estimate, standard step-size control; steps are
// clipped to land exactly on the output times so observables are read without interpolation.
// One system per thread: all state is thread-private (registers for small NS).  A step cap
// and any non-finite value end the solve with logL = -inf (the reference's NaN -> -inf and
// try/except -> -inf conventions, run_jarm_multinest.py:115-119,137-143).
#pragma once
#include "common.cuh"
#include "likelihoods.cuh"

namespace gl {

// blob layout (float64 words), written by gleipnir_b200/odegen.py::pack_blob
enum OdeBlob {
    OB_NSAMPLED = 0, OB_NPARAMS = 1, OB_NTIMES = 2, OB_LIKE = 3, OB_MAXSTEPS = 4, OB_RTOL = 5, OB_ATOL = 6,
    OB_NROWS = 7, OB_OFF_NOMINAL = 16, OB_OFF_MAP = 17, OB_OFF_ISLOG = 18, OB_OFF_TSPAN = 19, OB_OFF_OBSIDX = 20,
    OB_OFF_DATA = 21, OB_OFF_SD = 22, OB_OFF_MASK = 23, OB_HEADER = 32
};
#ifndef GL_ODE_LIKE_TYPES
#define GL_ODE_LIKE_TYPES
enum OdeLikeType { ODE_SNLPDF = 0, ODE_MSE = 1, ODE_SSE = 2 };
#endif

// In-place LU with partial pivoting of an N x N row-major matrix; returns false if singular.
template <int N>
__device__ __forceinline__ bool lu_factor(double *A, int *piv) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        int p = c;
        double best = fabs(A[c * N + c]);
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            const double v = fabs(A[r * N + c]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
        if (!(best > 0.0)) return false;
        if (p != c) {
#pragma unroll
            for (int q = 0; q < N; ++q) {
                const double t = A[c * N + q];
                A[c * N + q] = A[p * N + q];
                A[p * N + q] = t;
            }
        }
        const double inv = 1.0 / A[c * N + c];
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            const double l = A[r * N + c] * inv;
            A[r * N + c] = l;
#pragma unroll
            for (int q = c + 1; q < N; ++q) A[r * N + q] = fma(-l, A[c * N + q], A[r * N + q]);
        }
    }
    return true;
}

template <int N>
__device__ __forceinline__ void lu_solve(const double *A, const int *piv, double *b) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const int p = piv[c];
        if (p != c) { const double t = b[c]; b[c] = b[p]; b[p] = t; }
#pragma unroll
        for (int r = c + 1; r < N; ++r) b[r] = fma(-A[r * N + c], b[c], b[r]);
    }
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        double s = b[c];
#pragma unroll
        for (int q = c + 1; q < N; ++q) s = fma(-A[c * N + q], b[q], s);
        b[c] = s / A[c * N + c];
    }
}

// One Kaps-Rentrop step of size h from y; writes ynew and the error estimate.  false = singular.
template <class Net>
__device__ __forceinline__ bool ros4_step(const double *y, const double *p, double h, double *ynew, double *yerr) {
    constexpr int N = Net::NS;
    constexpr double GAM = 0.5, A21 = 2.0, A31 = 48.0 / 25.0, A32 = 6.0 / 25.0;
    constexpr double C21 = -8.0, C31 = 372.0 / 25.0, C32 = 12.0 / 5.0, C41 = -112.0 / 125.0, C42 = -54.0 / 125.0,
                     C43 = -2.0 / 5.0;
    constexpr double B1 = 19.0 / 9.0, B2 = 0.5, B3 = 25.0 / 108.0, B4 = 125.0 / 108.0;
    constexpr double E1 = 17.0 / 54.0, E2 = 7.0 / 36.0, E4 = 125.0 / 108.0;
    double A[N * N], g1[N], g2[N], g3[N], g4[N], yt[N];
    int piv[N];
    Net::jac(y, p, A);
    const double diag = 1.0 / (GAM * h), ih = 1.0 / h;
#pragma unroll
    for (int i = 0; i < N * N; ++i) A[i] = -A[i];
#pragma unroll
    for (int i = 0; i < N; ++i) A[i * N + i] += diag;
    if (!lu_factor<N>(A, piv)) return false;
    Net::rhs(y, p, g1);
    lu_solve<N>(A, piv, g1);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = fma(A21, g1[i], y[i]);
    Net::rhs(yt, p, g2);
#pragma unroll
    for (int i = 0; i < N; ++i) g2[i] = fma(C21 * ih, g1[i], g2[i]);
    lu_solve<N>(A, piv, g2);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = fma(A32, g2[i], fma(A31, g1[i], y[i]));
    Net::rhs(yt, p, g3);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        g4[i] = g3[i];                                // stage 4 reuses f(y3)
        g3[i] = fma(C32 * ih, g2[i], fma(C31 * ih, g1[i], g3[i]));
    }
    lu_solve<N>(A, piv, g3);
#pragma unroll
    for (int i = 0; i < N; ++i) g4[i] = fma(C43 * ih, g3[i], fma(C42 * ih, g2[i], fma(C41 * ih, g1[i], g4[i])));
    lu_solve<N>(A, piv, g4);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ynew[i] = y[i] + B1 * g1[i] + B2 * g2[i] + B3 * g3[i] + B4 * g4[i];
        yerr[i] = E1 * g1[i] + E2 * g2[i] + E4 * g4[i];
    }
    return true;
}

template <class Net>
struct LikeOde {
    const double *blob;
    int d;
    __device__ explicit LikeOde(const ModelDev &m) : blob(m.aux0), d(m.d) {}
    static __host__ __device__ int cache_size(int d) { return 0; }

    // logL for the sampled vector Y[j*stride] with component jrep replaced by v (jrep < 0: none)
    __device__ double solve(const double *Y, int stride, int jrep, double v) const {
        constexpr int N = Net::NS;
        const double *B = blob;
        const int np = (int)B[OB_NPARAMS], nt = (int)B[OB_NTIMES], like = (int)B[OB_LIKE];
        const int max_steps = (int)B[OB_MAXSTEPS], nrows = (int)B[OB_NROWS];
        const double rtol = B[OB_RTOL], atol = B[OB_ATOL];
        const double *nominal = B + (int)B[OB_OFF_NOMINAL], *map = B + (int)B[OB_OFF_MAP];
        const double *islog = B + (int)B[OB_OFF_ISLOG], *tspan = B + (int)B[OB_OFF_TSPAN];
        const double *obsidx = B + (int)B[OB_OFF_OBSIDX], *data = B + (int)B[OB_OFF_DATA];
        const double *sd = B + (int)B[OB_OFF_SD], *mask = B + (int)B[OB_OFF_MASK];
        double p[Net::NP];
        for (int i = 0; i < np; ++i) p[i] = nominal[i];
        for (int j = 0; j < d; ++j) {
            const double yj = j == jrep ? v : Y[j * stride];
            p[(int)map[j]] = islog[j] != 0.0 ? exp10(yj) : yj;           // params[mask] = 10**Y
        }
        double y[N], ynew[N], yerr[N];
        Net::initials(p, y);
        double acc = 0.0;                      // snlpdf / sse: running sum; mse: sum of per-row means
        // per-row accumulators for mse are folded on the fly: row r contributes sum_r / count_r
        auto misfit = [&](int it) {
            for (int r = 0; r < nrows; ++r) {
                if (mask[r * nt + it] == 0.0) continue;
                const double sim = Net::observable((int)obsidx[r], y);
                const double mu = data[r * nt + it];
                if (like == ODE_SNLPDF) {
                    const double s = sd[r * nt + it], z = (sim - mu) / s;
                    acc += -0.5 * z * z - log(s) - GL_HALF_LOG_2PI;
                } else if (like == ODE_SSE) {
                    const double e = mu - sim;
                    acc -= e * e;
                } else {
                    const double e = mu - sim;
                    acc -= e * e * sd[r * nt + it];       // mse: sd slot holds 1/count of the row
                }
            }
        };
        misfit(0);
        double t = tspan[0];
        double h = fmin(1e-3 * (tspan[nt - 1] - tspan[0]), tspan[1] - tspan[0]);
        int steps = 0;
        for (int it = 1; it < nt; ++it) {
            const double tn = tspan[it];
            while (t < tn) {
                const double rem = tn - t;
                const bool last = h >= rem;
                const double hh = last ? rem : h;
                if (++steps > max_steps) return -INFINITY;
                if (!ros4_step<Net>(y, p, hh, ynew, yerr)) return -INFINITY;
                double err = 0.0;
                bool finite = true;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double sc = atol + rtol * fmax(fabs(y[i]), fabs(ynew[i]));
                    err = fmax(err, fabs(yerr[i]) / sc);
                    finite = finite && isfinite(ynew[i]);
                }
                if (!finite || !(err == err)) {
                    if (hh < 1e-300) return -INFINITY;
                    h = hh * 0.1;
                    continue;
                }
                const double fac = err > 0.0 ? fmin(5.0, fmax(0.2, 0.9 * pow(err, -0.25))) : 5.0;
                if (err <= 1.0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) y[i] = ynew[i];
                    t = last ? tn : t + hh;
                    // a step clipped to land on an output time only lowers the running step size
                    // if its own error asks for that; otherwise the unclipped proposal stands
                    if (!last || fac < 1.0) h = hh * fac;
                    else h = fmax(h, hh * fac);
                } else {
                    h = hh * fac;
                    if (h < 1e-14 * fabs(tn)) return -INFINITY;
                }
            }
            misfit(it);
        }
        return (acc != acc) ? -INFINITY : acc;
    }

    __device__ __forceinline__ double eval(const double *x, int stride) const { return solve(x, stride, -1, 0.0); }
    __device__ __forceinline__ double init(const double *x, double *c, int stride, int lane, int G,
                                           unsigned gmask) const {
        return solve(x, stride, -1, 0.0);
    }
    __device__ __forceinline__ double trial(const double *x, const double *c, int stride, int j, double v,
                                            double &cnew) const {
        cnew = 0.0;
        return solve(x, stride, j, v);
    }
    __device__ __forceinline__ void commit(double *x, double *c, int stride, int j, double v, double cnew, int lane,
                                           int G, unsigned gmask) const {
        if (lane == 0) x[j * stride] = v;
    }
    __device__ __forceinline__ double refresh(const double *x, double *c, int stride, int lane, int G, unsigned gmask,
                                              double L) const {
        return L;
    }
};

}  // namespace gl
