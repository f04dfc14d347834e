// Closed-form prior pdf / ppf for the three recognised scipy.stats families
// (SampledParameter, gleipnir/sampled_parameter.py:40-80).  Arithmetic forms match
// oracle/priors.py.
#pragma once
#include "common.cuh"

namespace gl {

struct PriorTable {   // device copy, SoA
    int32_t kind[GL_MAX_DIM];
    double p0[GL_MAX_DIM];
    double p1[GL_MAX_DIM];
};

#define GL_LOG_SQRT_2PI 0.91893853320467274178

__device__ __forceinline__ double prior_pdf(int kind, double p0, double p1, double x) {
    if (kind == GL_PRIOR_UNIFORM) {
        if (x != x) return x;
        double y = (x - p0) / p1;
        return (y >= 0.0 && y <= 1.0) ? 1.0 / p1 : 0.0;
    } else if (kind == GL_PRIOR_NORM) {
        double z = (x - p0) / p1;
        return exp(-0.5 * z * z - GL_LOG_SQRT_2PI) / p1;
    } else {
        if (x != x) return x;
        return (x >= p0 && x <= p1) ? 1.0 / (x * log(p1 / p0)) : 0.0;
    }
}

__device__ __forceinline__ double prior_ppf(int kind, double p0, double p1, double u) {
    if (kind == GL_PRIOR_UNIFORM) {
        return u * p1 + p0;
    } else if (kind == GL_PRIOR_NORM) {
        return normcdfinv(u) * p1 + p0;
    } else {
        return p0 * exp(u * log(p1 / p0));
    }
}

// Metropolis prior test of one component move (samplers.py:105-107,110): returns
// (u < prior(new)/prior(cur)) with the ratio simplified algebraically per family, so that no
// division or exp is spent where the answer is already decided:
//   uniform     ratio in {1, 0, inf, nan}            ->  inside(new)
//   norm        exp(-0.5*(zn^2 - zc^2))              ->  true when zn^2 <= zc^2 (ratio >= 1 > u)
//   loguniform  cur/new inside the support            ->  inside(new) && (!inside(cur) || u < cur/new)
// inside(x) := 0 <= x - p0 <= p1 (uniform) / p0 <= x <= p1 (loguniform); NaN is outside.
// Same forms as oracle/priors.py::accept_prior.
__device__ __forceinline__ bool prior_accept(int kind, double p0, double p1, double cur, double nw, double u) {
    if (kind == GL_PRIOR_UNIFORM) {
        const double t = nw - p0;
        return t >= 0.0 && t <= p1;
    } else if (kind == GL_PRIOR_NORM) {
        const double zn = (nw - p0) / p1, zc = (cur - p0) / p1;
        const double a = zn * zn, b = zc * zc;
        if (a <= b) return true;
        return u < exp(-0.5 * (a - b));
    } else {
        const bool in_n = nw >= p0 && nw <= p1;
        if (!in_n) return false;
        const bool in_c = cur >= p0 && cur <= p1;
        return !in_c || u < cur / nw;
    }
}

}  // namespace gl
