"""TEST INFRASTRUCTURE ONLY (oracle) -- closed-form prior pdf / logpdf / ppf.

CPU restatement of what ``SampledParameter`` (``gleipnir/sampled_parameter.py:21-80``)
asks of a scipy.stats frozen distribution, for the three families the B200
path recognises.  ``scipy.stats`` 1.18.1 is the third-party arithmetic behind
the reference; the closed forms here are pinned against it and against the
known answers in the reference's ``tests/test_sampled_parameter.py:23-36`` by
``tests/test_oracle_priors.py``.

Prior table row = (kind, p0, p1):
    GL_PRIOR_UNIFORM    p0 = loc, p1 = scale     support [loc, loc+scale] (closed)
    GL_PRIOR_NORM       p0 = loc, p1 = scale
    GL_PRIOR_LOGUNIFORM p0 = a,   p1 = b         support [a, b]
"""
import numpy as np
from scipy.special import ndtri

GL_PRIOR_UNIFORM = 0
GL_PRIOR_NORM = 1
GL_PRIOR_LOGUNIFORM = 2

_LOG_SQRT_2PI = 0.91893853320467274178


def pdf(kind, p0, p1, x):
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(all="ignore"):
        if kind == GL_PRIOR_UNIFORM:
            # scipy uniform._pdf: 1.0*(x==x) on the closed support, 0 outside
            y = (x - p0) / p1
            return np.where((y >= 0.0) & (y <= 1.0), 1.0 / p1, 0.0) + 0.0 * np.where(np.isnan(x), np.nan, 0.0)
        if kind == GL_PRIOR_NORM:
            z = (x - p0) / p1
            return np.exp(-0.5 * z * z - _LOG_SQRT_2PI) / p1
        if kind == GL_PRIOR_LOGUNIFORM:
            inside = (x >= p0) & (x <= p1)
            return np.where(inside, 1.0 / (x * np.log(p1 / p0)), 0.0)
    raise ValueError(kind)


def logpdf(kind, p0, p1, x):
    # SampledParameter.logprior is log(pdf/_norm) (sampled_parameter.py:60)
    with np.errstate(all="ignore"):
        return np.log(pdf(kind, p0, p1, x))


def ppf(kind, p0, p1, u):
    u = np.asarray(u, dtype=np.float64)
    if kind == GL_PRIOR_UNIFORM:
        return u * p1 + p0                       # scipy: _ppf(q)*scale + loc with _ppf(q)=q
    if kind == GL_PRIOR_NORM:
        return ndtri(u) * p1 + p0
    if kind == GL_PRIOR_LOGUNIFORM:
        return p0 * np.exp(u * np.log(p1 / p0))  # a*(b/a)**q, scipy reciprocal._ppf
    raise ValueError(kind)


def accept_prior(kind, p0, p1, cur, new, u):
    """(u < prior(new)/prior(cur)) of samplers.py:105-107,110 with the ratio simplified per
    family (same forms as csrc/priors.cuh::prior_accept):
      uniform: the ratio is 1, 0, inf or nan -> accept iff new is inside the support;
      norm: exp(-0.5*(zn^2 - zc^2));  loguniform: cur/new inside the support."""
    cur = np.asarray(cur, dtype=np.float64)
    new = np.asarray(new, dtype=np.float64)
    with np.errstate(all="ignore"):
        if kind == GL_PRIOR_UNIFORM:
            t = new - p0
            return (t >= 0.0) & (t <= p1)
        if kind == GL_PRIOR_NORM:
            zn, zc = (new - p0) / p1, (cur - p0) / p1
            a, b = zn * zn, zc * zc
            return (a <= b) | (u < np.exp(-0.5 * (a - b)))
        if kind == GL_PRIOR_LOGUNIFORM:
            in_n = (new >= p0) & (new <= p1)
            in_c = (cur >= p0) & (cur <= p1)
            return in_n & (~in_c | (u < cur / new))
    raise ValueError(kind)


def parse_scipy_frozen(prior):
    """scipy frozen distribution -> (kind, p0, p1); mirrors the product-side parser."""
    name = prior.dist.name
    shapes, loc, scale = prior.dist._parse_args(*prior.args, **prior.kwds)
    if name == "uniform":
        return (GL_PRIOR_UNIFORM, float(loc), float(scale))
    if name == "norm":
        return (GL_PRIOR_NORM, float(loc), float(scale))
    if name in ("loguniform", "reciprocal"):
        a, b = shapes
        if float(loc) != 0.0 or float(scale) != 1.0:
            raise ValueError("loguniform with loc/scale is not supported")
        return (GL_PRIOR_LOGUNIFORM, float(a), float(b))
    raise ValueError("unsupported prior family: %s" % name)
