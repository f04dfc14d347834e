"""``SampledParameter`` -- host mirror of ``gleipnir/sampled_parameter.py:21-80``.

Same constructor, attributes (``name``, ``prior_dist``, ``pf``, ``logpf``, ``_norm``) and
scalar host methods as the reference: those simply forward to the user's prior object, as
upstream does, so any distribution object keeps working for scripts that only use them.

What is new is the batched device side used by the Nested Sampling engine: priors that
are scipy.stats frozen ``uniform``, ``norm`` or ``loguniform``/``reciprocal`` are parsed
into a ``(kind, p0, p1)`` row of the device prior table, and ``device_rvs`` /
``device_prior`` / ``device_logprior`` / ``device_invcdf`` run the Philox draw and the
pdf / ppf transforms as CUDA kernels on whole arrays.  Any other prior family raises when
the engine asks for its table row (no CPU fallback).
"""
import numpy as np

from . import _lib


def prior_table_row(prior):
    """scipy.stats frozen distribution -> (kind, p0, p1) of the device prior table."""
    dist = getattr(prior, "dist", None)
    name = getattr(dist, "name", None)
    if name is None or not hasattr(dist, "_parse_args"):
        raise TypeError("prior %r is not a scipy.stats frozen distribution; the device path supports "
                        "uniform, norm and loguniform/reciprocal only" % (prior,))
    shapes, loc, scale = dist._parse_args(*prior.args, **prior.kwds)
    loc, scale = float(loc), float(scale)
    if name == "uniform":
        return (_lib.GL_PRIOR_UNIFORM, loc, scale)
    if name == "norm":
        return (_lib.GL_PRIOR_NORM, loc, scale)
    if name in ("loguniform", "reciprocal"):
        if loc != 0.0 or scale != 1.0:
            raise ValueError("loguniform priors with loc/scale are not supported on the device path")
        a, b = (float(s) for s in shapes)
        return (_lib.GL_PRIOR_LOGUNIFORM, a, b)
    raise TypeError("prior family %r has no device kernel (supported: uniform, norm, loguniform/reciprocal)" % name)


class SampledParameter(object):
    """A named parameter with a prior.  See the module docstring."""

    def __init__(self, name, prior):
        self.name = name
        self.prior_dist = prior
        if hasattr(prior, "pdf"):
            self.pf, self.logpf = prior.pdf, prior.logpdf
        else:
            self.pf, self.logpf = prior.pmf, prior.logpmf
        self._norm = prior.cdf(np.inf)
        self._row = None

    # ---- scalar / numpy host protocol (forwards to the prior object, as upstream) ------
    def rvs(self, sample_shape):
        return self.prior_dist.rvs(sample_shape)

    def logprior(self, value):
        return np.log(self.pf(value) / self._norm)

    def prior(self, value):
        return self.pf(value) / self._norm

    def invcdf(self, value):
        return self.prior_dist.ppf(value)

    # ---- batched device protocol -----------------------------------------------------------
    def table_row(self):
        if self._row is None:
            self._row = prior_table_row(self.prior_dist)
        return self._row

    def _map(self, fn_name, values):
        import torch
        lib = _lib.load()
        x = torch.as_tensor(values, dtype=torch.float64, device="cuda").contiguous()
        out = torch.empty_like(x)
        row = _lib.prior_array([self.table_row()])
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(getattr(lib, fn_name)(row, x.data_ptr(), out.data_ptr(), x.numel(), stream))
        return out

    def device_prior(self, values):
        """pdf on a whole array (CUDA tensor in, CUDA tensor out)."""
        return self._map("gl_prior_pdf", values)

    def device_logprior(self, values):
        return self._map("gl_prior_logpdf", values)

    def device_invcdf(self, values):
        return self._map("gl_prior_ppf", values)

    def device_rvs(self, n, seed=0, first_row=0):
        """n Philox-keyed prior draws as a CUDA tensor (stream tag GL_TAG_USER)."""
        import torch
        lib = _lib.load()
        out = torch.empty(int(n), dtype=torch.float64, device="cuda")
        row = _lib.prior_array([self.table_row()])
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.gl_prior_draw(row, 1, int(seed), _lib.GL_TAG_USER, int(first_row), int(n),
                                     out.data_ptr(), stream))
        return out


def prior_table(sampled_parameters):
    """List of SampledParameter -> list of (kind, p0, p1) rows, in order."""
    return [sp.table_row() for sp in sampled_parameters]


def hypercube_transform(sampled_parameters, unit_cube):
    """Batched ``invcdf`` over a whole [n, d] block of unit-cube points (the transform the
    third-party wrappers build from ``SampledParameter.invcdf``, e.g. multinest.py:131-132)."""
    import torch
    lib = _lib.load()
    u = torch.as_tensor(unit_cube, dtype=torch.float64, device="cuda").contiguous()
    d = len(sampled_parameters)
    if u.dim() != 2 or u.shape[1] != d:
        raise ValueError("unit_cube must be [n, %d]" % d)
    out = torch.empty_like(u)
    rows = _lib.prior_array(prior_table(sampled_parameters))
    _lib.check(lib.gl_prior_transform(rows, d, u.data_ptr(), out.data_ptr(), u.shape[0],
                                      torch.cuda.current_stream().cuda_stream))
    return out
