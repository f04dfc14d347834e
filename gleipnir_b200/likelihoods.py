"""Device log-likelihood objects: the batched device protocol.

The reference's protocol is a Python callable ``loglikelihood(ndarray[d]) -> float``
(``nested_sampling.py:108``, ``samplers.py:83,108,144``).  A bare Python callable cannot
run inside a CUDA kernel, so the B200 engine accepts objects of the classes below: each
owns a device model handle (``gl_model``) that the walker and batch-evaluation kernels
dispatch on, and is *still host-callable on one vector* (``deviance_ic`` and user code
call it that way) -- that call is evaluated by the same CUDA kernel with a batch of one.
"""
import ctypes

import numpy as np

from . import _lib


def _current_device():
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else None
    except ImportError:          # pragma: no cover
        return None


class DeviceLogLikelihood(object):
    """Base class: owns a ``gl_model`` handle created lazily on first use."""

    ndim = None

    def __init__(self):
        self._handle = None

    def _create(self, lib):          # pragma: no cover - abstract
        raise NotImplementedError

    def device_model(self):
        """The model's device data (dense matrices, ODE blobs) lives on the device that is current at the
        first use; using the object from another device is an error, not a silent cross-device access."""
        if self._handle is None:
            lib = _lib.load()
            self._handle = self._create(lib)
            self._device_index = _current_device()
        elif getattr(self, "_device_index", None) not in (None, _current_device()):
            raise _lib.GleipnirB200Error("this likelihood object was created on cuda:%s and is being used with cuda:%s "
                                         "current" % (self._device_index, _current_device()))
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and _lib is not None and getattr(_lib, "_lib", None) is not None:
            try:
                _lib._lib.gl_model_destroy(h)
            except Exception:
                pass
            self._handle = None

    def evaluate(self, x):
        """Batched evaluation.  CUDA tensor [n, d] -> CUDA tensor [n]; numpy -> numpy."""
        lib = _lib.load()
        try:
            import torch
            is_tensor = isinstance(x, torch.Tensor)
        except ImportError:          # pragma: no cover
            is_tensor = False
        if is_tensor:
            dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
            with torch.cuda.device(dev):
                xt = x.to(dtype=torch.float64, device=dev).contiguous().reshape(-1, self.ndim)
                out = torch.empty(xt.shape[0], dtype=torch.float64, device=dev)
                _lib.check(lib.gl_loglike_eval(self.device_model(), xt.data_ptr(), xt.shape[0], out.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
            return out
        xa = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, self.ndim))
        out = np.empty(xa.shape[0], dtype=np.float64)
        _lib.check(lib.gl_loglike_eval_host(self.device_model(), _lib.as_double_ptr(xa), xa.shape[0],
                                            _lib.as_double_ptr(out)))
        return out

    def __call__(self, sampled_parameter_vector):
        return float(self.evaluate(np.asarray(sampled_parameter_vector, dtype=np.float64))[0])


class IIDGaussianLogLikelihood(DeviceLogLikelihood):
    """``-0.5*sum(((x-mu)/sigma)**2) + ndim*(-0.5*log(2*pi) - log(sigma))``
    (reference: tests/test_nested_sampling.py:34-36 with mu=0, sigma=1)."""

    def __init__(self, ndim, mu=0.0, sigma=1.0):
        super().__init__()
        self.ndim, self.mu, self.sigma = int(ndim), float(mu), float(sigma)

    def _create(self, lib):
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_iid_gaussian(self.ndim, self.mu, self.sigma, ctypes.byref(h)))
        return h


class EggCartonLogLikelihood(DeviceLogLikelihood):
    """``(2 + prod(cos(x)))**5`` (reference: examples/eggcarton/NS_eggcarton.py:25-27)."""

    def __init__(self, ndim=2):
        super().__init__()
        self.ndim = int(ndim)

    def _create(self, lib):
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_eggcarton(self.ndim, ctypes.byref(h)))
        return h


class Mixture1DLogLikelihood(DeviceLogLikelihood):
    """log of the mean of equal-width 1-D Gaussians, NaN -> -inf
    (reference: examples/1d_multimodal/NS_1d_multimodal.py:20-32)."""

    def __init__(self, positions=(0.1, 0.2, 0.5, 0.55, 0.9, 1.1), width=0.01):
        super().__init__()
        self.ndim = 1
        self.positions = np.ascontiguousarray(positions, dtype=np.float64)
        self.width = float(width)

    def _create(self, lib):
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_mixture1d(_lib.as_double_ptr(self.positions), len(self.positions),
                                                 self.width, ctypes.byref(h)))
        return h


class DenseGaussianLogLikelihood(DeviceLogLikelihood):
    """``-0.5*(x-mu)^T P (x-mu) - 0.5*logdet(2*pi*Sigma)`` with a dense precision matrix
    (synthetic C3 config; not in the reference)."""

    def __init__(self, mu, precision, log_norm=None):
        super().__init__()
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.precision = np.ascontiguousarray(precision, dtype=np.float64)
        self.ndim = int(self.mu.shape[0])
        if self.precision.shape != (self.ndim, self.ndim):
            raise ValueError("precision must be [d, d]")
        if log_norm is None:
            sign, logdet = np.linalg.slogdet(self.precision)
            log_norm = -0.5 * (self.ndim * np.log(2.0 * np.pi) - logdet)
        self.log_norm = float(log_norm)

    @classmethod
    def synthetic_c3(cls, ndim=50, seed=1234, lam_lo=1e-2, lam_hi=1.0):
        """Sigma = Q diag(lam) Q^T, Q from QR of default_rng(seed).standard_normal, lam log-spaced."""
        rng = np.random.default_rng(seed)
        q, _ = np.linalg.qr(rng.standard_normal((ndim, ndim)))
        lam = np.logspace(np.log10(lam_lo), np.log10(lam_hi), ndim)
        p = (q / lam[None, :]) @ q.T
        p = 0.5 * (p + p.T)
        log_norm = -0.5 * (ndim * np.log(2.0 * np.pi) + np.sum(np.log(lam)))
        return cls(np.zeros(ndim), p, log_norm)

    def _create(self, lib):
        h = ctypes.c_void_p()
        _lib.check(lib.gl_model_create_dense_gaussian(self.ndim, _lib.as_double_ptr(self.mu),
                                                      _lib.as_double_ptr(self.precision), self.log_norm,
                                                      ctypes.byref(h)))
        return h


def require_device_loglikelihood(loglikelihood):
    if not isinstance(loglikelihood, DeviceLogLikelihood):
        raise TypeError(
            "loglikelihood must be a gleipnir_b200.likelihoods.DeviceLogLikelihood (e.g. "
            "IIDGaussianLogLikelihood, EggCartonLogLikelihood, Mixture1DLogLikelihood, "
            "DenseGaussianLogLikelihood, or an ODE-network likelihood): a bare Python callable "
            "cannot run inside the CUDA walker kernels and there is no CPU fallback")
    return loglikelihood
