"""Device-side ``posteriors()``: weighted marginal histograms of a dead-point table
(reference: nested_sampling.py:266-293), via gl_posterior_weights / gl_posterior_histograms."""
import ctypes

import numpy as np

from . import _lib


def device_posteriors(log_l, weight, params, log_evidence, nbins=None, device=None):
    """log_l[rows], weight[rows], params[rows, d] (host arrays or CUDA tensors) ->
    (marginal[d, nbins], edges[d, nbins+1], centers[d, nbins]) as numpy arrays."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.GleipnirB200Error("no CUDA device is available and gleipnir_b200 has no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda" if device is None else device)
    f64 = dict(dtype=torch.float64, device=dev)
    def to_dev(a):
        if isinstance(a, torch.Tensor):
            return a.to(**f64).contiguous()
        return torch.from_numpy(np.array(a, dtype=np.float64, copy=True, order="C")).to(dev).contiguous()

    ll, w, x = to_dev(log_l), to_dev(weight), to_dev(params)
    rows, d = x.shape
    st = torch.cuda.current_stream(dev).cuda_stream
    nw = torch.empty(rows, **f64)
    stats = torch.empty(int(lib.gl_posterior_stats_bytes()), dtype=torch.uint8, device=dev)
    n_masked = ctypes.c_int64(0)
    _lib.check(lib.gl_posterior_weights(ll.data_ptr(), w.data_ptr(), x.data_ptr(), rows, d, float(log_evidence),
                                        nw.data_ptr(), stats.data_ptr(), ctypes.byref(n_masked), st))
    if nbins is None:
        nbins = 2 * int(np.cbrt(n_masked.value))               # Rice rule (nested_sampling.py:286)
    nbins = int(nbins)
    partial = torch.empty(int(lib.gl_posterior_partial_doubles(rows, d, nbins)), **f64)
    marg = torch.empty((d, nbins), **f64)
    edges = torch.empty((d, nbins + 1), **f64)
    centers = torch.empty((d, nbins), **f64)
    _lib.check(lib.gl_posterior_histograms(x.data_ptr(), nw.data_ptr(), rows, d, nbins, stats.data_ptr(),
                                           partial.data_ptr(), marg.data_ptr(), edges.data_ptr(), centers.data_ptr(),
                                           st))
    return marg.cpu().numpy(), edges.cpu().numpy(), centers.cpu().numpy()
