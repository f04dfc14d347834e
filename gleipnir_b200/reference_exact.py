"""Reference-exact validation mode: reproduce upstream Gleipnir's built-in Nested Sampling AS
SHIPPED -- numpy's global MT19937 stream, the reference's draw order, sampler quirks included --
bit for bit on the GPU (``csrc/reference_exact.cu``), for scipy ``uniform`` priors, the 'uniform'
proposal and the iid-Gaussian likelihood of the reference's own test-suite.  Strictly sequential;
meant for validating a port against a reference run, not for speed.

    out = run_reference_exact(seed=0, locs=[-5]*5, scales=[10]*5, const=5*(-0.5*np.log(2*np.pi)),
                              population_size=20, iterations=10, tuning_cycles=1, n_iterations=120)
    out["log_evidence"], out["dead"]      # == upstream run after np.random.seed(0)
"""
import numpy as np

from . import _lib


def run_reference_exact(seed, locs, scales, const, population_size, iterations=100, burn_in=0, tuning_cycles=0,
                        n_iterations=1000):
    lib = _lib.load()
    d, N, n = len(locs), int(population_size), int(n_iterations)
    alpha = N / (N + 1)
    # the reference's own expressions (nested_sampling.py:112,143,184), evaluated by CPython floats
    w = [1.0 - alpha ** 1]
    for it in range(2, n + 1):
        w.append(alpha ** (it - 1.0) - alpha ** it)
    w.append(alpha ** n)
    weights = np.array(w, dtype=np.float64)
    loc = np.ascontiguousarray(locs, dtype=np.float64)
    scale = np.ascontiguousarray(scales, dtype=np.float64)
    dead = np.zeros((n + N - 1, d + 2))
    widths = np.zeros(d)
    scal = np.zeros(4)
    _lib.check(lib.gl_reference_exact_run_host(int(seed) & 0xFFFFFFFF, N, d, _lib.as_double_ptr(loc),
                                               _lib.as_double_ptr(scale), float(const), int(iterations),
                                               int(burn_in), int(tuning_cycles), n, _lib.as_double_ptr(weights),
                                               _lib.as_double_ptr(dead), _lib.as_double_ptr(widths),
                                               _lib.as_double_ptr(scal)))
    Z, H, n_evals, info = scal
    return dict(dead=dead, widths=widths, evidence=Z, H=H, information=info, log_evidence=float(np.log(Z)),
                log_evidence_error=float(np.sqrt(info / N)), n_evals=int(n_evals))
