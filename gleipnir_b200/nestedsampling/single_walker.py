"""The sampler's per-call protocol on the device: one walker through ``gl_sampler_call_host``."""
import ctypes

import numpy as np

from .. import _lib
from ..likelihoods import require_device_loglikelihood
from ..sampled_parameter import prior_table


def run_single_walker(sampler, sampled_parameters, loglikelihood, start, ns_boundary, seed=None):
    lib = _lib.load()
    like = require_device_loglikelihood(loglikelihood)
    table = prior_table(sampled_parameters)
    d = len(table)
    cfg = sampler.device_config()
    if seed is None:
        seed = getattr(sampler, "_seed", None)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        sampler._seed = seed
    calls = getattr(sampler, "_calls", 0)
    have = 0 if sampler._first else 1
    widths = np.zeros(d) if sampler._first else np.ascontiguousarray(sampler._widths, dtype=np.float64).copy()
    start = np.ascontiguousarray(start, dtype=np.float64)
    point = np.zeros(d)
    logl = ctypes.c_double(0.0)
    _lib.check(lib.gl_sampler_call_host(_lib.prior_array(table), d, like.device_model(), cfg["iterations"],
                                        cfg["burn_in"], cfg["tuning_cycles"], cfg["proposal"], int(seed), calls,
                                        _lib.as_double_ptr(start), float(ns_boundary), _lib.as_double_ptr(widths), have,
                                        _lib.as_double_ptr(point), ctypes.byref(logl)))
    sampler._calls = calls + 1
    return point, logl.value, widths
