"""Concurrent independent Nested Sampling runs on one GPU ("replica batching").

A single small run (C1: 500 live points, C2: 10^4) offers the device a few tens of thousands of threads per
launch against the 3*10^5 it can hold, and spends its time in chains of short dependent launches.  Runs that
do not depend on each other -- the seeds of one problem (north_star asks for ln Z over 10 seeds), the models of a
selection sweep (``HypSelector.run_nested_sampling`` runs them one after another,
pysb_utilities/hyp_selector.py:236-248) -- therefore share the device: every run gets its own engine and CUDA
stream, ``gl_ns_advance_many`` enqueues all of them from one host thread each, and their kernels overlap.
Per-run results are bit-identical to running each object alone (same Philox counters, no shared state).
"""
import ctypes

import numpy as np

from . import _lib
from .engine import DeviceEngine
from .likelihoods import require_device_loglikelihood


def run_many(samplers, max_concurrent=16):
    """Run every ``NestedSampling`` object of ``samplers`` (closed-form stopping criteria, single GPU) and return
    ``[(log_evidence, log_evidence_error), ...]``; afterwards each object holds its results exactly as after its
    own ``run()``.  At most ``max_concurrent`` runs are resident at a time (their live sets and dead-point logs
    occupy device memory for the duration)."""
    import torch
    lib = _lib.load()
    samplers = list(samplers)
    out = [None] * len(samplers)
    for lo in range(0, len(samplers), int(max_concurrent)):
        group = samplers[lo:lo + int(max_concurrent)]
        prepared = []
        for ns in group:
            loglike = require_device_loglikelihood(ns.loglikelihood)
            loglike, table, n, k, seed, scfg, n_deaths, chunked, capacity = ns._prepare(loglike)
            if chunked:
                raise TypeError("run_many needs stopping criteria with a closed-form death count "
                                "(NumberOfIterations, RemainingPriorMass)")
            stream = torch.cuda.Stream(device=ns.device)
            eng = DeviceEngine(table, loglike, n, k, capacity, seed=seed, device=ns.device, stream=stream, **scfg)
            ns._engine = eng
            ns._adopt_initial(eng, loglike, 0, n)
            prepared.append((ns, eng, n, k, n_deaths, stream))
        torch.cuda.synchronize()                           # buffers allocated on the caller's stream are ready
        for ns, eng, n, k, n_deaths, stream in prepared:
            eng.init()
        m = len(prepared)
        handles = (ctypes.c_void_p * m)(*[p[1]._h for p in prepared])
        streams = (ctypes.c_void_p * m)(*[p[1]._stream() for p in prepared])
        deaths = (ctypes.c_int64 * m)(*[int(p[4]) for p in prepared])
        _lib.check(lib.gl_ns_advance_many(handles, m, deaths, streams))
        for ns, eng, n, k, n_deaths, stream in prepared:
            eng.finalize()
        for i, (ns, eng, n, k, n_deaths, stream) in enumerate(prepared):
            st = eng.state()                               # synchronises the engine's stream
            if k <= 8192:
                table_rows = eng.dead_table_device()
            else:
                table_rows, _idx, _step = eng.dead_table()
            out[lo + i] = ns._publish(st, table_rows, int(st.n_evals), eng.live.cpu().numpy(), n)
    return out


def run_seeds(make_sampler, seeds, max_concurrent=16):
    """``make_sampler(seed) -> NestedSampling`` for every seed, run concurrently; returns the samplers.
    The 10-seed evidence statistics north_star asks for in one call."""
    samplers = [make_sampler(int(s)) for s in seeds]
    run_many(samplers, max_concurrent=max_concurrent)
    return samplers


def evidence_statistics(samplers):
    """(mean ln Z, seed-to-seed standard deviation, mean reported error) over finished runs."""
    lz = np.array([s.log_evidence for s in samplers])
    err = np.array([s.log_evidence_error for s in samplers])
    return float(lz.mean()), float(lz.std(ddof=1)) if len(lz) > 1 else 0.0, float(err.mean())
