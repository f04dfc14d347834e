"""C2 replicas: gl_ns_advance_many with direct launches vs CUDA graphs -- identical results, wall time per mode.
python tools/replica_check.py [n_replicas ...]"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                             # noqa: E402
from gleipnir_b200 import _lib as L                      # noqa: E402
from gleipnir_b200.engine import DeviceEngine            # noqa: E402
from gleipnir_b200.likelihoods import EggCartonLogLikelihood   # noqa: E402

lib = L.load()
N, k, n_deaths = 10000, 1000, 138163
table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * 2
like = EggCartonLogLikelihood(2)


def run(n_rep, fn):
    best, states = None, None
    for rep in range(3):
        engs = [DeviceEngine(table, like, N, k, n_deaths, iterations=50, tuning_cycles=2, seed=i, stream=torch.cuda.Stream())
                for i in range(n_rep)]
        torch.cuda.synchronize()
        handles = (ctypes.c_void_p * n_rep)(*[e._h for e in engs])
        streams = (ctypes.c_void_p * n_rep)(*[e._stream() for e in engs])
        deaths = (ctypes.c_int64 * n_rep)(*[n_deaths] * n_rep)
        t0 = time.perf_counter()
        for e in engs:
            e.init()
        L.check(fn(handles, n_rep, deaths, streams))
        for e in engs:
            e.finalize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = [e.state() for e in engs]
        if best is None or dt < best:
            best, states = dt, st
    return best, np.array([s.log_z for s in states]), sum(s.n_evals for s in states)


for n_rep in [int(a) for a in sys.argv[1:]] or [1, 10, 16, 32]:
    t_a, z_a, ev = run(n_rep, lib.gl_ns_advance_many)
    t_b, z_b, _ = run(n_rep, lib.gl_ns_advance_many_graphs)
    print("replicas %2d: direct %.2f ms (%.2f per run, %.2f G evals/s) | graphs %.2f ms (%.2f per run, %.2f G evals/s) | "
          "identical ln Z: %s" % (n_rep, 1e3 * t_a, 1e3 * t_a / n_rep, ev / t_a / 1e9, 1e3 * t_b, 1e3 * t_b / n_rep,
                                  ev / t_b / 1e9, bool(np.array_equal(z_a, z_b))))
