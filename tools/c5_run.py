#!/usr/bin/env python
"""C5-shaped measurement: the SYNTHETIC JARM-like network (16 species, 18 sampled log10 rate
constants, 181 output times; the JARM model files are absent upstream) through the cooperative
ODE walker.  Bounded slice of a run: a few batches, reporting solves/s next to the SciPy stand-in."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=8192)
    ap.add_argument("--batch-k", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--tuning-cycles", type=int, default=1)
    a = ap.parse_args()
    import torch
    from test_gpu_ode import _jarm_setup
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    from oracle import ode
    net, ts, data, sampled, rows, desc = _jarm_setup()
    like = OdeNetworkLogLikelihood(net, sampled, ts, data, rtol=1e-6)
    nominal = np.log10(np.array([net.param_values[net.param_names.index(n)] for n in sampled]))
    table = [(0, float(v) - 2.0, 4.0) for v in nominal]        # NestedSampleIt default priors (:265)
    N, k = a.population, a.batch_k
    n_deaths = k * a.steps
    eng = DeviceEngine(table, like, N, k, n_deaths, iterations=a.iterations,
                       tuning_cycles=a.tuning_cycles, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.init()
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    eng.advance(n_deaths)
    st = eng.state()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("C5-like d=18 NS=16 N=%d k=%d batches=%d: init %d solves in %.2f s (%.0f solves/s); "
          "total %.3e solves in %.1f s -> %.0f solves/s; logZ so far %.2f" %
          (N, k, a.steps, N, t_init, N / t_init, st.n_evals, dt, st.n_evals / dt, st.log_z), flush=True)
    sidx = [net.param_names.index(n) for n in sampled]
    Y = eng.live[:20].cpu().numpy()
    t0 = time.perf_counter()
    for y in Y:
        ode.loglike(desc, sidx, y, ts, rows)
    print("SciPy LSODA stand-in (1 core): %.2f solves/s" % (20 / (time.perf_counter() - t0)))


if __name__ == "__main__":
    main()
