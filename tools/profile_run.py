#!/usr/bin/env python
"""Short, fixed-shape run of the hot path for ncu (launch list / --set full captures).

    python tools/profile_run.py --workload c2 --steps 4
runs `steps` batches of the named workload (init + steps x (kill, refill) + finalize)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--population", type=int, default=0)
    ap.add_argument("--batch-k", type=int, default=0)
    a = ap.parse_args()
    import torch
    from gleipnir_b200 import _lib as L
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    if a.workload == "c1":
        d, N, like, table, skw = 5, 500, lk.IIDGaussianLogLikelihood(5), [(0, -5.0, 10.0)] * 5, dict(iterations=10, tuning_cycles=1)
    elif a.workload == "c2":
        d, N, like = 2, 10000, lk.EggCartonLogLikelihood(2)
        table, skw = [(0, 0.0, 10.0 * np.pi)] * 2, dict(iterations=50, tuning_cycles=2)
    else:
        d, N, like = 50, 1000000, lk.DenseGaussianLogLikelihood.synthetic_c3(50)
        table, skw = [(0, -5.0, 10.0)] * 50, dict(iterations=10, tuning_cycles=1)
    N = a.population or N
    k = a.batch_k or N // 10
    n_deaths = k * a.steps
    eng = DeviceEngine(table, like, N, k, n_deaths, seed=1, **skw)
    eng.init()
    eng.advance(n_deaths)
    eng.finalize()
    st = eng.state()
    torch.cuda.synchronize()
    print("workload=%s N=%d k=%d deaths=%d evals=%d launches=%d logZ=%.6f" %
          (a.workload, N, k, st.n_dead, st.n_evals, st.n_launches, st.log_z))


if __name__ == "__main__":
    main()
