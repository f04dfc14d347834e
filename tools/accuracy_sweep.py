#!/usr/bin/env python
"""logZ accuracy vs batch size: mean/scatter over seeds against the exact evidence."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

EGG_TRUTH = 235.85594033225414      # trapezoid quadrature of exp((2+cos x cos y)^5) over one period


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--population", type=int, default=10000)
    ap.add_argument("--ks", default="100,1000,2500,5000")
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--tuning-cycles", type=int, default=2)
    ap.add_argument("--dim", type=int, default=2)
    a = ap.parse_args()
    import torch
    from scipy.special import erf
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import death_count
    N, d = a.population, a.dim
    if a.workload == "c2":
        like, table, truth = lk.EggCartonLogLikelihood(2), [(0, 0.0, 10.0 * np.pi)] * 2, EGG_TRUTH
        n_deaths = death_count(N, 1e-6)
        d = 2
    else:
        like, table = lk.IIDGaussianLogLikelihood(d), [(0, -5.0, 10.0)] * d
        truth = d * np.log(erf(5 / np.sqrt(2))) - d * np.log(10.0)
        n_deaths = 10 * N
    for k in [int(v) for v in a.ks.split(",")]:
        zs, errs, ts = [], [], []
        for seed in range(a.seeds):
            eng = DeviceEngine(table, like, N, k, n_deaths, iterations=a.iterations,
                               tuning_cycles=a.tuning_cycles, seed=seed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.init(); eng.advance(n_deaths); eng.finalize()
            st = eng.state()
            ts.append(time.perf_counter() - t0)
            zs.append(st.log_z)
            errs.append(np.sqrt(st.information / N))
        zs = np.array(zs)
        print("N=%d k=%5d deaths=%d: logZ mean %.4f (truth %.4f, bias %+.4f) seed-std %.4f  reported sqrt(H/N) %.4f  "
              "time %.1f ms" % (N, k, n_deaths, zs.mean(), truth, zs.mean() - truth, zs.std(ddof=1), np.mean(errs),
                                1e3 * np.median(ts)), flush=True)


if __name__ == "__main__":
    main()
