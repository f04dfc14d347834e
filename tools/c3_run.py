#!/usr/bin/env python
"""C3: synthetic 50-D correlated Gaussian, converged run at a chosen population."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=100000)
    ap.add_argument("--batch-k", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--tuning-cycles", type=int, default=1)
    ap.add_argument("--cutoff-nats", type=float, default=160.0, help="stop when ln X <= -cutoff")
    ap.add_argument("--dim", type=int, default=50)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per walker (0 = automatic)")
    a = ap.parse_args()
    import torch
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    d, N = a.dim, a.population
    k = a.batch_k or N // 10
    like = lk.DenseGaussianLogLikelihood.synthetic_c3(d)
    n_deaths = int(a.cutoff_nats * N)
    # keep only the scalar outputs: the dead log of a C3-scale run is tens of GB
    eng = DeviceEngine([(0, -5.0, 10.0)] * d, like.device_model(), N, k, min(n_deaths, 2 * k), iterations=a.iterations,
                       tuning_cycles=a.tuning_cycles, seed=1, walker_lanes=a.lanes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.init()
    eng.advance(n_deaths)
    st = eng.state()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("C3 d=%d N=%d k=%d deaths=%d: logZ(no final term) %.4f  truth %.4f  H %.2f  lnX %.1f  L* %.3f  "
          "evals %.3e  wall %.2f s  evals/s %.3e" % (d, N, k, st.n_dead, st.log_z, -d * np.log(10.0), st.information,
                                                    st.log_x, st.lstar, st.n_evals, dt, st.n_evals / dt), flush=True)


if __name__ == "__main__":
    main()
