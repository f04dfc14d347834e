"""Where the end-to-end time of NestedSampling.run() goes (C2 workload): engine creation,
device run, dead-table export + D2H copy, host publication.  python tools/e2e_breakdown.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from scipy.stats import uniform
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import RemainingPriorMass
    from gleipnir_b200.sampled_parameter import SampledParameter, prior_table
    d, N, k = 2, 10000, 1000
    like = EggCartonLogLikelihood(d)
    sps = [SampledParameter(i, uniform(loc=0.0, scale=10.0 * np.pi)) for i in range(d)]

    def sync():
        torch.cuda.synchronize()

    def timed(f):
        sync()
        t = time.perf_counter()
        r = f()
        sync()
        return r, (time.perf_counter() - t) * 1e3

    for rep in range(4):
        ns = NestedSampling(sps, like, N, sampler=MetropolisComponentWiseHardNSRejection(iterations=50, tuning_cycles=2),
                            stopping_criterion=RemainingPriorMass(1e-6), batch_k=k, seed=rep)
        _, t_run = timed(ns.run)
        n_deaths = ns._death_count()
        table = prior_table(sps)
        eng, t_create = timed(lambda: DeviceEngine(table, like, N, k, n_deaths, seed=rep, iterations=50, tuning_cycles=2))
        _, t_dev = timed(lambda: (eng.init(), eng.advance(n_deaths), eng.finalize()))
        st, t_state = timed(eng.state)
        tab, t_export = timed(eng.dead_table_device)
        live, t_live = timed(lambda: eng.live.cpu().numpy())
        _, t_df = timed(lambda: ns.dead_points)
        print("rep %d: run() %.2f ms | create %.2f  device %.2f  state %.3f  export+D2H %.2f  live %.3f  (DataFrame on demand %.2f)"
              % (rep, t_run, t_create, t_dev, t_state, t_export, t_live, t_df))


if __name__ == "__main__":
    main()
