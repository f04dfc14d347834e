"""Converged Nested Sampling runs on the JARM problem (C5) at a small population, several seeds concurrently
(run_many): ln Z per seed, seed scatter vs reported error, time to converged ln Z.
python tools/c5_converged.py [N] [k] [ln_X_stop] [n_seeds]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                             # noqa: E402
from gleipnir_b200 import jarm                            # noqa: E402
from gleipnir_b200.nestedsampling import NestedSampling   # noqa: E402
from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection   # noqa: E402
from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations   # noqa: E402
from gleipnir_b200.replicas import run_many                # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
k = int(sys.argv[2]) if len(sys.argv) > 2 else 125
lnx = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
n_deaths = int(lnx * N)
like = jarm.jarm_loglikelihood()
runs = [NestedSampling(jarm.jarm_sampled_parameters(), like, N,
                       sampler=MetropolisComponentWiseHardNSRejection(iterations=10, burn_in=10, tuning_cycles=1),
                       stopping_criterion=NumberOfIterations(n_deaths), batch_k=k, seed=100 + s) for s in range(seeds)]
like.counters(reset=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
run_many(runs)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
c = like.counters()
lz = np.array([r.log_evidence for r in runs])
err = np.array([r.log_evidence_error for r in runs])
out = dict(config="JARM (C5 likelihood), N=%d, batch_k=%d, sampler(10, burn 10, tc 1), NumberOfIterations(%d) = ln X_stop %.0f; "
                  "%d seeds advanced concurrently on one B200" % (N, k, n_deaths, -lnx, seeds),
           logZ=lz.tolist(), logZ_err_reported=err.tolist(), information=[r.information for r in runs],
           logZ_mean=float(lz.mean()), logZ_seed_scatter=float(lz.std(ddof=1)) if seeds > 1 else None,
           max_logL=[float(r.max_loglikelihood()) for r in runs],
           final_live_term_fraction=[float(np.exp(np.log(r.dead_points["weight"].to_numpy()[n_deaths:].sum()) +
                                                  np.log(np.mean(np.exp(r.dead_points["log_l"].to_numpy()[n_deaths:] -
                                                                        r.max_loglikelihood()))) + r.max_loglikelihood() -
                                                  r.log_evidence)) for r in runs],
           wall_s_all_seeds=wall, evals=c["evals"], evals_per_s=c["evals"] / wall, ode_steps_per_solve=c["steps"] / c["solves"])
print(json.dumps(out))
