#!/usr/bin/env python
"""Tiny end-to-end exercise of every kernel family, for compute-sanitizer."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from scipy.stats import norm, uniform
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200 import networks
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    from gleipnir_b200.sampled_parameter import SampledParameter
    from gleipnir_b200.sharded import LocalComm, run_sharded, shard_bounds
    # fused + multipass engine, posteriors, export
    for mp in (False, True):
        sps = [SampledParameter(i, uniform(0, 10 * np.pi)) for i in range(2)]
        ns = NestedSampling(sps, lk.EggCartonLogLikelihood(2), 300, sampler=MetropolisComponentWiseHardNSRejection(5, tuning_cycles=1),
                            stopping_criterion=NumberOfIterations(600), seed=1)
        ns.run()
        ns.posteriors()
        e = DeviceEngine([(0, -5.0, 10.0)] * 3, lk.IIDGaussianLogLikelihood(3), 200, 20, 200, iterations=3, tuning_cycles=1, seed=2)
        if mp:
            e.force_multipass_select(True)
        e.init(); e.advance(200); e.finalize(); e.dead_table(); e.dead_table_device()
    # dense (DMMA + incremental walker), mixture
    d = 12
    dg = lk.DenseGaussianLogLikelihood.synthetic_c3(d)
    e = DeviceEngine([(0, -5.0, 10.0)] * d, dg, 128, 16, 64, iterations=2, tuning_cycles=1, seed=3)
    e.init(); e.advance(64); e.finalize()
    e = DeviceEngine([(2, 1e-2, 2.0)], lk.Mixture1DLogLikelihood(), 64, 8, 32, iterations=2, tuning_cycles=1, proposal=1, seed=4)
    e.init(); e.advance(32); e.finalize()
    # sharded (emulated)
    shards = []
    for r in range(2):
        off, nl = shard_bounds(200, 2, r)
        shards.append(DeviceEngine([(0, -5.0, 10.0)] * 2, lk.IIDGaussianLogLikelihood(2), nl, 20, 100, seed=5, row_offset=off,
                                   global_population=200, iterations=2, tuning_cycles=1))
    run_sharded(shards, LocalComm(2), 100, 20)
    # ODE: thread-per-system and cooperative
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_ode.npz"))
    ol = OdeNetworkLogLikelihood(networks.dimerization(), ["kf", "kr"], g["dimerization/tspan"],
                                 {"A_dimer": (g["dimerization/data"], 10.0 * g["dimerization/sd"], None)})
    e = DeviceEngine([(1, -3.0, 1.0), (1, 0.0, 1.0)], ol, 32, 4, 8, iterations=1, tuning_cycles=0, seed=6)
    e.init(); e.advance(8); e.finalize()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_gpu_ode import _jarm_setup
    net, ts, data, sampled, rows, desc = _jarm_setup()
    jl = OdeNetworkLogLikelihood(net, sampled, ts, data, rtol=1e-5)
    nominal = np.log10(np.array([net.param_values[net.param_names.index(n)] for n in sampled]))
    e = DeviceEngine([(0, float(v) - 0.5, 1.0) for v in nominal], jl, 16, 4, 8, iterations=1, tuning_cycles=0, seed=7)
    e.init(); e.advance(8); e.finalize()
    # reference-exact + per-call sampler
    from gleipnir_b200.reference_exact import run_reference_exact
    run_reference_exact(0, [-5.0] * 2, [10.0] * 2, -1.8, 10, iterations=2, tuning_cycles=1, n_iterations=12)
    s = MetropolisComponentWiseHardNSRejection(iterations=3, tuning_cycles=1)
    s([SampledParameter('a', norm(0, 1))], lk.IIDGaussianLogLikelihood(1), np.array([0.1]), -5.0, seed=1)
    torch.cuda.synchronize()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
