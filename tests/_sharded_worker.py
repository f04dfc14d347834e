"""torchrun worker of tests/test_gpu_multi.py: one rank per GPU, native sharded runs (fused
peer-memory step and NCCL all-gather path) + the public NestedSampling API; results to <outdir>."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from scipy.stats import uniform
    from gleipnir_b200 import _lib as L
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir_b200.sampled_parameter import SampledParameter
    from gleipnir_b200.sharded import NcclComm, run_sharded_native, shard_bounds
    n, k, deaths, d, seed = 4000, 400, 2600, 2, 17           # 6 full batches + a short one
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * d
    like = EggCartonLogLikelihood(d)
    off, n_loc = shard_bounds(n, world, rank)
    res = {}
    for name, no_p2p in (("p2p", "0"), ("nccl", "1")):
        os.environ["GLEIPNIR_B200_NO_P2P"] = no_p2p
        comm = NcclComm()
        eng = DeviceEngine(table, like, n_loc, k, deaths, seed=seed, row_offset=off, global_population=n,
                           iterations=4, tuning_cycles=1)
        reran = run_sharded_native(eng, comm, deaths, k)
        st = eng.state()
        tab, idx, step = eng.dead_table()
        res[name + "_tab"], res[name + "_idx"], res[name + "_step"] = tab, idx, step
        res[name + "_state"] = np.array([st.log_z, st.information, st.log_x, float(st.n_dead), float(st.n_evals),
                                         float(reran), float(comm.lib.gl_comm_p2p_ready(comm.h))])
    os.environ["GLEIPNIR_B200_NO_P2P"] = "0"
    sps = [SampledParameter(i, uniform(loc=0.0, scale=10.0 * np.pi)) for i in range(d)]
    ns = NestedSampling(sps, like, n, sampler=MetropolisComponentWiseHardNSRejection(iterations=4, tuning_cycles=1),
                        stopping_criterion=NumberOfIterations(deaths), batch_k=k, seed=seed)
    logz, err = ns.run()
    res["api_logz"] = np.array([logz, err, float(ns._n_evals)])
    res["api_table"] = ns.dead_points.to_numpy()

    # a user-defined stopping criterion (no closed-form death count): the sharded run is driven batch by batch
    class StopAfter(object):
        def __init__(self, nmax):
            self.nmax = nmax

        def __call__(self, sampler):
            return sampler._n_iterations >= self.nmax
    whole = 2400                                             # a multiple of k: sliced runs stop on batch boundaries
    ns_a = NestedSampling(sps, like, n, sampler=MetropolisComponentWiseHardNSRejection(iterations=4, tuning_cycles=1),
                          stopping_criterion=NumberOfIterations(whole), batch_k=k, seed=seed)
    ns_b = NestedSampling(sps, like, n, sampler=MetropolisComponentWiseHardNSRejection(iterations=4, tuning_cycles=1),
                          stopping_criterion=StopAfter(whole), batch_k=k, seed=seed)
    ra, rb = ns_a.run(), ns_b.run()
    res["sliced_equal"] = np.array([float(ra == rb), float(np.array_equal(ns_a.dead_points.to_numpy(),
                                                                          ns_b.dead_points.to_numpy()))])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
