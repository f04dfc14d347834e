"""Sharded Nested Sampling run: the live set and the walker batch are split over the ranks
(GPUs) of one box; the only data-path collective is one all-gather of k log-likelihoods per
rank per step (the global likelihood threshold L*), plus one all-gather of 4 doubles per rank
for the final live-point term (north_star (e)).

``run_sharded`` is the SPMD loop.  It is written against two small protocols so that the same
code drives one CUDA engine per process over NCCL (production), several CUDA engines inside one
process (single-GPU emulation for tests), or -- in the CPU test-suite only -- numpy oracle
shards over gloo:

  shard ops    init() / candidates(k) / commit(gathered, world, rank, k) / refill() /
               final_partial() / finalize_sharded(gathered, world, rank)
               (``DeviceEngine`` implements them on the GPU)
  comm         world, ranks (global rank of each local shard), all_gather(list of local arrays)
               -> rank-major concatenation over the whole world
"""


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


class LocalComm(object):
    """All shards live in this process (emulation of G ranks on one device)."""

    def __init__(self, world):
        self.world = int(world)
        self.ranks = list(range(self.world))

    def all_gather(self, parts):
        if _is_torch(parts[0]):
            import torch
            return torch.cat(list(parts))
        import numpy as np
        return np.concatenate(list(parts))


class TorchDistComm(object):
    """One shard per process; torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.ranks = [dist.get_rank(group)]

    def all_gather(self, parts):
        import torch
        (mine,) = parts
        is_np = not _is_torch(mine)
        t = torch.from_numpy(mine) if is_np else mine
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out.numpy() if is_np else out


def shard_bounds(n_global, world, rank):
    """Contiguous block of global slots owned by `rank`."""
    if n_global % world:
        raise ValueError("population_size (%d) must be divisible by the number of ranks (%d)" % (n_global, world))
    n_local = n_global // world
    return rank * n_local, n_local


def run_sharded(shards, comm, n_deaths, batch_k, init=True):
    """Run the local `shards` (engines of the ranks ``comm.ranks``) to ``n_deaths`` global deaths."""
    if init:
        for sh in shards:
            sh.init()
    done = 0
    while done < n_deaths:
        k = min(int(batch_k), n_deaths - done)
        gathered = comm.all_gather([sh.candidates(k) for sh in shards])
        for sh, r in zip(shards, comm.ranks):
            sh.commit(gathered, comm.world, r, k)
        done += k
        if done < n_deaths:
            for sh in shards:
                sh.refill()
    partials = comm.all_gather([sh.final_partial() for sh in shards])
    for sh, r in zip(shards, comm.ranks):
        sh.finalize_sharded(partials, comm.world, r)
    return done
