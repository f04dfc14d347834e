"""Sharded Nested Sampling run: the live set and the walker batch are split over the ranks
(GPUs) of one box; the only data-path collective is one all-gather of k log-likelihoods per
rank per step (the global likelihood threshold L*), plus one all-gather of 4 doubles per rank
for the final live-point term (north_star (e)).

``run_sharded`` is the SPMD loop.  It is written against two small protocols so that the same
code drives one CUDA engine per process over NCCL (production), several CUDA engines inside one
process (single-GPU emulation for tests), or -- in the CPU test-suite only -- numpy oracle
shards over gloo:

  shard ops    init() / candidates(m) / commit(gathered, world, rank, k) / refill() /
               final_partial() / finalize_sharded(gathered, world, rank) / overflowed()
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


DEFAULT_CANDIDATE_SIGMAS = 8.0


def candidate_count(k, world, n_local, sigmas):
    """Candidates each rank contributes to a step of k global deaths (same arithmetic as
    ``gl_ns_shard_candidate_count``): the k deaths fall on a rank as Binomial(k, 1/world), so
    mean + `sigmas` standard deviations (+8) covers it; the commit proves each step exact and
    raises the overflow flag when it cannot.  ``sigmas <= 0``: the full k per rank."""
    import math
    m = int(k)
    if world > 1 and sigmas > 0.0:
        p = 1.0 / float(world)
        m = int(math.ceil(float(k) * p + sigmas * math.sqrt(float(k) * p * (1.0 - p)) + 8.0))
    m = min(m, int(k), int(n_local) - 1)
    if m * world < k:
        m = (int(k) + world - 1) // world
    return m


def _run_sharded_once(shards, comm, n_deaths, batch_k, sigmas):
    for sh in shards:
        sh.init()
    done = 0
    while done < n_deaths:
        k = min(int(batch_k), n_deaths - done)
        m = candidate_count(k, comm.world, shards[0].N, sigmas)
        gathered = comm.all_gather([sh.candidates(m) for sh in shards])
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


def run_sharded(shards, comm, n_deaths, batch_k, candidate_sigmas=DEFAULT_CANDIDATE_SIGMAS):
    """Run the local `shards` (engines of the ranks ``comm.ranks``) to ``n_deaths`` global deaths.
    A truncated candidate exchange that could not prove one of its steps exact (the flag is
    identical on every rank) is repeated from the start with the full exchange."""
    done = _run_sharded_once(shards, comm, n_deaths, batch_k, candidate_sigmas)
    if candidate_sigmas > 0.0 and any(sh.overflowed() for sh in shards):
        done = _run_sharded_once(shards, comm, n_deaths, batch_k, 0.0)
    return done


class NcclComm(object):
    """Native communicator (``gl_comm``): NCCL called from the library's own C++ loop, on the
    engine's stream.  The NCCL unique id is created on rank 0 and broadcast with
    torch.distributed (plumbing); ``gl_comm_create`` is collective."""

    def __init__(self, group=None):
        import ctypes

        import torch.distributed as dist

        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        nbytes = self.lib.gl_comm_unique_id_bytes()
        buf = (ctypes.c_uint8 * nbytes)()
        if self.rank == 0:
            _lib.check(self.lib.gl_comm_unique_id(buf))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (ctypes.c_uint8 * nbytes).from_buffer_copy(box[0])
        self.h = ctypes.c_void_p()
        _lib.check(self.lib.gl_comm_create(self.world, self.rank, ident, ctypes.byref(self.h)))

    def enable_p2p(self, cap):
        """Map every rank's exchange buffer into this process (CUDA IPC over NVLink) so that
        gl_ns_advance_sharded can run its fused one-kernel step; collective.  Returns True when
        the peer path is ready (False: the NCCL all-gather path stays in use)."""
        import ctypes
        import os

        import torch
        import torch.distributed as dist

        from . import _lib
        if self.lib.gl_comm_p2p_ready(self.h):
            return True
        if getattr(self, "_p2p_failed", False) or os.environ.get("GLEIPNIR_B200_NO_P2P") == "1":
            return False
        nbytes = self.lib.gl_comm_p2p_handle_bytes()
        buf = (ctypes.c_uint8 * nbytes)()
        rc = self.lib.gl_comm_p2p_alloc(self.h, max(int(cap), 16384), buf)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(buf) if rc == 0 else None, group=self.group)
        if any(h is None for h in handles):               # some rank could not allocate: nobody uses it
            self._p2p_failed = True
            return False
        allb = b"".join(handles)
        arr = (ctypes.c_uint8 * len(allb)).from_buffer_copy(allb)
        rc = self.lib.gl_comm_p2p_open(self.h, arr)
        oks = [None] * self.world
        dist.all_gather_object(oks, rc == 0, group=self.group)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if not all(oks):                                  # every rank takes the same path: NCCL all-gather
            import warnings
            self._p2p_failed = True
            self.lib.gl_comm_p2p_disable(self.h)
            if rc != 0:
                warnings.warn("peer-memory exchange unavailable (%s); using the NCCL all-gather path"
                              % (self.lib.gl_last_error() or b"?").decode())
            return False
        return True

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value:
            self.lib.gl_comm_destroy(h)
            self.h = None


_comm_cache = {}


def get_nccl_comm(group=None):
    """One native communicator per process group (creating one costs a NCCL bootstrap)."""
    key = id(group)
    if key not in _comm_cache:
        _comm_cache[key] = NcclComm(group)
    return _comm_cache[key]


def gather_dead_tables(engine, group=None):
    """All ranks' dead-point logs on every rank: (rows [n, d+2], slot idx, step), unsorted.
    One padded tensor all-gather over NCCL instead of pickled objects."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = int(engine.state().n_rows)
    d = engine.d
    counts = torch.zeros(world, dtype=torch.int64, device=engine.device)
    dist.all_gather_into_tensor(counts, torch.tensor([rows], dtype=torch.int64, device=engine.device), group=group)
    counts = counts.cpu().numpy()
    mx = int(counts.max())
    mine = torch.zeros((mx, d + 4), dtype=torch.float64, device=engine.device)
    mine[:rows, 0] = engine.dead_logl[:rows]
    mine[:rows, 1] = engine.dead_w[:rows]
    mine[:rows, 2:2 + d] = engine.dead_x[:rows]
    mine[:rows, 2 + d] = engine.dead_idx[:rows].to(torch.float64)        # exact below 2^53
    mine[:rows, 3 + d] = engine.dead_step[:rows].to(torch.float64)
    allr = torch.empty((world * mx, d + 4), dtype=torch.float64, device=engine.device)
    dist.all_gather_into_tensor(allr, mine, group=group)
    allr = allr.cpu().numpy().reshape(world, mx, d + 4)
    parts = [allr[r, :counts[r]] for r in range(world)]
    cat = np.concatenate(parts)
    return cat[:, :d + 2], cat[:, d + 2].astype(np.int64), cat[:, d + 3].astype(np.int64)


def gather_dead_table_device(engine, batch_k, n_deaths, group=None):
    """All ranks' dead-point logs assembled on the device into the reference's table (death order):
    one padded all-gather over NCCL, one kernel (gl_export_gathered_table), one pinned D2H copy.
    Returns the host array [rows, d+2]; every rank gets the same table.  batch_k <= 8192."""
    import torch
    import torch.distributed as dist

    from . import _lib
    world = dist.get_world_size(group)
    rows = int(engine.state().n_rows)
    d = engine.d
    dev = engine.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([rows], dtype=torch.int64, device=dev), group=group)
    counts_h = counts.cpu()
    mx, total = int(counts_h.max()), int(counts_h.sum())
    mine = torch.empty((mx, d + 4), dtype=torch.float64, device=dev)
    mine[:rows, 0] = engine.dead_logl[:rows]
    mine[:rows, 1] = engine.dead_w[:rows]
    mine[:rows, 2:2 + d] = engine.dead_x[:rows]
    mine[:rows, 2 + d] = engine.dead_idx[:rows].to(torch.float64)        # exact below 2^53
    mine[:rows, 3 + d] = engine.dead_step[:rows].to(torch.float64)
    allr = torch.empty((world * mx, d + 4), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allr, mine, group=group)
    n_steps = (int(n_deaths) + int(batch_k) - 1) // int(batch_k)
    table = torch.empty((total, d + 2), dtype=torch.float64, device=dev)
    _lib.check(engine.lib.gl_export_gathered_table(allr.data_ptr(), counts.data_ptr(), world, mx, d, int(batch_k), n_steps,
                                                   int(n_deaths), table.data_ptr(), engine._stream()))
    host = torch.empty((total, d + 2), dtype=torch.float64, pin_memory=True)
    host.copy_(table, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host.numpy()


def run_sharded_native(engine, nccl_comm, n_deaths, batch_k, candidate_sigmas=DEFAULT_CANDIDATE_SIGMAS,
                       check_overflow=True):
    """Whole sharded run enqueued from C++ (gl_ns_advance_sharded / gl_ns_finalize_sharded).
    ``check_overflow`` reads the state back (one sync) and repeats the run with the full exchange
    when the truncated one could not prove a step exact; returns True when that happened."""
    from . import _lib
    st = engine._stream()
    nccl_comm.enable_p2p(batch_k)             # fused one-kernel step over NVLink peer memory when available

    def once(sigmas):
        engine.init()
        _lib.check(engine.lib.gl_ns_advance_sharded(engine._h, nccl_comm.h, int(n_deaths), int(batch_k), 0,
                                                    float(sigmas), st))
        _lib.check(engine.lib.gl_ns_finalize_sharded(engine._h, nccl_comm.h, st))

    once(candidate_sigmas)
    if not check_overflow:
        return False
    # the verdict of a truncated exchange is identical on every rank, a watchdog timeout is rank-local: take the
    # maximum over the ranks so that all of them repeat the run, or all of them raise
    flag = int(engine.state().shard_overflow)
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([flag], dtype=torch.int32, device=engine.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            flag = int(t.item())
    except ImportError:          # pragma: no cover
        pass
    if flag == 2:
        raise RuntimeError("sharded run: a peer did not deliver its candidates within GLEIPNIR_PEER_TIMEOUT_S "
                           "(peer-memory exchange)")
    if flag and candidate_sigmas > 0.0:
        once(0.0)
        return True
    return False


def run_sharded_sliced(engine, nccl_comm, batch_k, max_deaths, should_stop):
    """Sharded run driven batch by batch from the host (gl_ns_step_sharded): after every batch ``should_stop(state)``
    is evaluated -- the run state is redundantly identical on every rank, so every rank takes the same decision with
    no extra collective.  Full candidate exchange (a truncated one could only be validated after the run)."""
    from . import _lib
    st = engine._stream()
    nccl_comm.enable_p2p(batch_k)
    engine.init()
    done = 0
    while done + batch_k <= max_deaths:
        _lib.check(engine.lib.gl_ns_step_sharded(engine._h, nccl_comm.h, 1, int(batch_k), int(done), 0.0, st))
        done += batch_k
        if should_stop(engine.state()):
            break
    _lib.check(engine.lib.gl_ns_finalize_sharded(engine._h, nccl_comm.h, st))
    if int(engine.state().shard_overflow) == 2:
        raise RuntimeError("sharded run: a peer did not deliver its candidates within GLEIPNIR_PEER_TIMEOUT_S")
    return done
