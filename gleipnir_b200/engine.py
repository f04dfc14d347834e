"""``DeviceEngine`` -- thin Python handle on the C-ABI Nested Sampling engine
(``gl_ns_*`` in ``include/gleipnir_b200.h``).  torch provides device memory and the
stream; every kernel is ours.
"""
import ctypes

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.GleipnirB200Error("no CUDA device is available and gleipnir_b200 has no CPU fallback")
    return torch


def _on_device(fn):
    """Run a DeviceEngine method with the engine's device current: the C library launches on the current
    device (kernels, cudaMalloc of model data, peer buffers), so `device='cuda:1'` must make it current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with _torch().cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


class DeviceEngine(object):
    """One device-resident NS run: live set, dead-point log, accumulators, widths.

    ``model_handle`` is a DeviceLogLikelihood (preferred: the engine then keeps it alive) or a raw
    ``gl_model`` handle, which must outlive the engine."""

    def __init__(self, prior_table, model_handle, population_size, batch_k, n_deaths_capacity, iterations=100,
                 burn_in=0, tuning_cycles=0, proposal=0, seed=0, device=None, walker_lanes=0, row_offset=0,
                 global_population=None, stream=None):
        torch = _torch()
        self.lib = _lib.load()
        if hasattr(model_handle, "device_model"):       # a DeviceLogLikelihood: keep it alive with the engine
            self._like = model_handle
            model_handle = model_handle.device_model()
        self.d = len(prior_table)
        self.N = int(population_size)
        self.stream = stream                             # torch.cuda.Stream of this engine (None: the current stream)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self._construct(torch, prior_table, model_handle, batch_k, n_deaths_capacity, iterations, burn_in,
                            tuning_cycles, proposal, seed, walker_lanes, row_offset, global_population)

    def _construct(self, torch, prior_table, model_handle, batch_k, n_deaths_capacity, iterations, burn_in,
                   tuning_cycles, proposal, seed, walker_lanes, row_offset, global_population):
        cap = int(n_deaths_capacity) + self.N
        self.cfg = _lib.gl_ns_config(population_size=self.N, batch_k=int(batch_k), iterations=int(iterations),
                                     burn_in=int(burn_in), tuning_cycles=int(tuning_cycles),
                                     proposal=int(proposal), walker_lanes=int(walker_lanes), seed=int(seed) & (2 ** 64 - 1),
                                     dead_capacity=cap, row_offset=int(row_offset),
                                     global_population=int(global_population or self.N))
        ws = self.lib.gl_ns_workspace_bytes(ctypes.byref(self.cfg), self.d)
        if ws < 0:
            _lib.check(int(ws))
        f64 = dict(dtype=torch.float64, device=self.device)
        self.live = torch.empty((self.N, self.d), **f64)
        self.logl = torch.empty(self.N, **f64)
        self.dead_x = torch.empty((cap, self.d), **f64)
        self.dead_logl = torch.empty(cap, **f64)
        self.dead_w = torch.empty(cap, **f64)
        self.dead_idx = torch.empty(cap, dtype=torch.int64, device=self.device)
        self.dead_step = torch.empty(cap, dtype=torch.int32, device=self.device)
        self.workspace = torch.empty(int(ws), dtype=torch.uint8, device=self.device)
        self._bufs = _lib.gl_ns_buffers(self.live.data_ptr(), self.logl.data_ptr(), self.dead_x.data_ptr(),
                                        self.dead_logl.data_ptr(), self.dead_w.data_ptr(),
                                        self.dead_idx.data_ptr(), self.dead_step.data_ptr(),
                                        self.workspace.data_ptr())
        self._priors = _lib.prior_array(prior_table)
        self._model = model_handle
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.gl_ns_create(ctypes.byref(self.cfg), self._priors, self.d, model_handle,
                                         ctypes.byref(self._bufs), ctypes.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.gl_ns_destroy(h)
            self._h = ctypes.c_void_p()

    def _stream(self):
        if self.stream is not None:
            return self.stream.cuda_stream
        return _torch().cuda.current_stream(self.device).cuda_stream

    def _torch_stream(self):
        return self.stream if self.stream is not None else _torch().cuda.current_stream(self.device)

    def adopt_population(self, live, logl):
        """Start the run from a caller-supplied population (resuming a run, or a population conditioned by the
        caller) instead of prior draws: every later ``init()`` -- also the ones the sharded drivers issue -- copies
        these rows.  ``live`` [N, d] and ``logl`` [N]: numpy arrays or tensors (any device)."""
        self._adopted = (live, logl)

    @_on_device
    def init(self, live=None, logl=None):
        """Draw + evaluate the initial population (or adopt a caller-supplied one)."""
        torch = _torch()
        skip = 0
        if live is None and getattr(self, "_adopted", None) is not None:
            live, logl = self._adopted
        if live is not None:
            def dev(a):
                if not isinstance(a, torch.Tensor):
                    a = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
                return a.to(device=self.device, dtype=torch.float64)
            self.live.copy_(dev(live).reshape(self.N, self.d))
            self.logl.copy_(dev(logl).reshape(self.N))
            skip = 1
        _lib.check(self.lib.gl_ns_init(self._h, skip, self._stream()))

    @_on_device
    def advance(self, n_deaths_total):
        _lib.check(self.lib.gl_ns_advance(self._h, int(n_deaths_total), self._stream()))

    @_on_device
    def kill(self, k):
        _lib.check(self.lib.gl_ns_kill(self._h, int(k), self._stream()))

    @_on_device
    def refill(self):
        _lib.check(self.lib.gl_ns_refill(self._h, self._stream()))

    @_on_device
    def inject(self, points, logls):
        torch = _torch()
        p = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64)).to(self.device)
        l = torch.as_tensor(np.ascontiguousarray(logls, dtype=np.float64)).to(self.device)
        _lib.check(self.lib.gl_ns_inject(self._h, p.data_ptr(), l.data_ptr(), int(l.numel()), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()     # p, l must outlive the kernels

    @_on_device
    def finalize(self):
        _lib.check(self.lib.gl_ns_finalize(self._h, self._stream()))

    def force_smem_walker(self, enable=True):
        """Use the generic shared-memory walker even where the register-resident one applies."""
        _lib.check(self.lib.gl_ns_force_smem_walker(self._h, 1 if enable else 0))

    def force_multipass_select(self, enable=True):
        _lib.check(self.lib.gl_ns_force_multipass_select(self._h, 1 if enable else 0))

    # ---- sharded run (one engine per rank; see gleipnir_b200/sharded.py) ---------------------
    @_on_device
    def candidates(self, k):
        torch = _torch()
        out = torch.empty(int(k), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gl_ns_shard_candidates(self._h, int(k), out.data_ptr(), self._stream()))
        return out

    @_on_device
    def commit(self, gathered, world, rank, k):
        """`gathered` holds m = len/world candidates per rank (m <= k: truncated exchange)."""
        self._gathered = gathered                       # must outlive the enqueued kernels
        n = int(gathered.numel())
        m = n // int(world)
        scratch = None
        if n > 24576:                                   # keys of the fused commit go to a global scratch
            torch = _torch()
            if getattr(self, "_key_scratch", None) is None or self._key_scratch.numel() < n:
                self._key_scratch = torch.empty(n, dtype=torch.int64, device=self.device)
            scratch = self._key_scratch.data_ptr()
        _lib.check(self.lib.gl_ns_shard_commit_truncated(self._h, gathered.data_ptr(), int(world), int(rank), int(k),
                                                         m, scratch, self._stream()))

    def overflowed(self):
        """True when a truncated candidate exchange could not prove one of its steps exact."""
        flag = int(self.state().shard_overflow)
        if flag == 2:
            raise RuntimeError("sharded run: a peer did not deliver its candidates within 10 s (peer-memory exchange)")
        return bool(flag)

    @_on_device
    def final_partial(self):
        torch = _torch()
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gl_ns_shard_final_partial(self._h, out.data_ptr(), self._stream()))
        return out

    @_on_device
    def finalize_sharded(self, gathered, world, rank):
        self._gathered = gathered
        _lib.check(self.lib.gl_ns_shard_finalize(self._h, gathered.data_ptr(), int(world), int(rank), self._stream()))

    def set_select_path(self, path):
        """0 auto, 1 multi-launch radix, 2 fused single-CTA radix, 3 fused sorted-merge, 4 cluster radix."""
        _lib.check(self.lib.gl_ns_set_select_path(self._h, int(path)))

    def profile_enable(self, enable=True):
        _lib.check(self.lib.gl_ns_profile_enable(self._h, 1 if enable else 0))

    @_on_device
    def profile_read(self):
        """(walk_ms, walk_launches, select_ms, select_launches) from CUDA events recorded on the
        launching stream around the walker / selection kernels."""
        wm, sm_ = ctypes.c_double(0), ctypes.c_double(0)
        wl, sl = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self.lib.gl_ns_profile_read(self._h, ctypes.byref(wm), ctypes.byref(wl), ctypes.byref(sm_),
                                               ctypes.byref(sl)))
        return wm.value, wl.value, sm_.value, sl.value

    @_on_device
    def state(self):
        st = _lib.gl_ns_state()
        _lib.check(self.lib.gl_ns_read_state(self._h, ctypes.byref(st), self._stream()))
        return st

    @_on_device
    def dead_table_device(self):
        """`dead_points` values assembled and ordered on the device (gl_ns_export_dead):
        host array [rows, d+2] in death order after ONE device->host copy.  Batches larger than
        8192 rows are left in slot order by the kernel; use dead_table() for those."""
        torch = _torch()
        st = self.state()
        rows = int(st.n_rows)
        ts = self._torch_stream()
        with torch.cuda.stream(ts):
            table = torch.empty((rows, self.d + 2), dtype=torch.float64, device=self.device)
            _lib.check(self.lib.gl_ns_export_dead(self._h, table.data_ptr(), self._stream()))
            # pinned staging (torch's caching host allocator recycles the block between runs): the copy
            # runs at PCIe rate instead of through a pageable bounce buffer
            host = torch.empty((rows, self.d + 2), dtype=torch.float64, pin_memory=True)
            host.copy_(table, non_blocking=True)
            ts.synchronize()
        return host.numpy()

    @_on_device
    def dead_table(self):
        """Dead-point log as a host array [rows, d+2] (log_l, weight, x...) in death order,
        plus the slot index and batch id of every row."""
        torch = _torch()
        st = self.state()
        rows = int(st.n_rows)
        logl = self.dead_logl[:rows].cpu().numpy()
        w = self.dead_w[:rows].cpu().numpy()
        x = self.dead_x[:rows].cpu().numpy()
        idx = self.dead_idx[:rows].cpu().numpy()
        step = self.dead_step[:rows].cpu().numpy()
        order = death_order(logl, idx, step)
        table = np.concatenate((logl[order, None], w[order, None], x[order]), axis=1)
        return table, idx[order], step[order]


def sort_key_host(logl):
    """Host copy of the device ordering key (NaN first, -0.0 == +0.0)."""
    x = np.asarray(logl, dtype=np.float64) + 0.0
    b = x.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    k = np.where(neg, ~b, b | np.uint64(0x8000000000000000))
    return np.where(np.isnan(x), np.uint64(0), k)


def death_order(logl, idx, step):
    """Permutation putting the device's dead rows (index order inside a batch) into death
    order: batches in sequence, ascending (logL, slot) inside a batch, survivors (step -1)
    last in slot order."""
    step = np.asarray(step, dtype=np.int64)
    batch = np.where(step < 0, np.iinfo(np.int64).max, step)
    key = np.where(step < 0, np.uint64(0), sort_key_host(logl))
    return np.lexsort((idx, key, batch))
