#!/usr/bin/env python
"""bench.py -- logL evals/s of the B200 Nested Sampling path on BASELINE.json's headline config.

Headline workload (BASELINE.json configs[4], SURVEY 8d "C5"): the JARM likelihood
(examples/pysb_jarm/run_jarm_multinest.py:72-143) on the synthetic JARM-shaped network
(34 species, 84 reactions, 50 parameters, 18 sampled; gleipnir_b200/jarm.py), N = 10^5 live points,
sampler MetropolisComponentWiseHardNSRejection(iterations=10, burn_in=10, tuning_cycles=1) -- what
NestedSampleIt configures for the built-in sampler (nestedsample_it.py:416-427).

A "step" is ONE Nested Sampling batch of the running C5 run: the k worst live points die (selection,
evidence / information update, dead-row record) and k walkers refill them, each doing 20*d tuning
+ 20*d production trial moves = 721 likelihood evaluations of 4 stiff ODE solves.  k = one resident wave of the
walker kernel per GPU (148 SMs x 4 blocks x 2 walkers = 1184; sharded runs keep 3.5 sigma of headroom).  A converged C5 run is ~4*10^3 such steps (hours), so it is
not a bench step; the converged-run figures carried on the same line are C2 (eggcarton, N = 10^4) and
C3 (50-D dense Gaussian, N = 10^6) as `secondary`.

`value` = likelihood evaluations actually integrated on the device (counted by the kernel; trial
moves the prior rejects are not integrated and not counted) / device time, inputs resident.
`e2e` = the same through the public `NestedSampling(...).run()` call on a bounded run (host priors
and config in, dead-point table out, inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C5 = dict(name="C5 JARM ODE likelihood (synthetic JARM-shaped network: 34 species, 84 reactions, 18 sampled of 50 "
               "parameters; 2 conditions x (pre-equilibration + 181-time solve), 5 data terms + 4 cycle terms), "
               "N=1e5 live points, one NS batch per step",
          population=100000, iterations=10, burn_in=10, tuning_cycles=1, d=18)


def c5_walkers_per_gpu(like, world):
    """One resident wave of walk_coop_kernel: SMs x resident blocks per SM x 2 walkers per 128-thread block (a walker =
    one warp per condition).  Sharded runs keep 3.5 sigma of headroom: a rank's share of a step's k*world deaths is
    Binomial, and walkers beyond the resident wave would run as a second wave."""
    import math
    from gleipnir_b200 import _lib
    cap = _lib.device_info()[0] * like.network.coop_min_blocks() * 2
    return cap if world == 1 else int(cap - 3.5 * math.sqrt(cap))

C2 = dict(name="C2 eggcarton d=2 N=1e4/GPU RemainingPriorMass(1e-6)", d=2, population=10000,
                iterations=50, burn_in=0, tuning_cycles=2, cutoff=1e-6, truth_logz=235.88)


def death_count(n, cutoff):
    import math
    alpha = n / (n + 1)
    m = max(1, int(math.ceil(math.log(cutoff) / math.log(alpha))))
    while m > 1 and alpha ** (m - 1) <= cutoff:
        m -= 1
    while alpha ** m > cutoff:
        m += 1
    return m


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))



# ------------------------------------------------------------------------------------------
# CPU side: the SciPy stand-in for the reference's likelihood (oracle/ode.py), reference arm
# ------------------------------------------------------------------------------------------
# ScipyOdeSimulator's own default solver tolerances (scipy `ode` defaults behind pysb: rtol 1e-6,
# atol 1e-12); the parity tests use the same function at 1e-12.
CPU_RTOL, CPU_ATOL = 1e-6, 1e-12


_C5_CPU = {}


def _c5_cpu_setup():
    """Per process, once: the problem, the prior box and a first evaluation (imports, SciPy's LSODA wrapper warm)."""
    if _C5_CPU:
        return
    from gleipnir_b200 import jarm
    from oracle import ode
    like = jarm.jarm_loglikelihood()
    table = jarm.jarm_prior_table()
    _C5_CPU.update(ode=ode, desc=like.network.description(), proto=like.oracle_protocol(),
                   lo=np.array([t[0] for t in table]), sc=np.array([t[1] for t in table]))
    try:                                    # one BLAS thread per process: the processes already fill the cores
        from threadpoolctl import threadpool_limits
        _C5_CPU["blas"] = threadpool_limits(limits=1)
    except ImportError:                     # pragma: no cover
        pass
    _c5_cpu_chunk((12345, 1))


def _c5_cpu_chunk(args):
    """`n` evaluations of the JARM likelihood at prior draws with the SciPy stand-in; (evals, seconds)."""
    seed, n = args
    _c5_cpu_setup()
    c = _C5_CPU
    rng = np.random.default_rng(seed)
    Y = c["lo"] + c["sc"] * rng.uniform(size=(n, len(c["lo"])))
    t0 = time.perf_counter()
    for y in Y:
        c["ode"].protocol_loglike(c["desc"], c["proto"], y, rtol=CPU_RTOL, atol=CPU_ATOL)
    return n, time.perf_counter() - t0


class C5CpuPool(object):
    """Worker processes that live across steps: process start-up, imports and problem set-up stay outside the timed
    region (timing them would under-report the reference: a step is ~0.5 s of solves per process)."""

    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs, initializer=_c5_cpu_setup) if procs > 1 else None
        if self.pool is None:
            _c5_cpu_setup()
        else:
            self.pool.map(_c5_cpu_chunk, [(i, 1) for i in range(procs)])      # every worker up and warm

    def sample(self, n_per_proc, seed0=0):
        """(evaluations, wall seconds) of one step: n_per_proc evaluations in every process."""
        t0 = time.perf_counter()
        if self.pool is None:
            res = [_c5_cpu_chunk((seed0, n_per_proc))]
        else:
            res = self.pool.map(_c5_cpu_chunk, [(seed0 + i, n_per_proc) for i in range(self.procs)], chunksize=1)
        return sum(r[0] for r in res), time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def c5_cpu_sample(procs, n_per_proc, seed0=0):
    pool = C5CpuPool(procs)
    try:
        return pool.sample(n_per_proc, seed0)
    finally:
        pool.close()


C5_CPU_SAMPLE = ("SciPy stand-in for ScipyOdeSimulator (oracle/ode.py: odeint/LSODA, analytic Jacobian, rtol %g atol %g "
                 "= the reference solver's defaults; PySB itself is not installable here): JARM likelihood at prior "
                 "draws, %d persistent worker processes x %d evaluations per step (process start-up and problem set-up "
                 "outside the timed region)")


def run_reference_arm(args):
    """--impl reference: the reference's CPU path for the headline workload on all host cores.  The
    sampler's own cost is negligible next to 4 stiff solves per evaluation, so a step is
    a bounded batch of likelihood evaluations, one process per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_per_proc = 24
    per_step, total = [], 0
    pool = C5CpuPool(cores)                                  # start-up and set-up outside the timed steps
    for i in range(args.warmup + args.steps):
        ev, wall = pool.sample(n_per_proc, seed0=1000 * i)
        if i >= args.warmup:
            per_step.append(wall)
            total += ev
    pool.close()
    value = total / sum(per_step)
    sample = C5_CPU_SAMPLE % (CPU_RTOL, CPU_ATOL, cores, n_per_proc)
    line = dict(metric="logL evals/sec", value=value, unit="evals/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * float(np.mean(per_step)), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=C5["name"], sample=sample),
                cpu_baseline=dict(value=value, unit="evals/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


# C2 with the reference-as-shipped port and with the intended-algorithm oracle (BASELINE.md section 2 control)
def _c2_shipped_sample(args):
    seed, population, max_iter = args[:3]
    import warnings
    from scipy.stats import uniform
    from oracle import reference_port as rp
    w = C2
    np.random.seed(seed)
    pri = [uniform(loc=0.0, scale=10.0 * np.pi) for _ in range(w["d"])]

    def egg(x):
        return (2.0 + (np.cos(x)).prod()) ** 5
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = rp.run_shipped(pri, egg, population,
                           rp.ShippedSamplerState(iterations=w["iterations"], burn_in=w["burn_in"],
                                                  tuning_cycles=w["tuning_cycles"]),
                           prior_mass_cutoff=w["cutoff"], max_iterations=max_iter)
    return r["n_evals"], time.perf_counter() - t0


def c2_cpu_baselines(N, n_deaths):
    """Two single-core CPU figures for C2: the reference as shipped (scipy scalar calls on every trial move)
    and the intended algorithm with closed-form priors (oracle/engine.py) -- BASELINE.md section 2 asks for both."""
    w = C2
    sweeps = 20 * w["tuning_cycles"] + w["iterations"] + w["burn_in"]
    ev1, t1 = _c2_shipped_sample((0, N, 1))
    n_it = 300
    ev2, t2 = _c2_shipped_sample((0, N, n_it))
    t_iter = (t2 - t1) / (n_it - 1.0)
    evals_full = N + (n_deaths - 1) * (1 + sweeps * w["d"])
    full_s = t1 + (n_deaths - 1) * t_iter
    shipped = dict(value=evals_full / full_s, unit="evals/s", cores=1, kind="port",
                   sample="reference-as-shipped port (oracle/reference_port.py): N=%d init (%.2f s, %d evals) + %d "
                          "iterations (%.1f s, %d evals); value = full-run evals / (t_init + (n_deaths-1)*t_iteration)"
                          % (N, t1, ev1, n_it, t2, ev2),
                   init_evals_per_s=ev1 / t1, iteration_evals_per_s=(ev2 - ev1) / (t2 - t1),
                   time_to_converged_logZ_s_extrapolated=full_s)
    from oracle import priors as pr
    from oracle.engine import Engine
    from oracle.likelihoods import Model
    table = [(pr.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * w["d"]
    n_small, k_small = 500, 1
    t0 = time.perf_counter()
    eng = Engine(table, Model("eggcarton", w["d"]), n_small, batch_k=k_small, iterations=w["iterations"],
                 tuning_cycles=w["tuning_cycles"], seed=3)
    eng.run(60)
    dt = time.perf_counter() - t0
    ev = n_small + 59 * (1 + sweeps * w["d"])
    intended = dict(value=ev / dt, unit="evals/s", cores=1, kind="port",
                    sample="intended algorithm, closed-form priors (oracle/engine.py, the checker of the GPU engine): "
                           "N=%d, 60 iterations, %d evals in %.1f s" % (n_small, ev, dt))
    return shipped, intended


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def measure_dfma_peak(torch, lib, L):
    sm, _, _ = L.device_info()
    blocks, iters = sm * 16, 4096
    scratch = torch.empty(blocks * 256, dtype=torch.float64, device="cuda")
    flops = ctypes.c_int64(0)
    st = torch.cuda.current_stream().cuda_stream
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.gl_microbench_dfma(blocks, iters, scratch.data_ptr(), ctypes.byref(flops), st))
        e1.record()
        e1.synchronize()
        if i >= 2:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3))
    return best / 1e12


class Env(object):
    """Process-group plumbing shared by the workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        from gleipnir_b200 import _lib as L
        self.L, self.lib = L, L.load()
        self.comm = None
        if self.world > 1:
            from gleipnir_b200.sharded import get_nccl_comm
            self.comm = get_nccl_comm()
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")     # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(v) for v in t]


def c5_flops_per_step(net):
    """SURVEY 8d's algorithmic count per accepted step of an s-stage Rosenbrock method on n species and r
    reactions: 1 Jacobian (2r*order = 4r), 1 LU (2/3 n^3), s x (RHS 3r + solve 2 n^2); RODAS4 has s = 6.
    Also the FMAs the sparse generated code really executes per step (structural nonzeros only)."""
    n, r, s = net.ns, len(net.reactions), 6
    dense = 4.0 * r + 2.0 / 3.0 * n ** 3 + s * (3.0 * r + 2.0 * n * n)
    pl = net.plan()
    kr4 = pl.flops_executed(r)                         # counted for 4 stages (3 RHS) by odeplan
    per_solve = (kr4 - 4.0 * r - 4.0 * 3.0 * r) / 4.0 if False else None
    lu = 2.0 * sum(len(pl.leaf_rows[l]) * len(pl.leaf_cols[l]) for l in range(pl.nx))
    Q = pl.filled
    for c, cols in pl.lu_ops:
        rows = int(Q[c + 1:pl.nc, c].sum())
        lu += 2.0 * rows * len(cols) + rows
    solve = 2.0 * (pl.nnz_l + pl.nnz_u + pl.nc + sum(len(a) + len(b) for a, b in zip(pl.leaf_rows, pl.leaf_cols)))
    sparse = 4.0 * r + lu + s * (3.0 * r + solve)
    return dense, sparse


def finite_population(env, like, table, n, seed):
    """n prior draws CONDITIONED on a finite likelihood, as device tensors: the live set of a C5 run at the moment
    its first phase ends.  ~59 % of the prior box has likelihood -inf (the half-normal term of
    run_jarm_multinest.py:39-40 has no density below the measured value), so a run's first 0.59*N/k batches only
    clear those points under a wall of -inf; every batch after that -- thousands -- runs under a finite wall, and
    that is the state the timed steps start from."""
    torch = env.torch
    lo = torch.tensor([a for _, a, b in table], dtype=torch.float64, device="cuda")
    sc = torch.tensor([b for _, a, b in table], dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    xs, ls, have = [], [], 0
    while have < n:
        x = lo + sc * torch.rand((max(4096, n), len(table)), generator=g, dtype=torch.float64, device="cuda")
        l = like.evaluate(x)
        m = torch.isfinite(l)
        xs.append(x[m])
        ls.append(l[m])
        have += int(m.sum())
    return torch.cat(xs)[:n].contiguous(), torch.cat(ls)[:n].contiguous()


def run_c5(args, env):
    """Headline: K timed NS batches of the running C5 run.  Returns the pieces of the JSON line."""
    torch, L, world, rank = env.torch, env.L, env.world, env.rank
    from gleipnir_b200 import jarm
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.sharded import shard_bounds
    w = C5
    d, N = w["d"], w["population"]
    like = jarm.jarm_loglikelihood()
    k = (args.batch_k if args.batch_k else c5_walkers_per_gpu(like, world)) * world
    table = [(L.GL_PRIOR_UNIFORM, a, b) for a, b in jarm.jarm_prior_table()]
    n_steps = args.warmup + args.steps
    skw = dict(iterations=w["iterations"], burn_in=w["burn_in"], tuning_cycles=w["tuning_cycles"], proposal=0, seed=2024)
    cap = (n_steps + 2) * k
    n_prof = 1                                              # extra step(s) with events around every walker launch
    if world > 1:
        off, n_loc = shard_bounds(N, world, rank)
        eng = DeviceEngine(table, like, n_loc, k, cap, row_offset=off, global_population=N, **skw)
        env.comm.enable_p2p(k)
    else:
        eng = DeviceEngine(table, like, N, k, cap, **skw)
    st_ = eng._stream()

    def advance(total, so_far):
        """One step of the running run: refill the waiting batch (k walkers), kill the next k worst."""
        if world > 1:
            L.check(eng.lib.gl_ns_step_sharded(eng._h, env.comm.h, 1, int(k), int(so_far), 0.0, st_))
        else:
            if so_far > 0:
                eng.refill()
            eng.kill(k)
    t0 = time.perf_counter()
    x0, l0 = finite_population(env, like, table, eng.N, 11 + rank)      # setup, untimed
    eng.init(x0, l0)
    advance(k, 0)                                           # first kill: from here on a step = refill + kill
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    like.counters(reset=True)
    for i in range(args.warmup):
        advance((i + 2) * k, (i + 1) * k)
    env.barrier()
    like.counters(reset=True)
    clocks = ClockSampler(env.local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    env.barrier()
    times = []
    for i in range(args.warmup, n_steps):
        env.flush.zero_()                                   # evict L2 between timed steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        advance((i + 2) * k, (i + 1) * k)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    env.barrier()
    clk = clocks.stop() if rank == 0 else None
    ctr = like.counters(reset=True)
    st = eng.state()
    launches_per_step = w["tuning_cycles"] + 1 + 1          # walker phases + the kill / fused sharded step
    # per-kernel time for the roofline: events around every walker launch, in 2 extra steps
    eng.profile_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(n_steps, n_steps + n_prof):
        advance((i + 2) * k, (i + 1) * k)
    p1.record()
    torch.cuda.synchronize()
    prof_ms = p0.elapsed_time(p1)
    walk_ms, walk_launches, sel_ms, sel_launches = eng.profile_read()
    ctr_prof = like.counters(reset=True)
    eng.profile_enable(False)
    t_dev = env.reduce([sum(times) * 1e-3], "MAX")[0]
    evals, solves, steps_ode = env.reduce([ctr["evals"], ctr["solves"], ctr["steps"]], "SUM")
    nominal_evals = args.steps * k * (1 + (20 * w["tuning_cycles"] + w["iterations"] + w["burn_in"]) * d)
    return dict(like=like, eng=eng, k=k, N=N, t_dev=t_dev, evals=evals, solves=solves, ode_steps=steps_ode,
                nominal_evals=nominal_evals, clocks=clk, t_init=t_init, launches=args.steps * launches_per_step * world,
                walk_ms=walk_ms, walk_launches=walk_launches, sel_ms=sel_ms, sel_launches=sel_launches,
                ctr_prof=ctr_prof, state=st, ms_steps=times, n_prof=n_prof, prof_ms=prof_ms)


def run_c5_e2e(args, env, k):
    """The public call, bounded: NestedSampling(..., initial_live_points=host array).run() with NumberOfIterations(2k)
    at N = 10^5 -- priors, config and the 14.4 MB live set in from the host, its 10^5 likelihood evaluations, one
    refilled batch, finalisation, the dead-point table back on the host.  The live set is the finite-likelihood
    population of finite_population() (drawn outside the timed region)."""
    torch, world = env.torch, env.world
    from gleipnir_b200 import jarm
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations
    w = C5
    n_runs = max(1, min(args.steps, 2))
    t_tot, ev_tot, rows = 0.0, 0.0, 0
    like = jarm.jarm_loglikelihood()
    table = [(env.L.GL_PRIOR_UNIFORM, a, b) for a, b in jarm.jarm_prior_table()]
    x_host = finite_population(env, like, table, w["population"], 77)[0].cpu().numpy()      # same on every rank
    for i in range(n_runs):
        sps = jarm.jarm_sampled_parameters()
        ns = NestedSampling(sps, like, w["population"], sampler=MetropolisComponentWiseHardNSRejection(
            iterations=w["iterations"], burn_in=w["burn_in"], tuning_cycles=w["tuning_cycles"]),
            stopping_criterion=NumberOfIterations(2 * k), batch_k=k, seed=700 + i, initial_live_points=x_host)
        env.barrier()
        like.counters(reset=True)
        t0 = time.perf_counter()
        ns.run()
        torch.cuda.synchronize()
        t_tot += time.perf_counter() - t0
        ev_tot += like.counters(reset=True)["evals"]
        rows = len(ns.dead_points)
    t_max = env.reduce([t_tot], "MAX")[0]
    ev = env.reduce([ev_tot], "SUM")[0]
    d = w["d"]
    return dict(value=ev / t_max, unit="evals/s",
                h2d_bytes_per_step=int(24 * d + ctypes.sizeof(env.L.gl_ns_config) + w["population"] * d * 8),
                d2h_bytes_per_step=int(rows * 8 * (d + 2) + w["population"] * d * 8 + 2 * ctypes.sizeof(env.L.gl_ns_state)),
                ms_per_step=1e3 * t_max / n_runs, runs=n_runs,
                call="NestedSampling(jarm priors, jarm likelihood, 1e5, sampler(10, burn 10, tc 1), NumberOfIterations(2k), "
                     "batch_k=k, initial_live_points=<host array, finite-likelihood prior draws>).run(): live set host->device, "
                     "its 1e5 evaluations, one refilled batch, final term, dead_points on the host")


def run_c2(args, env, reps):
    """Secondary: converged C2 runs (eggcarton, N = 10^4 per GPU), device-timed; as round 1's headline."""
    torch, L, world, rank = env.torch, env.L, env.world, env.rank
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    from gleipnir_b200.sharded import run_sharded_native, shard_bounds
    w = C2
    d = w["d"]
    N = w["population"] * world
    k = N // 10
    n_deaths = death_count(N, w["cutoff"])
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * d
    like = EggCartonLogLikelihood(d)

    def one_run(seed, profile=False):
        skw = dict(iterations=w["iterations"], burn_in=w["burn_in"], tuning_cycles=w["tuning_cycles"], proposal=0, seed=seed)
        if world > 1:
            off, n_loc = shard_bounds(N, world, rank)
            eng = DeviceEngine(table, like, n_loc, k, n_deaths, row_offset=off, global_population=N, **skw)
        else:
            eng = DeviceEngine(table, like, N, k, n_deaths, **skw)
        if profile:
            eng.profile_enable(True)
        env.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            run_sharded_native(eng, env.comm, n_deaths, k)
        else:
            eng.init()
            eng.advance(n_deaths)
            eng.finalize()
        e1.record()
        e1.synchronize()
        return eng, e0.elapsed_time(e1)
    for i in range(3):
        one_run(i)
    env.barrier()
    times, evals, logzs = [], 0, []
    for i in range(reps):
        eng, ms = one_run(100 + i)
        st = eng.state()
        times.append(ms)
        evals += st.n_evals
        logzs.append(st.log_z)
    env.barrier()
    eng, ms = one_run(100, profile=True)
    walk_ms, walk_launches, sel_ms, sel_launches = eng.profile_read()
    t_dev = env.reduce([sum(times) * 1e-3], "MAX")[0]
    ev = env.reduce([float(evals)], "SUM")[0]
    out = dict(workload=w["name"], population_global=N, batch_k=k, n_deaths=n_deaths, runs=reps,
               value=ev / t_dev, unit="evals/s", time_to_converged_logZ_s=t_dev / reps,
               logZ_mean=float(np.mean(logzs)), logZ_std=float(np.std(logzs)), logZ_exact=235.85594,
               walker_avg_launch_us=1e3 * walk_ms / max(walk_launches, 1),
               select_avg_launch_us=1e3 * sel_ms / max(sel_launches, 1), select_share_of_step=sel_ms / ms,
               scaling="weak (1e4 live points per GPU)")
    if world == 1:
        out["replicas"] = run_c2_replicas(env, N, k, n_deaths)
    if world == 1 and rank == 0 and not args.no_cpu:
        shipped, intended = c2_cpu_baselines(N, n_deaths)
        out["cpu_baseline_as_shipped"] = shipped
        out["cpu_baseline_intended_algorithm"] = intended
    return out


def run_c2_replicas(env, N, k, n_deaths, n_rep=10):
    """C2 again as 10 independent seeds advanced CONCURRENTLY (gl_ns_advance_many: one engine + stream per seed) --
    the 10-seed evidence statistics north_star asks for, and what fills a device that one N = 10^4 run leaves idle."""
    torch, L = env.torch, env.L
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    w = C2
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * w["d"]
    like = EggCartonLogLikelihood(w["d"])
    best = None
    for rep in range(3):
        engs = [DeviceEngine(table, like, N, k, n_deaths, iterations=w["iterations"], burn_in=w["burn_in"],
                             tuning_cycles=w["tuning_cycles"], seed=1000 * rep + i, stream=torch.cuda.Stream())
                for i in range(n_rep)]
        torch.cuda.synchronize()
        handles = (ctypes.c_void_p * n_rep)(*[e._h for e in engs])
        streams = (ctypes.c_void_p * n_rep)(*[e._stream() for e in engs])
        deaths = (ctypes.c_int64 * n_rep)(*[n_deaths] * n_rep)
        env.flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for e in engs:
            e.init()
        L.check(env.lib.gl_ns_advance_many(handles, n_rep, deaths, streams))
        for e in engs:
            e.finalize()
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        states = [e.state() for e in engs]
        if best is None or ms < best[0]:
            best = (ms, states)
    ms, states = best
    lz = np.array([s.log_z for s in states])
    return dict(seeds=n_rep, wall_ms_all_runs=ms, ms_per_run_amortised=ms / n_rep,
                value=float(sum(s.n_evals for s in states)) / (ms * 1e-3), unit="evals/s",
                logZ_mean=float(lz.mean()), logZ_seed_scatter=float(lz.std(ddof=1)),
                logZ_err_reported=float(np.mean([np.sqrt(max(s.information, 0.0) / N) for s in states])),
                timing="CUDA events on the default stream around init + gl_ns_advance_many + finalize of all runs, "
                       "device synchronised on both sides; best of 3")


def run_c3(args, env):
    """Secondary: ONE converged C3 run (synthetic 50-D dense Gaussian, N = 10^6 in total, stop at ln X = -160):
    strong scaling -- the live set is sharded over the ranks."""
    torch, L, world, rank = env.torch, env.L, env.world, env.rank
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.sharded import shard_bounds
    d, N = 50, 1000000
    k = N // 10
    n_deaths = 160 * N
    like = lk.DenseGaussianLogLikelihood.synthetic_c3(d)
    table = [(L.GL_PRIOR_UNIFORM, -5.0, 10.0)] * d
    skw = dict(iterations=10, burn_in=0, tuning_cycles=1, proposal=0, seed=1)
    # keep only the scalar outputs: the dead log of a C3-scale run is tens of GB
    if world > 1:
        off, n_loc = shard_bounds(N, world, rank)
        eng = DeviceEngine(table, like, n_loc, k, 2 * k, row_offset=off, global_population=N, **skw)
        env.comm.enable_p2p(k)
    else:
        eng = DeviceEngine(table, like, N, k, 2 * k, **skw)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.init()
    if world > 1:
        L.check(eng.lib.gl_ns_advance_sharded(eng._h, env.comm.h, int(n_deaths), int(k), 0, 0.0, eng._stream()))
    else:
        eng.advance(n_deaths)
    e1.record()
    e1.synchronize()
    st = eng.state()
    t_dev = env.reduce([e0.elapsed_time(e1) * 1e-3], "MAX")[0]
    ev = env.reduce([float(st.n_evals)], "SUM")[0]
    return dict(workload="C3 synthetic 50-D correlated Gaussian N=1e6 (total), sampler(10, tc=1), stop at ln X = -160",
                population_global=N, batch_k=k, n_deaths=int(st.n_dead), value=ev / t_dev, unit="evals/s",
                time_to_converged_logZ_s=t_dev, logZ_without_final_term=st.log_z, logZ_truth=-d * float(np.log(10.0)),
                information=st.information, scaling="strong (N fixed, sharded over the ranks)",
                note="dead-point log not kept (tens of GB); evals = N + deaths*(1 + 30*d)")


def run_b200_arm(args):
    env = Env()
    torch, L, lib, world, rank = env.torch, env.L, env.lib, env.world, env.rank
    parity = None
    if world > 1:
        parity = sharded_parity_check(env)
    c5 = run_c5(args, env)
    e2e = run_c5_e2e(args, env, c5["k"])
    secondary = {}
    if not args.no_secondary:
        secondary["C2"] = run_c2(args, env, reps=5)
        try:
            secondary["C3"] = run_c3(args, env)
        except Exception as exc:                        # the headline must not die on a secondary config
            secondary["C3"] = dict(error=str(exc)[:300])
    env.barrier()
    if rank != 0:
        if world > 1:
            env.dist.destroy_process_group()
        return 0
    value = c5["evals"] / c5["t_dev"]
    like = c5["like"]
    net = like.network
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dfma_peak = measure_dfma_peak(torch, lib, L)
    dense_flop, sparse_flop = c5_flops_per_step(net)
    # roofline of the dominant kernel from the 2 profiled steps on this rank: ODE steps integrated inside the
    # walker launches / the CUDA-event time of those launches
    walk_s = c5["walk_ms"] * 1e-3
    steps_prof = c5["ctr_prof"]["steps"]
    achieved = steps_prof * dense_flop / walk_s / 1e12 if walk_s > 0 else None
    step_ms = float(np.mean(c5["ms_steps"]))
    roofline = dict(kernel="walk_coop_kernel<Net=jarm> (RODAS4, one system per warp, one warp per condition, 16 warps per SM)",
                    bound="fp64", achieved=achieved, peak=dfma_peak, unit="TFLOP/s",
                    frac=(achieved / dfma_peak) if achieved and dfma_peak else None, traffic=None,
                    traffic_source="not re-measured in this run (bound is fp64, not hbm); ncu --set full of the batch-evaluation "
                                   "form of the same solver code, profiles/r02_h_jarm_eval_register_gather_4blocks.txt: 3.1 MB of "
                                   "DRAM traffic in a 62 ms launch = 1e-5 of the HBM roof (the state lives in registers and "
                                   "shared memory); the walker launch itself: profiles/r02_m_walk_coop_kernel_c5_production_launch.txt",
                    peak_source="gl_microbench_dfma measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                    algorithmic_flop_per_ode_step=dense_flop,
                    algorithmic_flop_formula="SURVEY 8d: 4r + 2/3 n^3 + s(3r + 2n^2), n=34 species, r=84 reactions, s=6 stages",
                    executed_sparse_flop_per_ode_step=sparse_flop,
                    executed_sparse_tflops=(steps_prof * sparse_flop / walk_s / 1e12) if walk_s > 0 else None,
                    ode_steps_per_s=steps_prof / walk_s if walk_s > 0 else None,
                    ode_steps_per_solve=c5["ode_steps"] / max(c5["solves"], 1),
                    solves_per_eval=c5["solves"] / max(c5["evals"], 1),
                    rhs_evals_per_s=6.0 * steps_prof / walk_s if walk_s > 0 else None,
                    avg_launch_ms=c5["walk_ms"] / max(c5["walk_launches"], 1), launches_timed=int(c5["walk_launches"]),
                    walker_share_of_step=c5["walk_ms"] / c5["prof_ms"],
                    select_avg_launch_us=1e3 * c5["sel_ms"] / max(c5["sel_launches"], 1),
                    timing="CUDA events around the walker launches of %d extra step(s) after the timed region" % c5["n_prof"],
                    note="latency-bound: a walker is a chain of 721 dependent likelihood evaluations, each 2 x (pre-equilibration "
                         "+ main) stiff solves of ~10^2-10^3 dependent Rosenbrock steps; fp64 pipe utilisation and stall "
                         "reasons from ncu are in profiles/r02_h_* (batch-evaluation form) and profiles/r02_m_* (a real walker launch)")
    cpu = None
    if world == 1 and not args.no_cpu:
        ev, wall = c5_cpu_sample(1, 400)
        cpu = dict(value=ev / wall, unit="evals/s", cores=1, kind="port",
                   sample=C5_CPU_SAMPLE % (CPU_RTOL, CPU_ATOL, 1, 400) + " (%.1f s)" % wall)
    w = C5
    line = dict(metric="logL evals/sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * c5["t_dev"] / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=w["name"], population_global=c5["N"], batch_k=c5["k"], walkers_per_gpu_per_step=c5["k"] // world,
                            sampler="MetropolisComponentWiseHardNSRejection(iterations=10, burn_in=10, tuning_cycles=1)",
                            ode="RODAS4, rtol %g atol %g" % (like.rtol, like.atol),
                            evals_counted="integrated on the device (kernel counter); the reference's own count, which "
                                          "includes trial moves rejected by the prior alone, is nominal_evals",
                            nominal_evals=c5["nominal_evals"], setup_s=c5["t_init"],
                            start_state="live set = prior draws conditioned on a finite likelihood (the state of a run when its "
                                        "-inf phase ends: 59 % of the prior box lies outside the half-normal term's support)",
                            l2="flushed between timed steps (256 MB memset)",
                            parallelism=("live set sharded over %d ranks (N/world each); per step one fused kernel exchanges the "
                                         "candidate logL over NVLink peer memory; k = %d walkers per GPU (weak in k)" % (world, c5["k"] // world))
                            if world > 1 else "single"),
                clocks=c5["clocks"], gpu_launches=int(c5["launches"]), e2e=e2e, roofline=roofline, cpu_baseline=cpu,
                time_to_converged_logZ_s=None,
                time_to_converged_note="a converged C5 run at N=1e5 is ~N*(H + ln(1/X_stop))/k batches (hours on one GPU) "
                                       "and is not measured here; converged runs: secondary.C2, secondary.C3",
                secondary=secondary)
    if parity is not None:
        line["parity_check"] = parity
    print(json.dumps(line))
    if world > 1:
        env.dist.destroy_process_group()
    return 0


def sharded_parity_check(env):
    """world > 1 only (the driver's GPU test box has one GPU): before timing, the fused peer-memory step, the NCCL
    all-gather path and the CPU oracle of the sharded algorithm (oracle/sharded.py, every rank's shard run through
    the same SPMD protocol) on a small eggcarton run -- this rank's dead rows compared slot for slot."""
    L, world, rank = env.L, env.world, env.rank
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    from gleipnir_b200.sharded import LocalComm, run_sharded, run_sharded_native, shard_bounds
    from oracle import priors as pr
    from oracle.likelihoods import Model
    from oracle.sharded import OracleShard
    d, N, k, n_deaths, seed = 2, 4000, 400, 2400, 17
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * d
    like = EggCartonLogLikelihood(d)
    off, n_loc = shard_bounds(N, world, rank)
    res = {}
    for mode in ("nccl", "p2p"):                       # NCCL first: the peer mapping, once made, stays in use
        if mode == "nccl":
            os.environ["GLEIPNIR_B200_NO_P2P"] = "1"
        else:
            os.environ.pop("GLEIPNIR_B200_NO_P2P", None)
        eng = DeviceEngine(table, like, n_loc, k, n_deaths, row_offset=off, global_population=N, iterations=4,
                           tuning_cycles=1, seed=seed)
        run_sharded_native(eng, env.comm, n_deaths, k, candidate_sigmas=0.0)
        tab, idx, step = eng.dead_table()
        res[mode] = (tab, idx, step, eng.state().log_z, bool(env.lib.gl_comm_p2p_ready(env.comm.h)))
    same = bool(np.array_equal(res["p2p"][0], res["nccl"][0]) and np.array_equal(res["p2p"][1], res["nccl"][1]))
    ora = []
    for rk in range(world):
        o, nl = shard_bounds(N, world, rk)
        ora.append(OracleShard([(pr.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * d, Model("eggcarton", d), nl, batch_k=k,
                               seed=seed, row_offset=o, global_population=N, iterations=4, tuning_cycles=1))
    run_sharded(ora, LocalComm(world), n_deaths, k, candidate_sigmas=0.0)
    want = ora[rank].result
    tab, idx, step = res["p2p"][:3]
    nd = len(want["dead_index"])
    od = np.lexsort((idx[:nd], step[:nd]))
    oo = np.lexsort((want["dead_index"], want["dead_batch"]))
    ok_oracle = bool(np.array_equal(idx[:nd][od], want["dead_index"][oo]) and
                     np.allclose(tab[:nd][od], want["dead"][:nd][oo], rtol=1e-12, atol=1e-13) and
                     abs(res["p2p"][3] - want["logZ"]) <= 1e-12 * abs(want["logZ"]))
    flags = env.reduce([1.0 if same else 0.0, 1.0 if ok_oracle else 0.0, 1.0 if res["p2p"][4] else 0.0,
                        0.0 if res["nccl"][4] else 1.0], "MIN")
    return dict(config="eggcarton N=%d k=%d deaths=%d world=%d, full exchange" % (N, k, n_deaths, world),
                peer_memory_equals_nccl_path_bitwise=bool(flags[0]), equals_cpu_oracle_shards=bool(flags[1]),
                peer_memory_path_ran=bool(flags[2]), nccl_path_ran=bool(flags[3]), logZ=res["p2p"][3],
                tolerance="slots identical, values 1e-12 relative")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-k", type=int, default=0,
                    help="C5 walkers per GPU per step (default: one resident wave, 1184 on a 148-SM part)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C3 secondary configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
