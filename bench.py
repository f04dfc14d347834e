#!/usr/bin/env python
"""bench.py -- logL evals/s and time-to-converged logZ of the B200 Nested Sampling path.

Workload (BASELINE.json configs[1], concretised in SURVEY 8d "C2"): 2-D egg-carton
log-likelihood (2 + cos x cos y)^5, priors uniform(0, 10*pi), population 10^4 per GPU,
MetropolisComponentWiseHardNSRejection(iterations=50, tuning_cycles=2), stop at
RemainingPriorMass(1e-6) -> 138 163 deaths, truth ln Z ~= 235.88.

A "step" is ONE whole converged run: initial population (Philox draw + logL), every
batch (k-worst selection, evidence/information update, walker refill) and the final
live-point term.  `value` = logL evaluations of the run / device time (inputs resident);
`ms_per_step` = time-to-converged ln Z.  `e2e` = the same through the public
`NestedSampling(...).run()` call: prior table + config go host->device, the dead-point
log comes back device->host, inside the timed region.

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

WORKLOAD = dict(name="C2 eggcarton d=2 N=1e4/GPU RemainingPriorMass(1e-6)", d=2, population=10000,
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
# reference arm / cpu baseline: the as-shipped restatement on host cores
# ------------------------------------------------------------------------------------------
def _cpu_sample(args):
    """One bounded sample of the workload with the reference-as-shipped port; (evals, seconds)."""
    seed, population, max_iter = args[:3]
    import warnings
    from scipy.stats import uniform
    from oracle import reference_port as rp
    w = WORKLOAD
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


def cpu_sample_parallel(procs, population, max_iter, seed0=0):
    import multiprocessing as mp
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_sample((seed0, population, max_iter))]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_sample, [(seed0 + i, population, max_iter) for i in range(procs)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle/reference_port.py, pinned
    bit-for-bit to the real reference) on all host cores, one process per seed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOAD
    cores = os.cpu_count() or 1
    population, max_iter = w["population"], 120    # 10^4 rvs + 10^4 logL + 119 sampler calls per process
    per_step = []
    total_evals = 0
    for i in range(args.warmup + args.steps):
        ev, wall = cpu_sample_parallel(cores, population, max_iter, seed0=1000 * i)
        if i >= args.warmup:
            per_step.append(wall)
            total_evals += ev
    value = total_evals / sum(per_step)
    n_deaths = death_count(population, w["cutoff"])
    evals_full = population + (n_deaths - 1) * (1 + (20 * w["tuning_cycles"] + w["iterations"] + w["burn_in"]) * w["d"])
    line = dict(metric="logL evals/sec", value=value, unit="evals/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * float(np.mean(per_step)), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=w["name"], sample="%d processes x (N=%d init + %d iterations)" %
                            (cores, population, max_iter)),
                cpu_baseline=dict(value=value, unit="evals/s", cores=cores, kind="port",
                                  sample="reference-as-shipped port (oracle/reference_port.py, bit-exact vs the "
                                         "real reference on golden runs): %d procs x (N=%d init + %d iterations)" %
                                         (cores, population, max_iter)),
                e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                time_to_converged_logZ_s_extrapolated=evals_full / (value / cores),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


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


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gleipnir_b200 import _lib as L
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    lib = L.load()
    w = WORKLOAD
    d = w["d"]
    N = w["population"] * world            # weak scaling: 10^4 live points per GPU, one global run
    k = args.batch_k if args.batch_k else N // 10
    n_deaths = death_count(N, w["cutoff"])
    from gleipnir_b200.sharded import get_nccl_comm, run_sharded_native, shard_bounds
    comm = get_nccl_comm() if world > 1 else None
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * d
    like = EggCartonLogLikelihood(d)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_run(seed, profile=False):
        skw = dict(iterations=w["iterations"], burn_in=w["burn_in"], tuning_cycles=w["tuning_cycles"], proposal=0,
                   seed=seed, walker_lanes=args.walker_lanes)
        if world > 1:                      # one shard per rank; L* by all-gather (NCCL)
            off, n_loc = shard_bounds(N, world, rank)
            eng = DeviceEngine(table, like, n_loc, k, n_deaths, row_offset=off, global_population=N,
                               **skw)
        else:
            eng = DeviceEngine(table, like, N, k, n_deaths, **skw)
        if profile:
            eng.profile_enable(True)
        flush.zero_()                      # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            run_sharded_native(eng, comm, n_deaths, k)
        else:
            eng.init()
            eng.advance(n_deaths)
            eng.finalize()
        e1.record()
        e1.synchronize()
        return eng, e0.elapsed_time(e1)

    for i in range(args.warmup):
        one_run(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)                    # let nvidia-smi deliver its first samples
    barrier()                              # (rank 0 slept: re-align the ranks before timing)
    times, evals, launches, logzs, errs = [], 0, 0, [], []
    walk_ms = walk_launches = sel_ms = sel_launches = 0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        eng, ms = one_run(100 + i, profile=False)
        st = eng.state()
        times.append(ms)
        evals += st.n_evals
        launches += st.n_launches
        logzs.append(st.log_z)
        errs.append(float(np.sqrt(max(st.information, 0.0) / N)))
        last_rows = int(st.n_rows)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    # per-kernel durations for the roofline: CUDA events around every walker / selection launch.  An
    # event record between two launches defeats their programmatic-dependent-launch overlap, so these
    # run as extra steps of the same workload right after the timed region, not inside it.
    prof_total_ms = 0.0
    for i in range(min(args.steps, 3)):
        eng, ms = one_run(100 + i, profile=True)
        a, b, c, e = eng.profile_read()
        walk_ms += a; walk_launches += b; sel_ms += c; sel_launches += e
        prof_total_ms += ms
    barrier()

    # ---- end-to-end through the public API (host in, host out) -----------------------------
    from scipy.stats import uniform
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import RemainingPriorMass
    from gleipnir_b200.sampled_parameter import SampledParameter

    def one_e2e(seed):
        sps = [SampledParameter(i, uniform(loc=0.0, scale=10.0 * np.pi)) for i in range(d)]
        ns = NestedSampling(sps, like, N, sampler=MetropolisComponentWiseHardNSRejection(
            iterations=w["iterations"], burn_in=w["burn_in"], tuning_cycles=w["tuning_cycles"]),
            stopping_criterion=RemainingPriorMass(w["cutoff"]), batch_k=k, seed=seed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ns.run()
        torch.cuda.synchronize()
        return ns, time.perf_counter() - t0
    for i in range(min(args.warmup, 2)):
        one_e2e(i)
    barrier()
    e2e_t, e2e_evals = 0.0, 0
    for i in range(args.steps):
        ns, t = one_e2e(500 + i)
        e2e_t += t
        e2e_evals += ns._n_evals / world           # sharded runs report the global count on every rank
    barrier()
    clk = clocks.stop() if rank == 0 else None      # sampled over both timed regions

    # max over ranks (device time), sum of evals over ranks
    dev_t = torch.tensor([sum(times) * 1e-3, e2e_t], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(evals), float(e2e_evals), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    t_dev, t_e2e = float(dev_t[0]), float(dev_t[1])
    value = float(tot[0]) / t_dev
    e2e_value = float(tot[1]) / t_e2e

    # ---- roofline of the dominant kernel (walker) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dfma_peak = measure_dfma_peak(torch, lib, L)
    sweeps = 20 * w["tuning_cycles"] + w["iterations"] + w["burn_in"]
    walk_evals_per_step = (k / world) * d * sweeps          # walkers of this rank (k/world on average)
    n_walk_steps = walk_launches / (w["tuning_cycles"] + 1)
    avg_launch_s = (walk_ms * 1e-3) / max(walk_launches, 1)
    flop_per_eval = FLOP_PER_TRIAL_EGGCARTON_D2
    achieved = (walk_evals_per_step / (w["tuning_cycles"] + 1)) * flop_per_eval / avg_launch_s / 1e12
    # dependent-issue latency of DFMA: what actually bounds a Metropolis chain (one warp per walker, ~25
    # dependent FP64 operations per round), as opposed to the pipe's throughput peak above
    lat = ctypes.c_double(0.0)
    lat_scratch = torch.zeros(64, dtype=torch.float64, device="cuda")
    L.check(lib.gl_microbench_dfma_latency(4096, lat_scratch.data_ptr(), ctypes.byref(lat)))
    hbm_peak = peaks.get("hbm_gbs")
    walk_bytes = (k / world) * 8.0 * (2 * d + 1)        # algorithmic: seed row in, point + logL out per replacement
    roofline = dict(kernel="walk_reg_kernel<LikeEggCarton, 2>", bound="fp64", achieved=achieved, peak=dfma_peak,
                    unit="TFLOP/s", frac=achieved / dfma_peak if dfma_peak else None,
                    traffic=158208, traffic_source="dram__bytes_read.sum + dram__bytes_write.sum of one production "
                    "launch, ncu --set full (profiles/r01_g_walk_reg_kernel_c2.txt); algorithmic bytes per launch: "
                    "%d" % int(walk_bytes),
                    peak_source="gl_microbench_dfma measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                    flop_per_trial_move=flop_per_eval,
                    dependent_dfma_latency_cycles=lat.value,
                    note="latency-bound by construction: a Metropolis chain is one dependent sequence and C2 offers "
                         "k*G = 32 000 threads to a 303 000-thread machine; neither HBM nor a tensor pipe is the limiter",
                    walker_hbm=dict(bound="hbm", achieved=walk_bytes / avg_launch_s / 1e9, peak=hbm_peak, unit="GB/s",
                                    frac=(walk_bytes / avg_launch_s / 1e9 / hbm_peak) if hbm_peak else None),
                    avg_launch_us=avg_launch_s * 1e6, launches_timed=int(walk_launches),
                    walker_share_of_step=walk_ms / prof_total_ms, select_share_of_step=sel_ms / prof_total_ms,
                    timing="CUDA events around every launch in %d extra steps after the timed region "
                           "(%.3f ms/step with the events in place)" % (min(args.steps, 3), prof_total_ms / min(args.steps, 3)),
                    select_avg_launch_us=1e3 * sel_ms / max(sel_launches, 1),
                    select_hbm=dict(bound="hbm", achieved=((8.0 * N / world) / (1e-3 * sel_ms / sel_launches) / 1e9)
                                    if sel_ms > 0 and sel_launches else None,
                                    peak=peaks.get("hbm_gbs"), unit="GB/s"))
    # ---- CPU baseline: bounded sample on one host core ---------------------------------------
    # init (N rvs + N logL) and the per-iteration sampler cost are timed separately so that the
    # extrapolation to the full run is t_init + (n_deaths - 1) * t_iteration
    cpu = None
    if world == 1:
        ev1, t1 = cpu_sample_parallel(1, N, 1)
        n_it = 600                                  # ~12 s of host work on one core
        cpu_evals, cpu_wall = cpu_sample_parallel(1, N, n_it)
        t_iter = (cpu_wall - t1) / (n_it - 1.0)
        evals_full = N + (n_deaths - 1) * (1 + sweeps * d)
        cpu_full_s = t1 + (n_deaths - 1) * t_iter
        cpu_value = evals_full / cpu_full_s
        cpu = dict(value=cpu_value, unit="evals/s", cores=1, kind="port",
                   sample="reference-as-shipped port (oracle/reference_port.py), N=%d init (%.2f s) + "
                          "%d iterations (%d evals, %.1f s); value = full-run evals / (t_init + "
                          "(n_deaths-1)*t_iteration)" % (N, t1, n_it, cpu_evals, cpu_wall),
                   time_to_converged_logZ_s_extrapolated=cpu_full_s)
    line = dict(metric="logL evals/sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=w["name"], population_per_gpu=w["population"], population_global=N, batch_k=k, n_deaths=n_deaths,
                            walker_lanes=args.walker_lanes or "auto",
                            sampler="MetropolisComponentWiseHardNSRejection(iterations=50, tuning_cycles=2)",
                            l2="flushed between timed runs (256 MB memset)",
                            parallelism=("live set sharded over %d ranks; per step one fused kernel exchanges "
                                         "~k/world + 8 sigma candidate logL per rank over NVLink peer memory" % world)
                            if world > 1 else "single"),
                clocks=clk, gpu_launches=int(float(tot[2])),
                e2e=dict(value=e2e_value, unit="evals/s", h2d_bytes_per_step=int(24 * d + ctypes.sizeof(L.gl_ns_config)),
                         # dead-point table + final live points + two state snapshots
                         d2h_bytes_per_step=int(last_rows * world * 8 * (d + 2) + N * d * 8 + 2 * ctypes.sizeof(L.gl_ns_state)),
                         ms_per_step=1e3 * t_e2e / args.steps),
                roofline=roofline,
                cpu_baseline=cpu,
                time_to_converged_logZ_s=t_dev / args.steps,
                logZ_mean=float(np.mean(logzs)), logZ_std=float(np.std(logzs)), logZ_err_reported=float(np.mean(errs)),
                logZ_truth=w["truth_logz"])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# fp64 operations per trial move of walk_kernel<LikeEggCarton>, d=2 (DESIGN.md "Roofline"):
# ncu instruction mix of the r01 baseline capture (profiles/r01_a_walk_kernel_c2_baseline.txt):
# 57 DFMA + 14 DMUL + 5 DADD per trial move = 2*57 + 14 + 5 = 133 flop (2 cos, 2 prior pdfs,
# ratio, proposal, (2+chi)^5, u52 conversions).
FLOP_PER_TRIAL_EGGCARTON_D2 = 133.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-k", type=int, default=0)
    ap.add_argument("--walker-lanes", type=int, default=0, help="lanes per walker (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
