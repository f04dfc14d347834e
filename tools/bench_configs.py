#!/usr/bin/env python
"""Secondary measurements for the parity-test configs (not bench.py lines): prior transforms and
k-worst selection against HBM peak, DFMA/DMMA peaks, dense-Gaussian batch kernel (C3),
ODE batch evaluation (C4), and converged runs of C1.  Writes one JSON document."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(torch, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    import torch
    from gleipnir_b200 import _lib as L
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    lib = L.load()
    out = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    sm = L.device_info()[0]
    st = None
    # ---- FP64 peaks -------------------------------------------------------------------------
    scratch = torch.empty(sm * 16 * 256, dtype=torch.float64, device="cuda")
    fl = ctypes.c_int64(0)
    t = timeit(torch, lambda: L.check(lib.gl_microbench_dfma(sm * 16, 4096, scratch.data_ptr(), ctypes.byref(fl), st)))
    out["dfma_peak_tflops"] = fl.value / t / 1e12
    t = timeit(torch, lambda: L.check(lib.gl_microbench_dmma(sm * 16, 4096, scratch.data_ptr(), ctypes.byref(fl), st)))
    out["dmma_peak_tflops"] = fl.value / t / 1e12
    # ---- prior draw / ppf / pdf (HBM bound) -----------------------------------------------------
    n = 1 << 27                                         # 1 GiB of doubles: larger than L2
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for name, row in (("uniform", (0, -5.0, 10.0)), ("norm", (1, 0.5, 2.0)), ("loguniform", (2, 1e-3, 1e2))):
        pa = L.prior_array([row])
        t = timeit(torch, lambda: L.check(lib.gl_prior_draw(pa, 1, 7, 4, 0, n, x.data_ptr(), st)), reps=5)
        out["prior_draw_%s" % name] = dict(GBps=8.0 * n / t / 1e9, frac_of_hbm=8.0 * n / t / 1e9 / hbm)
        torch.rand(n, dtype=torch.float64, out=y) if False else y.uniform_(1e-6, 1 - 1e-6)
        t = timeit(torch, lambda: L.check(lib.gl_prior_ppf(pa, y.data_ptr(), x.data_ptr(), n, st)), reps=5)
        out["prior_ppf_%s" % name] = dict(GBps=16.0 * n / t / 1e9, frac_of_hbm=16.0 * n / t / 1e9 / hbm)
        t = timeit(torch, lambda: L.check(lib.gl_prior_pdf(pa, x.data_ptr(), y.data_ptr(), n, st)), reps=5)
        out["prior_pdf_%s" % name] = dict(GBps=16.0 * n / t / 1e9, frac_of_hbm=16.0 * n / t / 1e9 / hbm)
    del y
    # ---- k-worst selection --------------------------------------------------------------------------
    for nn, k in ((10 ** 6, 10 ** 5), (1 << 27, 10 ** 5)):
        v = x[:nn].normal_()
        idx = torch.empty(k, dtype=torch.int64, device="cuda")
        ls = torch.empty(1, dtype=torch.float64, device="cuda")
        ws = torch.empty(int(lib.gl_select_workspace_bytes(nn)), dtype=torch.uint8, device="cuda")
        t = timeit(torch, lambda: L.check(lib.gl_select_kworst(v.data_ptr(), nn, k, idx.data_ptr(), ls.data_ptr(),
                                                               ws.data_ptr(), st)), reps=5)
        out["select_n%d_k%d" % (nn, k)] = dict(us=t * 1e6, GBps=8.0 * nn * 8 / t / 1e9,
                                               frac_of_hbm=8.0 * nn * 8 / t / 1e9 / hbm, passes=8)
    del x
    # ---- C3: dense Gaussian batch evaluation, DMMA vs scalar ---------------------------------------------
    d, n3 = 50, 10 ** 6
    like = lk.DenseGaussianLogLikelihood.synthetic_c3(d)
    xs = torch.empty((n3, d), dtype=torch.float64, device="cuda").uniform_(-5, 5)
    o = torch.empty(n3, dtype=torch.float64, device="cuda")
    m = like.device_model()
    flop = (2.0 * d * d + 2.0 * d) * n3
    t = timeit(torch, lambda: L.check(lib.gl_loglike_eval(m, xs.data_ptr(), n3, o.data_ptr(), st)))
    out["c3_dense_eval_dmma"] = dict(ms=t * 1e3, evals_per_s=n3 / t, tflops=flop / t / 1e12,
                                     frac_of_dmma_peak=flop / t / 1e12 / out["dmma_peak_tflops"],
                                     GBps=(8.0 * d + 8) * n3 / t / 1e9)
    L.check(lib.gl_model_set_tensor_path(m, 0))
    t = timeit(torch, lambda: L.check(lib.gl_loglike_eval(m, xs.data_ptr(), n3, o.data_ptr(), st)))
    out["c3_dense_eval_scalar"] = dict(ms=t * 1e3, evals_per_s=n3 / t, tflops=flop / t / 1e12)
    L.check(lib.gl_model_set_tensor_path(m, 1))
    del xs, o
    # ---- C3: a slice of the run (N = 10^6, k = 10^5, 3 steps) ------------------------------------------------
    eng = DeviceEngine([(0, -5.0, 10.0)] * d, m, 10 ** 6, 10 ** 5, 3 * 10 ** 5, iterations=10, tuning_cycles=1, seed=1)
    eng.profile_enable(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.init(); eng.advance(3 * 10 ** 5)
    s = eng.state()
    t1 = time.perf_counter() - t0
    wm, wl, sm_, sl = eng.profile_read()
    out["c3_run_slice"] = dict(N=10 ** 6, k=10 ** 5, steps=3, wall_s=t1, evals=int(s.n_evals),
                               evals_per_s=s.n_evals / t1, walk_ms_per_launch=wm / max(wl, 1),
                               select_ms_per_step=sm_ / 3)
    del eng
    # ---- C4: ODE batch of 10^5 parameter sets ---------------------------------------------------------
    from gleipnir_b200 import networks
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                             "ref_ode.npz"))
    ol = OdeNetworkLogLikelihood(networks.dimerization(), ["kf", "kr"], g["dimerization/tspan"],
                                 {"A_dimer": (g["dimerization/data"], 10.0 * g["dimerization/sd"], None)})
    n4 = 10 ** 5
    ya = torch.empty((n4, 2), dtype=torch.float64, device="cuda")
    L.check(lib.gl_prior_draw(L.prior_array([(1, -3.0, 1.0), (1, 0.0, 1.0)]), 2, 0, 4, 0, n4, ya.data_ptr(), st))
    oo = torch.empty(n4, dtype=torch.float64, device="cuda")
    mo = ol.device_model()
    t = timeit(torch, lambda: L.check(lib.gl_loglike_eval(mo, ya.data_ptr(), n4, oo.data_ptr(), st)), reps=5)
    out["c4_ode_batch_1e5"] = dict(ms=t * 1e3, evals_per_s=n4 / t, finite=int(torch.isfinite(oo).sum().item()))
    # CPU stand-in on a sample of 200
    from oracle import ode
    desc = networks.dimerization().description()
    rows = [(1, g["dimerization/data"], 10.0 * g["dimerization/sd"], np.ones(51))]
    yh = ya[:200].cpu().numpy()
    t0 = time.perf_counter()
    for yv in yh:
        ode.loglike(desc, [0, 1], yv, g["dimerization/tspan"], rows)
    out["c4_ode_cpu_scipy_standin"] = dict(evals_per_s=200 / (time.perf_counter() - t0), cores=1,
                                           note="SciPy LSODA stand-in for ScipyOdeSimulator (PySB absent)")
    # ---- C1: converged runs --------------------------------------------------------------------------------
    for dd in (2, 5):
        lk1 = lk.IIDGaussianLogLikelihood(dd)
        def run():
            e = DeviceEngine([(0, -5.0, 10.0)] * dd, lk1.device_model(), 500, 50, 5000, iterations=10, tuning_cycles=1,
                             seed=3)
            e.init(); e.advance(5000); e.finalize()
            return e
        t = timeit(torch, run, reps=5)
        e = run()
        out["c1_gauss_d%d" % dd] = dict(ms_to_converged=t * 1e3, evals=int(e.state().n_evals),
                                        evals_per_s=e.state().n_evals / t, logZ=e.state().log_z)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
