"""Dense-Gaussian DMMA batch kernel: parity vs numpy and throughput vs the measured DMMA peak."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                             # noqa: E402
from gleipnir_b200 import _lib as L                      # noqa: E402
from gleipnir_b200 import likelihoods as lk              # noqa: E402

lib = L.load()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


sm, _, _ = L.device_info()
scratch = torch.empty(sm * 16 * 256, dtype=torch.float64, device="cuda")
fl = ctypes.c_int64(0)
t = timeit(lambda: L.check(lib.gl_microbench_dmma(sm * 16, 4096, scratch.data_ptr(), ctypes.byref(fl), st)))
peak = fl.value / t / 1e12
for d in (50, 12, 33):
    n = 10 ** 6
    like = lk.DenseGaussianLogLikelihood.synthetic_c3(d)
    xs = torch.empty((n, d), dtype=torch.float64, device="cuda").uniform_(-5, 5)
    got = like.evaluate(xs[:1003]).cpu().numpy()
    r = xs[:1003].cpu().numpy() - like.mu[None, :]
    want = -0.5 * np.einsum("ni,ij,nj->n", r, like.precision, r) + like.log_norm
    err = np.max(np.abs(got - want) / np.abs(want))
    o = torch.empty(n, dtype=torch.float64, device="cuda")
    m = like.device_model()
    t = timeit(lambda: L.check(lib.gl_loglike_eval(m, xs.data_ptr(), n, o.data_ptr(), st)))
    flop = (2.0 * d * d + 2.0 * d) * n
    print("d=%d: max rel err %.2e | %.3f ms, %.2f TFLOP/s = %.1f %% of the measured DMMA peak (%.1f), %.0f GB/s"
          % (d, err, t * 1e3, flop / t / 1e12, 100 * flop / t / 1e12 / peak, peak, (8.0 * d + 8) * n / t / 1e9))
