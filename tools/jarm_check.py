"""C5 sanity run on the GPU: protocol likelihood vs the SciPy oracle, work counters, a timed batch
evaluation and a few NS steps.  python tools/jarm_check.py [n_batch] [N] [steps]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import torch                                        # noqa: E402
from gleipnir_b200 import jarm, networks            # noqa: E402
from gleipnir_b200.engine import DeviceEngine       # noqa: E402
from oracle import ode                               # noqa: E402
from oracle import priors as pr                      # noqa: E402


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    like = jarm.jarm_loglikelihood(rtol=float(os.environ.get('JARM_RTOL', '1e-8')))
    net = like.network
    desc = net.description()
    proto = like.oracle_protocol()
    nominal = np.log10(net.param_values[networks.JARM_CALIBRATED])
    rng = np.random.default_rng(1)
    Y = nominal[None, :] + rng.uniform(-1.0, 1.0, size=(8, 18))
    Y[0] = nominal
    t0 = time.time()
    got = like.evaluate(Y)
    print("first eval (compile+load) s", time.time() - t0)
    want = np.array([ode.protocol_loglike(desc, proto, y) for y in Y])
    print("gpu ", got)
    print("cpu ", want)
    print("rel ", np.abs(got - want) / np.abs(want))
    print("counters", like.counters(reset=True))
    table = jarm.jarm_prior_table()
    lo = np.array([t[0] for t in table]); sc = np.array([t[1] for t in table])
    X = torch.as_tensor(lo + sc * rng.uniform(size=(nb, 18)), device="cuda")
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.time()
        out = like.evaluate(X)
        torch.cuda.synchronize()
        dt = time.time() - t0
        c = like.counters(reset=True)
        print("batch %d prior draws: %.3f s, %.0f evals/s, finite %.3f, solves/eval %.2f, steps/solve %.1f, steps/s %.3g"
              % (nb, dt, nb / dt, float(torch.isfinite(out).float().mean()), c["solves"] / c["evals"],
                 c["steps"] / c["solves"], c["steps"] / dt))
    # CPU oracle on the same draws (bounded)
    t0 = time.time(); n_cpu = 0
    xs = X[:64].cpu().numpy()
    cpu = []
    for y in xs:
        cpu.append(ode.protocol_loglike(desc, proto, y)); n_cpu += 1
        if time.time() - t0 > 20: break
    dt = time.time() - t0
    g = out[:n_cpu].cpu().numpy(); cpu = np.array(cpu)
    fin = np.isfinite(cpu) & np.isfinite(g)
    print("oracle: %d evals in %.1f s = %.1f evals/s/core; finite both %d; max rel diff %.3g; inf mismatch %d"
          % (n_cpu, dt, n_cpu / dt, fin.sum(), np.max(np.abs(g[fin] - cpu[fin]) / np.abs(cpu[fin])) if fin.any() else -1,
             int((np.isfinite(cpu) != np.isfinite(g)).sum())))
    # NS steps
    ptab = [(pr.GL_PRIOR_UNIFORM, a, b) for a, b in table]
    k = max(1, N // 10)
    eng = DeviceEngine(ptab, like, N, k, (steps + 1) * k, iterations=int(os.environ.get('JARM_IT', '10')),
                       burn_in=int(os.environ.get('JARM_BURN', '10')), tuning_cycles=int(os.environ.get('JARM_TC', '1')), seed=5)
    torch.cuda.synchronize(); t0 = time.time()
    eng.init(); torch.cuda.synchronize()
    print("init N=%d: %.3f s" % (N, time.time() - t0), like.counters(reset=True))
    for s in range(steps):
        t0 = time.time()
        eng.advance((s + 2) * k); torch.cuda.synchronize()
        dt = time.time() - t0
        c = like.counters(reset=True)
        st = eng.state()
        print("step %d: %.3f s, evals %d (%.0f/s), solves/eval %.2f, steps/solve %.1f, steps/s %.3g, L* %.3f"
              % (s, dt, c["evals"], c["evals"] / dt, c["solves"] / max(1, c["evals"]), c["steps"] / max(1, c["solves"]),
                 c["steps"] / dt, st.lstar))


if __name__ == "__main__":
    main()
