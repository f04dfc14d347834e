"""Where a C5 walker launch's time goes: per-block (SM id, start, end, ODE steps) trace of one NS batch.
python tools/walker_trace.py [N] [k] [steps]   (live set = finite-likelihood prior draws, as in bench.py)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
from gleipnir_b200 import _lib as L                 # noqa: E402
from gleipnir_b200 import jarm                      # noqa: E402
from gleipnir_b200.engine import DeviceEngine       # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    like = jarm.jarm_loglikelihood()
    table = [(L.GL_PRIOR_UNIFORM, a, b) for a, b in jarm.jarm_prior_table()]
    lo = torch.tensor([a for _, a, b in table], dtype=torch.float64, device="cuda")
    sc = torch.tensor([b for _, a, b in table], dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    xs, ls, have = [], [], 0
    while have < N:
        x = lo + sc * torch.rand((max(4096, N), len(table)), generator=g, dtype=torch.float64, device="cuda")
        l = like.evaluate(x)
        m = torch.isfinite(l)
        xs.append(x[m]); ls.append(l[m]); have += int(m.sum())
    x0, l0 = torch.cat(xs)[:N].contiguous(), torch.cat(ls)[:N].contiguous()
    eng = DeviceEngine(table, like, N, k, (steps + 3) * k, iterations=10, burn_in=10, tuning_cycles=1, proposal=0, seed=2024)
    eng.init(x0, l0)
    eng.kill(k)
    eng.refill(); eng.kill(k)                         # one warm step
    torch.cuda.synchronize()
    for s in range(steps):
        like.counters(reset=True)
        like.trace(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.refill(); eng.kill(k)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        c = like.counters(reset=True)
        tr = like.trace()
        like.trace(False)
        print("step %d: %.3f s, %d evals (%.0f/s), %.3g ODE steps/s, steps/solve %.1f, %d trace records"
              % (s, ms * 1e-3, c["evals"], c["evals"] / ms * 1e3, c["steps"] / ms * 1e3, c["steps"] / max(c["solves"], 1), len(tr)))
        for phase in sorted(set(tr[:, 0] >> 32)):
            t = tr[(tr[:, 0] >> 32) == phase]
            t = t[t[:, 3] > 0]                         # blocks that had walkers
            sm = t[:, 0] & 0xffffffff
            t0 = t[:, 1].min()
            wall = (t[:, 2].max() - t0) * 1e-9
            dur = (t[:, 2] - t[:, 1]) * 1e-9
            print("  phase %d: %d blocks on %d SMs, launch %.3f s; block duration min/median/max %.3f/%.3f/%.3f s; "
                  "block-time utilisation %.3f" % (phase, len(t), len(set(sm.tolist())), wall, dur.min(), np.median(dur),
                                                   dur.max(), dur.sum() / (len(t) * wall)))
            fin, work, nb = [], [], []
            for s_ in sorted(set(sm.tolist())):
                q = t[sm == s_]
                fin.append((q[:, 2].max() - t0) * 1e-9); work.append(q[:, 3].sum()); nb.append(len(q))
            fin, work = np.array(fin), np.array(work, dtype=float)
            print("    per SM: blocks %d..%d; finish min/median/max %.3f/%.3f/%.3f s; ODE steps min/median/max %.3g/%.3g/%.3g "
                  "(CV %.3f); corr(work, finish) %.3f; mean(finish)/max(finish) %.3f; mean(work)/max(work) %.3f"
                  % (min(nb), max(nb), fin.min(), np.median(fin), fin.max(), work.min(), np.median(work), work.max(),
                     work.std() / work.mean(), np.corrcoef(work, fin)[0, 1], fin.mean() / fin.max(), work.mean() / work.max()))
            bw = t[:, 3].astype(float)
            print("    per block: ODE steps CV %.3f, max/mean %.3f; rate while running (steps/s per block) min/median/max %.3g/%.3g/%.3g"
                  % (bw.std() / bw.mean(), bw.max() / bw.mean(), (bw / dur).min(), np.median(bw / dur), (bw / dur).max()))


if __name__ == "__main__":
    main()
