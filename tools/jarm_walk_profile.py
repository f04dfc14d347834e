"""One short refill of the C5 walkers for an ncu capture of walk_coop_kernel itself: 800 walkers (one resident
wave), ONE sweep each (19 evaluations per walker instead of 721, so that ncu's replay passes stay affordable).
python tools/jarm_walk_profile.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
from gleipnir_b200 import jarm                      # noqa: E402
from gleipnir_b200.engine import DeviceEngine       # noqa: E402
from oracle import priors as pr                      # noqa: E402

like = jarm.jarm_loglikelihood()
table = [(pr.GL_PRIOR_UNIFORM, a, b) for a, b in jarm.jarm_prior_table()]
lo = torch.tensor([a for _, a, b in table], dtype=torch.float64, device="cuda")
sc = torch.tensor([b for _, a, b in table], dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(3)
N, k = 8000, 800
xs, ls = [], []
while sum(len(x) for x in xs) < N:
    x = lo + sc * torch.rand((20000, 18), generator=g, dtype=torch.float64, device="cuda")
    l = like.evaluate(x)
    m = torch.isfinite(l)
    xs.append(x[m]); ls.append(l[m])
eng = DeviceEngine(table, like, N, k, 4 * k, iterations=1, burn_in=0, tuning_cycles=0, seed=1)
eng.init(torch.cat(xs)[:N].contiguous(), torch.cat(ls)[:N].contiguous())
like.counters(reset=True)
eng.kill(k)
eng.refill()
eng.kill(k)
torch.cuda.synchronize()
print(like.counters(), eng.state().lstar)
