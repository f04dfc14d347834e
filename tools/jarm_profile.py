"""One batch evaluation of the C5 likelihood for an ncu capture.  python tools/jarm_profile.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
from gleipnir_b200 import jarm                      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
like = jarm.jarm_loglikelihood(rtol=float(os.environ.get("JARM_RTOL", "1e-8")), max_steps=int(os.environ.get("JARM_MAXSTEPS", "100000")))
table = jarm.jarm_prior_table()
rng = np.random.default_rng(1)
lo = np.array([t[0] for t in table]); sc = np.array([t[1] for t in table])
X = torch.as_tensor(lo + sc * rng.uniform(size=(n, 18)), device="cuda")
out = like.evaluate(X)
torch.cuda.synchronize()
print(like.counters(), float(torch.isfinite(out).float().mean()))
