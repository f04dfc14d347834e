"""TEST INFRASTRUCTURE ONLY (oracle) -- one rank's shard of a sharded run, in numpy.

Implements the shard-ops protocol of ``gleipnir_b200/sharded.py`` on top of the oracle engine
so that (a) world_size-2 gloo tests can exercise the product's SPMD loop on CPUs and (b) the
CUDA shards can be checked value-for-value.  Semantics = ``csrc/engine.cu`` sharded entry
points: candidates are the rank's k worst in ascending slot order; the global k worst are
selected from the rank-major gathered array by (value, position) -- position order is
global-index order; every rank accumulates identically; walkers are seeded from the rank's own
survivors.
"""
import numpy as np

from .engine import Engine, select_k_worst


class OracleShard(Engine):

    def candidates(self, k):
        self._cand = np.sort(select_k_worst(self.logl, k))
        return self.logl[self._cand].copy()

    def commit(self, gathered, world, rank, k):
        gathered = np.asarray(gathered, dtype=np.float64)
        m = len(gathered) // world                          # candidates per rank (m <= k: truncated exchange)
        order = select_k_worst(gathered, k)                 # death order over the whole world
        pos = np.sort(order)
        if m < k:                                           # exactness proof: no rank had all m selected
            per_rank = np.bincount(pos // m, minlength=world)
            if (per_rank >= m).any():
                self.shard_overflow = True
        mine = pos[(pos >= rank * m) & (pos < (rank + 1) * m)] - rank * m
        local = self._cand[mine]
        logw = self.logX - np.log(self.n_global + 1.0)
        w = np.exp(logw)
        for i in local:                                      # rows this rank records (slot order)
            self.dead_rows.append(np.concatenate(([self.logl[i], w], self.live[i])))
            self.dead_index.append(int(i) + self.row_offset)
            self.dead_batch.append(self.step)
        self.acc.add_terms(logw, gathered[pos])              # identical on every rank
        self.logX += np.log((self.n_global - k + 1.0) / (self.n_global + 1.0))
        self.n_dead += k
        self.last_dead = local
        self.lstar = gathered[order[-1]]

    def overflowed(self):
        return bool(getattr(self, "shard_overflow", False))

    def init(self, *a, **kw):
        # a full restart, like gl_ns_init on the device (a run repeated after an overflow)
        from .engine import Accumulators
        self.shard_overflow = False
        self.acc = Accumulators()
        self.logX, self.n_dead, self.step, self.n_evals = 0.0, 0, 0, 0
        self.dead_rows, self.dead_index, self.dead_batch = [], [], []
        self.last_dead = np.zeros(0, dtype=np.int64)
        self.lstar = -np.inf
        return super().init(*a, **kw)

    def refill(self, injected=None):
        if len(self.last_dead) == 0:
            self.step += 1
            return
        self.walk(injected)

    def final_partial(self):
        dead = np.zeros(self.N, dtype=bool)
        dead[self.last_dead] = True
        surv = self.logl[~dead]
        fin = np.isfinite(surv)
        with np.errstate(all="ignore"):
            if fin.any():
                m = float(np.max(surv[fin]))
                s = float(np.sum(np.exp(surv[fin] - m)))
            else:
                m, s = -np.inf, 0.0
            return np.array([m, s, float(np.sum(np.exp(surv))), float(len(surv))])

    def finalize_sharded(self, gathered, world, rank):
        g = np.asarray(gathered, dtype=np.float64).reshape(world, 4)
        with np.errstate(all="ignore"):
            cnt = g[:, 3].sum()
            lin = g[:, 2].sum()
            X = np.exp(self.logX)
            lm = lin / cnt
            self.acc.Z += X * lm
            dH = X * lm * np.log(lm)
            if not np.isnan(dH):
                self.acc.H += dH
            ok = np.isfinite(g[:, 0])
            if ok.any():
                mx = g[ok, 0].max()
                loglm = mx + np.log(np.sum(g[ok, 1] * np.exp(g[ok, 0] - mx))) - np.log(cnt)
                lw = self.logX + loglm
                m_new = max(self.acc.M, lw)
                scale = np.exp(self.acc.M - m_new) if np.isfinite(self.acc.M) else 0.0
                t = np.exp(lw - m_new)
                self.acc.Zs = self.acc.Zs * scale + t
                self.acc.Hs = self.acc.Hs * scale + t * loglm
                self.acc.M = m_new
        dead = np.zeros(self.N, dtype=bool)
        dead[self.last_dead] = True
        aw = np.exp(self.logX) / cnt
        for i in np.nonzero(~dead)[0]:
            self.dead_rows.append(np.concatenate(([self.logl[i], aw], self.live[i])))
        self.result = dict(logZ=self.acc.logZ, information=self.acc.information, dead=np.array(self.dead_rows),
                           dead_index=np.array(self.dead_index), dead_batch=np.array(self.dead_batch),
                           n_dead=self.n_dead, n_evals=self.n_evals, widths=self.widths.copy())
        return self.result
