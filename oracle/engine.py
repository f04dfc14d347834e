"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the batched NS engine.

This is the checker for the CUDA engine (``gleipnir_b200/csrc``): same Philox
counters, same arithmetic forms, numpy instead of kernels.  It restates

  outer loop     nested_sampling.py:83-213   generalised from 1 to k deaths per step
  sampler        samplers.py:85-114          the *correct* tuning-loop move applied to
                                             tuning AND production (quirks
                                             Q1/Q2 are oracle facts of reference_port.py,
                                             not behaviour to ship)
  priors         sampled_parameter.py:40-80  closed forms (oracle/priors.py)
  stop tests     stopping_criterion.py:32,57 (turned into a death count by the caller)

Batch semantics (k deaths per step; k=1 is the reference's sequencing):
  * the k worst by (logL, index) -- NaN first, lowest index on ties, i.e. np.argmin
    order (nested_sampling.py:115,156) -- die together; L* = the largest of them;
  * live count while removing the j-th of a batch is N-j, shrink factor (N-j)/(N-j+1)
    (mean-of-Beta convention of ``_alpha = N/(N+1)``, nested_sampling.py:64); the
    product telescopes, so every death of a batch has weight X/(N+1) and afterwards
    X <- X*(N-k+1)/(N+1).  k=1: w_n = alpha^(n-1)/(N+1) = alpha^(n-1) - alpha^n (:143);
  * each dead slot is refilled by a walker started from a uniformly chosen survivor
    (int(U*N), redraw while dead -- :147-149) and moved under the hard wall logL > L*;
  * proposal widths: 0.5*range of 100 prior draws (samplers.py:70-81), then per
    tuning cycle of 20 sweeps x0.66 / x1.33 when the acceptance rate pooled over the
    step's walkers is <0.2 / >0.6 (docstring semantics, samplers.py:26-30), persisted;
  * evidence/information are kept both in the reference's linear form (:118-125) and
    as scaled log-sum-exp accumulators (superset: no under/overflow).
"""
import numpy as np

from . import philox as px
from . import priors as pr


def sort_key(logl):
    """float64 -> uint64 key, ascending; NaN -> 0 (first), -0.0 == +0.0."""
    x = np.asarray(logl, dtype=np.float64) + 0.0
    b = x.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    k = np.where(neg, ~b, b | np.uint64(0x8000000000000000))
    return np.where(np.isnan(x), np.uint64(0), k)


def select_k_worst(logl, k):
    """Indices of the k smallest by (key, index), in death order."""
    key = sort_key(logl)
    order = np.lexsort((np.arange(len(key)), key))
    return order[:k]


class Accumulators(object):
    """Z, H in linear space (reference form) and as (M, Zs, Hs) with Z = Zs*e^M."""

    def __init__(self):
        self.Z = 0.0
        self.H = 0.0
        self.M = -np.inf
        self.Zs = 0.0
        self.Hs = 0.0

    def add_terms(self, logw, logl):
        """Add sum_i exp(logw)*exp(logl_i) (all deaths of a batch share logw)."""
        logl = np.asarray(logl, dtype=np.float64)
        with np.errstate(all="ignore"):
            w = np.exp(logw)
            for l in logl:                                   # linear form, death order
                dZ = w * np.exp(l)
                self.Z += dZ
                dH = dZ * l
                if not np.isnan(dH):
                    self.H += dH
            lw = logw + logl
            fin = np.isfinite(lw)
            if fin.any():
                m_new = max(self.M, float(np.max(lw[fin])))
                scale = np.exp(self.M - m_new) if np.isfinite(self.M) else 0.0
                t = np.exp(lw[fin] - m_new)
                self.Zs = self.Zs * scale + float(np.sum(t))
                self.Hs = self.Hs * scale + float(np.sum(t * logl[fin]))
                self.M = m_new

    def add_final(self, logX, logl_surv):
        """nested_sampling.py:184-194: Z += X*mean(exp(L)); H += X*lm*log(lm)."""
        with np.errstate(all="ignore"):
            X = np.exp(logX)
            lm = np.exp(logl_surv).mean()
            self.Z += X * lm
            dH = X * lm * np.log(lm)
            if not np.isnan(dH):
                self.H += dH
            fin = np.isfinite(logl_surv)
            if fin.any():
                mx = float(np.max(logl_surv[fin]))
                loglm = mx + np.log(np.sum(np.exp(logl_surv[fin] - mx))) - np.log(len(logl_surv))
                lw = logX + loglm
                m_new = max(self.M, lw)
                scale = np.exp(self.M - m_new) if np.isfinite(self.M) else 0.0
                t = np.exp(lw - m_new)
                self.Zs = self.Zs * scale + t
                self.Hs = self.Hs * scale + t * loglm
                self.M = m_new

    @property
    def logZ(self):
        with np.errstate(all="ignore"):
            return self.M + np.log(self.Zs) if self.Zs > 0.0 else -np.inf

    @property
    def information(self):
        return (-self.logZ + self.Hs / self.Zs) if self.Zs > 0.0 else 0.0

    @property
    def information_linear(self):
        with np.errstate(all="ignore"):
            return (-np.log(self.Z) + self.H / self.Z) if self.Z > 0.0 else 0.0


class Engine(object):
    def __init__(self, prior_table, model, population_size, batch_k=1, iterations=100, burn_in=0,
                 tuning_cycles=0, proposal="uniform", seed=0, row_offset=0, global_population=None):
        self.pt = [tuple(p) for p in prior_table]
        self.row_offset = int(row_offset)
        self.n_global = int(global_population or population_size)
        self.model = model
        self.N, self.d = int(population_size), len(prior_table)
        self.k = int(batch_k)
        self.iterations, self.burn_in, self.tc = iterations, burn_in, tuning_cycles
        self.proposal = proposal
        self.seed = int(seed)
        self.acc = Accumulators()
        self.logX = 0.0
        self.n_dead = 0
        self.step = 0
        self.n_evals = 0
        self.dead_rows = []          # death order
        self.dead_index = []
        self.dead_batch = []
        self.last_dead = np.zeros(0, dtype=np.int64)
        self.lstar = -np.inf

    # -- a2/a6: initial population --------------------------------------------------------
    def draw_rows(self, row0, n):
        g = np.arange(row0, row0 + n, dtype=np.uint64)
        x = np.empty((n, self.d))
        for j, (kind, p0, p1) in enumerate(self.pt):
            u, _ = px.uniforms(self.seed, g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), j, px.GL_TAG_INIT)
            x[:, j] = pr.ppf(kind, p0, p1, u)
        return x

    def init_widths(self):
        w = np.empty(self.d)
        i = np.arange(100, dtype=np.uint64)
        for j, (kind, p0, p1) in enumerate(self.pt):
            u, _ = px.uniforms(self.seed, i, 0, j, px.GL_TAG_WIDTH)
            rs = pr.ppf(kind, p0, p1, u)
            w[j] = 0.5 * (rs.max() - rs.min())
        return w

    def init(self, live=None, logl=None):
        self.live = self.draw_rows(self.row_offset, self.N) if live is None else np.array(live, dtype=np.float64)
        if logl is None:
            self.logl = self.model(self.live)
            self.n_evals += self.N
        else:
            self.logl = np.array(logl, dtype=np.float64)
        self.widths = self.init_widths()

    # -- a8/a13/a14/a15: select + record one batch ---------------------------------------
    def kill(self, k):
        idx = select_k_worst(self.logl, k)
        logw = self.logX - np.log(self.n_global + 1.0)
        w = np.exp(logw)
        for i in idx:
            self.dead_rows.append(np.concatenate(([self.logl[i], w], self.live[i])))
            self.dead_index.append(int(i))
            self.dead_batch.append(self.step)
        self.acc.add_terms(logw, self.logl[idx])
        self.logX += np.log((self.n_global - k + 1.0) / (self.n_global + 1.0))
        self.n_dead += k
        self.last_dead = idx
        self.lstar = self.logl[idx[-1]]
        return idx

    # -- a11/a12/a16-a18: refill the dead slots ------------------------------------------
    def _propose(self, r, cur, w):
        if self.proposal == "normal":
            return pr.ndtri(r) * w + cur              # samplers.py:100
        return r * w + (cur + (-0.5 * w))             # samplers.py:102: ppf(r, loc=cur-w/2, scale=w)

    def _sweeps(self, slots, cur, cur_l, widths, t0, n_sweeps, count):
        """n_sweeps component-wise sweeps for all walkers at once (vectorised over walkers)."""
        d = self.d
        acc = np.zeros(d)
        with np.errstate(all="ignore"):
            for t in range(t0, t0 + n_sweeps):
                for j, (kind, p0, p1) in enumerate(self.pt):
                    r, u = px.uniforms(self.seed, slots + self.row_offset, self.step, t * d + j, px.GL_TAG_MOVE)
                    xj = cur[:, j]
                    nj = self._propose(r, xj, widths[j])
                    okp = pr.accept_prior(kind, p0, p1, xj, nj, u)
                    trial = cur.copy()
                    trial[:, j] = nj
                    tl = self.model(trial)
                    ok = okp & (tl > self.lstar)
                    cur[ok, j] = nj[ok]
                    cur_l[ok] = tl[ok]
                    if count:
                        acc[j] += ok.sum()
        self.n_evals += n_sweeps * d * len(slots)
        return acc

    def pick_survivors(self, slots):
        dead = np.zeros(self.N, dtype=bool)
        dead[slots] = True
        picks = np.empty(len(slots), dtype=np.int64)
        for q, g in enumerate(slots):
            a = 0
            while True:
                u, _ = px.uniforms(self.seed, int(g) + self.row_offset, self.step, a, px.GL_TAG_PICK)
                i = int(float(u) * self.N)
                if not dead[i]:
                    break
                a += 1
            picks[q] = i
        return picks

    def walk(self, injected=None):
        """Refill ``self.last_dead``; ``injected`` = (points, logls) bypasses the sampler."""
        slots = np.asarray(self.last_dead, dtype=np.int64)
        if injected is not None:
            pts, lls = injected
            self.live[slots] = pts
            self.logl[slots] = lls
            self.step += 1
            return
        picks = self.pick_survivors(slots)
        start = self.live[picks].copy()
        start_l = self.logl[picks].copy()
        widths = self.widths.copy()
        cur, cur_l = start.copy(), start_l.copy()
        for c in range(self.tc):
            acc = self._sweeps(slots, cur, cur_l, widths, c * 20, 20, True)
            with np.errstate(all="ignore"):
                rate = acc / (20.0 * len(slots))
            widths = np.where(rate < 0.2, widths * 0.66, np.where(rate > 0.6, widths * 1.33, widths))
        self.widths = widths
        cur, cur_l = start.copy(), start_l.copy()
        self._sweeps(slots, cur, cur_l, widths, self.tc * 20, self.iterations + self.burn_in, False)
        self.live[slots] = cur
        self.logl[slots] = cur_l
        self.step += 1

    # -- a15: final term -------------------------------------------------------------------
    def finalize(self):
        dead = np.zeros(self.N, dtype=bool)
        dead[self.last_dead] = True
        surv = np.nonzero(~dead)[0]
        self.acc.add_final(self.logX, self.logl[surv])
        aw = np.exp(self.logX) / len(surv)
        for i in surv:
            self.dead_rows.append(np.concatenate(([self.logl[i], aw], self.live[i])))
        return dict(logZ=self.acc.logZ, information=self.acc.information,
                    logZ_err=np.sqrt(max(self.acc.information, 0.0) / self.N),
                    Z_linear=self.acc.Z, H_linear=self.acc.H, information_linear=self.acc.information_linear,
                    dead=np.array(self.dead_rows), dead_index=np.array(self.dead_index),
                    dead_batch=np.array(self.dead_batch), n_dead=self.n_dead, n_evals=self.n_evals,
                    widths=self.widths.copy())

    def run(self, n_deaths, injected_stream=None):
        """Full run to exactly ``n_deaths`` deaths (the closed-form stop count)."""
        self.init()
        return self.continue_run(n_deaths, injected_stream)

    def continue_run(self, n_deaths, injected_stream=None):
        pos = 0
        while True:
            k = min(self.k, n_deaths - self.n_dead)
            self.kill(k)
            if self.n_dead >= n_deaths:
                break
            if injected_stream is not None:
                pts, lls = injected_stream
                self.walk((pts[pos:pos + k], lls[pos:pos + k]))
                pos += k
            else:
                self.walk()
        return self.finalize()
