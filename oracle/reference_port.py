"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference AS SHIPPED.

A from-scratch, array-based restatement of what
``gleipnir/nestedsampling/nested_sampling.py:83-213`` (outer loop),
``gleipnir/nestedsampling/samplers.py:56-152`` (sampler, *including* its
behaviour-defining quirks Q1/Q2) and
``gleipnir/sampled_parameter.py:40-80`` do, consuming numpy's global legacy
RNG and scipy.stats frozen distributions in exactly the reference's order, so
that for a given ``np.random.seed`` it reproduces the reference's outputs
bit-for-bit.  Pinned by ``tests/golden/ref_runs.npz`` (recorded from the real
reference by ``oracle/gen_golden.py``) in ``tests/test_oracle_reference_port.py``.

Uses: (1) the `cpu_baseline` / ``--impl reference`` leg of ``bench.py`` (the
reference itself cannot travel to the GPU box); (2) checker for the outer-loop
bookkeeping replay tests.  Never imported by the product path.
"""
import math

import numpy as np
from scipy.stats import norm as _sp_norm
from scipy.stats import uniform as _sp_uniform


class ShippedSamplerState(object):
    """Mutable sampler state that persists across calls (samplers.py:48-53,124)."""

    def __init__(self, iterations=100, burn_in=0, tuning_cycles=0, proposal="uniform"):
        self.iterations, self.burn_in = iterations, burn_in
        self.tuning_cycles, self.proposal = tuning_cycles, proposal
        self.widths = None

    def _propose(self, r, center, width):
        if self.proposal == "normal":
            return _sp_norm.ppf(r, loc=center, scale=width)            # samplers.py:100,136
        return _sp_uniform.ppf(r, loc=center - (width / 2.0), scale=width)   # :102,138

    def __call__(self, priors, loglikelihood, start, boundary, counter=None):
        """One sampler call.  ``priors`` = list of scipy frozen dists (norm == 1 assumed
        is NOT assumed: the density is divided by cdf(inf) as sampled_parameter.py:70)."""
        d = len(priors)
        norms = [p.cdf(np.inf) for p in priors] if not hasattr(self, "_norms") else self._norms
        self._norms = norms
        if self.widths is None:                                       # samplers.py:70-81
            w = []
            for p in priors:
                rs = p.rvs(100)
                w.append(0.5 * (np.max(rs) - np.min(rs)))
            self.widths = np.array(w)
        nevals = 1
        cur_l = loglikelihood(start)                                  # :83
        steps = self.widths.copy()
        acc = np.zeros(d)
        cur = start.copy()
        for _cycle in range(self.tuning_cycles):                      # :90-121
            for _k in range(20):
                rs = np.random.random(d)
                us = np.random.random(d)
                for j in range(d):
                    trial = cur.copy()
                    xj = cur[j]
                    nj = self._propose(rs[j], xj, self.widths[j])     # widths, not steps
                    trial[j] = nj
                    ratio = (priors[j].pdf(nj) / norms[j]) / (priors[j].pdf(xj) / norms[j])
                    tl = loglikelihood(trial)
                    nevals += 1
                    if (us[j] < ratio) and (tl > boundary):
                        cur[j] = nj
                        cur_l = tl
                        acc[j] += 1.0
                # Q2: the adaptation block sits inside the k loop (samplers.py:115-121)
                rate = acc / 20.0
                steps[rate < 0.2] *= 0.66
                steps[rate > 0.6] *= 1.33
                acc[:] = 0.0
        self.widths = steps.copy()                                    # :124 (persists)
        cur = start.copy()                                            # :125 (cur_l NOT reset)
        for _i in range(self.iterations + self.burn_in):              # :127-149
            rs = np.random.random(d)
            us = np.random.random(d)
            for j in range(d):
                trial = cur.copy()
                xj = cur[j]
                nj = self._propose(rs[j], xj, self.widths[j])
                # Q1: trial[j] is never set -> prior and likelihood see the current point
                ratio = (priors[j].pdf(trial[j]) / norms[j]) / (priors[j].pdf(xj) / norms[j])
                tl = loglikelihood(trial)
                nevals += 1
                if (us[j] < ratio) and (tl > boundary):
                    cur[j] = nj
                    cur_l = tl
        if counter is not None:
            counter[0] += nevals
        return cur, cur_l


def run_shipped(priors, loglikelihood, population_size, sampler, n_iterations=None,
                prior_mass_cutoff=None, replacement_stream=None, max_iterations=None):
    """The outer loop as shipped.  Exactly one of ``n_iterations`` (NumberOfIterations,
    stopping_criterion.py:32) / ``prior_mass_cutoff`` (RemainingPriorMass, :57).

    ``replacement_stream``: optional iterable of (point, logl); when given it stands in
    for the sampler call at nested_sampling.py:152 (deterministic replay tests).

    Returns a dict: logZ, logZ_err, evidence, evidence_error, information, H,
    dead (ndarray [rows, d+2]: log_l, weight, x...), dead_index (slot of each of the
    first n rows), n_iterations, n_evals.
    """
    N, d = population_size, len(priors)
    alpha = N / (N + 1)                                               # :64
    live = np.empty((N, d))
    for i in range(N):                                                # :94-101 (row-major RNG order)
        for j in range(d):
            live[i, j] = priors[j].rvs(1)[0]
    logl = np.array([loglikelihood(row) for row in live])             # :108
    counter = [N]
    stream = iter(replacement_stream) if replacement_stream is not None else None

    def stop(n):
        if max_iterations is not None and n >= max_iterations:
            return True
        if n_iterations is not None:
            return n >= n_iterations
        return alpha ** n <= prior_mass_cutoff

    Z = 0.0
    H = 0.0
    info = 0.0
    dead_rows, dead_index = [], []
    n = 1
    w = 1.0 - alpha ** n                                              # :112
    ndx = int(np.argmin(logl))
    while True:
        ll = logl[ndx]
        dZ = w * np.exp(ll)                                           # :118,160
        Z += dZ
        dH = dZ * ll
        if np.isnan(dH):
            dH = 0.0
        H += dH
        if Z > 0.0:
            info = -np.log(Z) + H / Z                                 # :125,167
        dead_rows.append(np.concatenate(([ll, w], live[ndx])))
        dead_index.append(ndx)
        if stop(n):
            break
        n += 1
        w = alpha ** (n - 1.0) - alpha ** n                           # :143
        if stream is None:
            pick = int(np.random.random(1)[0] * N)                    # :147-149
            while pick == ndx:
                pick = int(np.random.random(1)[0] * N)
            new_x, new_l = sampler(priors, loglikelihood, live[pick], ll, counter)
        else:
            new_x, new_l = next(stream)
        logl[ndx] = new_l                                             # :153-154
        live[ndx] = new_x
        ndx = int(np.argmin(logl))                                    # :156

    wf = alpha ** n                                                   # :184
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        lk = np.exp(logl)
        surv = np.array([v for i, v in enumerate(lk) if i != ndx])
        lm = surv.mean()
        Z += wf * lm
        dH = wf * lm * np.log(lm)
        if np.isnan(dH):
            dH = 0.0
        H += dH
        if Z > 0.0:
            info = -np.log(Z) + H / Z
        aw = wf / len(surv)
        for i in range(N):
            if i != ndx:
                dead_rows.append(np.concatenate(([logl[i], aw], live[i])))
        logz_err = np.sqrt(info / N)                                  # :205
        out = dict(logZ=np.log(Z), logZ_err=logz_err, evidence=Z, evidence_error=np.exp(logz_err),
                   information=info, H=H, dead=np.array(dead_rows), dead_index=np.array(dead_index),
                   n_iterations=n, n_evals=counter[0], final_live=live.copy(), final_logl=logl.copy())
    return out


def posteriors(dead, evidence, nbins=None):
    """nested_sampling.py:266-293 on a dead table [rows, d+2] -> list of (marginal, edges, centers)."""
    logl, w = dead[:, 0], dead[:, 1]
    nw = (w * np.exp(logl)) / evidence
    mask = nw > 0.0
    if nbins is None:
        nbins = 2 * int(np.cbrt(len(nw[mask])))                       # Rice rule
    out = []
    for j in range(2, dead.shape[1]):
        marginal, edge = np.histogram(dead[mask, j], weights=nw[mask], density=True, bins=nbins)
        out.append((marginal, edge, (edge[:-1] + edge[1:]) / 2.0))
    return out


def n_stop_for_prior_mass(alpha, cutoff):
    """Smallest n >= 1 with alpha**n <= cutoff, evaluated with the literal predicate."""
    n = max(1, int(math.ceil(math.log(cutoff) / math.log(alpha))))
    while n > 1 and alpha ** (n - 1) <= cutoff:
        n -= 1
    while alpha ** n > cutoff:
        n += 1
    return n
