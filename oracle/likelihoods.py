"""TEST INFRASTRUCTURE ONLY (oracle) -- analytic log-likelihoods, batched numpy.

Restates the user likelihoods the reference ships with its tests/examples:
    iid Gaussian   tests/test_nested_sampling.py:34-36  (= examples/gaussians/*:21-23)
    egg carton     examples/eggcarton/NS_eggcarton.py:25-27
    1-D mixture    examples/1d_multimodal/NS_1d_multimodal.py:20-32
plus the builder-defined dense correlated Gaussian (not in
the reference).  Each takes x[n, d] and returns logL[n].
"""
import numpy as np

_HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


def iid_gaussian(x, mu=0.0, sigma=1.0):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    d = x.shape[1]
    z = (x - mu) / sigma
    acc = np.zeros(x.shape[0])
    for j in range(d):                     # fixed left-to-right order (device does the same)
        acc = acc + z[:, j] * z[:, j]
    return -0.5 * acc + d * (-_HALF_LOG_2PI - np.log(sigma))


def eggcarton(x):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    chi = np.ones(x.shape[0])
    for j in range(x.shape[1]):
        chi = chi * np.cos(x[:, j])
    t = 2.0 + chi
    t2 = t * t
    return t2 * t2 * t


def mixture1d(x, positions, width):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    positions = np.asarray(positions, dtype=np.float64)
    z = (x[:, :1] - positions[None, :]) / width
    l = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi * width * width)
    with np.errstate(divide="ignore"):
        out = np.log(l.mean(axis=1))
    return np.where(np.isnan(out), -np.inf, out)


def dense_gaussian(x, mu, P, log_norm):
    """-0.5 (x-mu)^T P (x-mu) + log_norm,  log_norm = -0.5*logdet(2*pi*Sigma)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    r = x - np.asarray(mu)[None, :]
    return -0.5 * np.einsum("ni,ij,nj->n", r, np.asarray(P), r) + log_norm


def make_c3_dense_gaussian(d=50, seed=1234, lam_lo=1e-2, lam_hi=1.0):
    """The synthetic C3 target of SURVEY 8d: Sigma = Q diag(lam) Q^T."""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.logspace(np.log10(lam_lo), np.log10(lam_hi), d)
    P = (Q / lam[None, :]) @ Q.T
    P = 0.5 * (P + P.T)
    log_norm = -0.5 * (d * np.log(2.0 * np.pi) + np.sum(np.log(lam)))
    return np.zeros(d), P, float(log_norm)


class Model(object):
    """Tiny tagged union so the oracle engine can dispatch like the device does."""

    def __init__(self, kind, d, **params):
        self.kind, self.d, self.params = kind, d, params

    def __call__(self, x):
        k, p = self.kind, self.params
        if k == "iid_gaussian":
            return iid_gaussian(x, p.get("mu", 0.0), p.get("sigma", 1.0))
        if k == "eggcarton":
            return eggcarton(x)
        if k == "mixture1d":
            return mixture1d(x, p["positions"], p["width"])
        if k == "dense_gaussian":
            return dense_gaussian(x, p["mu"], p["P"], p["log_norm"])
        if k == "callable":
            return np.array([p["fn"](row) for row in np.atleast_2d(x)], dtype=np.float64)
        raise ValueError(k)
