"""Oracle pin: the as-shipped restatement (oracle/reference_port.py) must reproduce the
REAL reference bit-for-bit on the seeded runs recorded in tests/golden/ref_runs.npz
(incl. the SURVEY 6.2 anchor: seed 0, tests/test_nested_sampling.py config ->
(-9.852104048247114, 0.30841489785853227), information 1.902394984421778, 17 989 evals)."""
import os
import warnings

import numpy as np
import pytest
from scipy.stats import norm, uniform

from oracle import reference_port as rp


def _gauss(d):
    const = -0.5 * np.log(2 * np.pi)
    return lambda x: -0.5 * np.sum(x ** 2) + d * const


def _egg(x):
    return (2.0 + (np.cos(x)).prod()) ** 5


CASES = {
    "anchor_gauss5_N20": (0, 5, lambda: uniform(loc=-5.0, scale=10.0), _gauss(5), 20,
                          dict(iterations=10, tuning_cycles=1), dict(n_iterations=120)),
    "gauss2_N50_tc0": (3, 2, lambda: uniform(loc=-5.0, scale=10.0), _gauss(2), 50,
                       dict(iterations=5, tuning_cycles=0), dict(n_iterations=200)),
    "egg_N40_rpm": (7, 2, lambda: uniform(loc=0.0, scale=10.0 * np.pi), _egg, 40,
                    dict(iterations=8, burn_in=2, tuning_cycles=2), dict(prior_mass_cutoff=1e-2)),
    "gauss3_normprior_normprop": (11, 3, lambda: norm(0.5, 2.0), _gauss(3), 30,
                                  dict(iterations=6, tuning_cycles=1, proposal="normal"), dict(n_iterations=90)),
}


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_runs.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_is_bit_exact(g, name):
    seed, d, mk, ll, N, skw, stop = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(seed)
        pri = [mk() for _ in range(d)]
        r = rp.run_shipped(pri, ll, N, rp.ShippedSamplerState(**skw), **stop)
    assert r["logZ"] == g[name + "/logZ"]
    assert r["logZ_err"] == g[name + "/logZ_err"]
    assert r["evidence"] == g[name + "/evidence"]
    assert r["information"] == g[name + "/information"]
    assert np.array_equal(r["dead"], g[name + "/dead"], equal_nan=True)
    post = rp.posteriors(r["dead"], r["evidence"])
    assert np.array_equal(post[0][0], g[name + "/post_marginal0"])
    assert np.array_equal(post[0][1], g[name + "/post_edges0"])


def test_anchor_values(g):
    assert g["anchor_gauss5_N20/logZ"] == -9.852104048247114
    assert g["anchor_gauss5_N20/logZ_err"] == 0.30841489785853227
    assert g["anchor_gauss5_N20/information"] == 1.902394984421778
    assert g["anchor_gauss5_N20/dead"].shape == (139, 7)


def test_eval_count_formula():
    # N + (n-1)*(1 + 20*tc*d + (it+burn)*d)  
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(0)
        pri = [uniform(loc=-5.0, scale=10.0) for _ in range(5)]
        r = rp.run_shipped(pri, _gauss(5), 20, rp.ShippedSamplerState(iterations=10, tuning_cycles=1),
                           n_iterations=120)
    assert r["n_evals"] == 20 + 119 * (1 + 20 * 1 * 5 + 10 * 5) == 17989


def test_prior_mass_stop_count_matches_literal_loop():
    for N, cutoff in [(40, 1e-2), (500, 1e-6), (10000, 1e-6), (7, 0.5)]:
        alpha = N / (N + 1)
        n = 1
        while not (alpha ** n <= cutoff):
            n += 1
        assert rp.n_stop_for_prior_mass(alpha, cutoff) == n
