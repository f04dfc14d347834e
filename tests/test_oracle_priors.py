"""Oracle pin: closed-form priors against scipy.stats values recorded from the reference's
SampledParameter (tests/golden/ref_priors.npz, made by oracle/gen_golden.py) and against
the known answers in the reference's tests/test_sampled_parameter.py:23-36."""
import os

import numpy as np
import pytest
from scipy.stats import loguniform, norm, uniform

from oracle import priors as pr

FAM = {"uniform": (pr.GL_PRIOR_UNIFORM, -5.0, 10.0), "norm": (pr.GL_PRIOR_NORM, 0.5, 2.0),
       "loguniform": (pr.GL_PRIOR_LOGUNIFORM, 1e-3, 1e2)}


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_priors.npz"))


@pytest.mark.parametrize("name", sorted(FAM))
def test_pdf_logpdf_ppf_match_reference(g, name):
    kind, p0, p1 = FAM[name]
    x, u = g[name + "/x"], g[name + "/u"]
    np.testing.assert_allclose(pr.pdf(kind, p0, p1, x), g[name + "/prior"], rtol=1e-13, atol=0)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(pr.logpdf(kind, p0, p1, x), g[name + "/logprior"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(pr.ppf(kind, p0, p1, u), g[name + "/invcdf"], rtol=1e-13, atol=1e-15)


def test_reference_known_answers(g):
    # tests/test_sampled_parameter.py:23-36 (norm(0,1) at 0.5)
    assert np.isclose(pr.logpdf(pr.GL_PRIOR_NORM, 0.0, 1.0, 0.5), -1.0439385332046727, rtol=1e-15)
    assert np.isclose(pr.pdf(pr.GL_PRIOR_NORM, 0.0, 1.0, 0.5), 0.3520653267642995, rtol=1e-15)
    assert abs(pr.ppf(pr.GL_PRIOR_NORM, 0.0, 1.0, 0.5)) < 1e-16
    assert g["known/logprior_0.5"] == pytest.approx(-1.0439385332046727, rel=1e-15)
    assert g["known/prior_0.5"] == pytest.approx(0.3520653267642995, rel=1e-15)


def test_parse_scipy_frozen():
    assert pr.parse_scipy_frozen(uniform(loc=-5.0, scale=10.0)) == (pr.GL_PRIOR_UNIFORM, -5.0, 10.0)
    assert pr.parse_scipy_frozen(norm(0.5, 2.0)) == (pr.GL_PRIOR_NORM, 0.5, 2.0)
    assert pr.parse_scipy_frozen(loguniform(1e-3, 1e2)) == (pr.GL_PRIOR_LOGUNIFORM, 1e-3, 1e2)
