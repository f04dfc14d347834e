"""Oracle pin (ODE): the SciPy stand-in for ScipyOdeSimulator against the one golden PySB
trajectory the reference ships (model_0_AB_complex_data.npy -> tests/golden/ref_ode.npz),
plus host-side checks of the network code generator (no GPU)."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import ode


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_ode.npz"))


def test_model_0_matches_reference_golden_trajectory(g):
    from gleipnir_b200 import networks
    d = networks.model_0().description()
    _, obs = ode.simulate(d, d["param_values"], g["model_0/tspan"])
    # ScipyOdeSimulator default tolerance is rtol 1e-6: the fixture itself is only that accurate
    np.testing.assert_allclose(obs[-10:, 2], g["model_0/AB_complex_last10"], rtol=1e-6)


def test_loglike_forms(g):
    from gleipnir_b200 import networks
    dd = networks.dimerization().description()
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    rows = [(1, data, 10.0 * sd, np.ones(51))]
    _, obs = ode.simulate(dd, np.array([1e-3, 1.0, 2000.0]), ts)
    sim = obs[:, 1]
    from scipy.stats import norm
    want = np.sum(norm(loc=data, scale=10.0 * sd).logpdf(sim))          # run_NS_dimerization_model.py:26,62
    assert np.isclose(ode.loglike(dd, [0, 1], [-3.0, 0.0], ts, rows), want, rtol=1e-12)
    assert np.isclose(ode.loglike(dd, [0, 1], [-3.0, 0.0], ts, rows, "mse"), -np.mean((data - sim) ** 2), rtol=1e-12)
    assert np.isclose(ode.loglike(dd, [0, 1], [-3.0, 0.0], ts, rows, "sse"), -np.sum((data - sim) ** 2), rtol=1e-12)


def test_generated_source_and_jacobian_consistency():
    from gleipnir_b200 import networks
    for mk in networks.SHIPPED.values():
        net = mk()
        src = net.cuda_struct()
        assert "static constexpr int NS = %d" % net.ns in src
        # the generated Jacobian is the derivative of the generated rhs: check through the oracle's
        # numeric forms built from the same description
        d = net.description()
        f, jac = ode.rhs_jac(d)
        y = np.linspace(1.0, 2.0, net.ns) * 100.0
        p = d["param_values"]
        J = jac(y, 0.0, p)
        for q in range(net.ns):
            e = np.zeros(net.ns)
            e[q] = 1e-3
            np.testing.assert_allclose((f(y + e, 0, p) - f(y - e, 0, p)) / 2e-3, J[:, q], rtol=1e-6, atol=1e-9)


def test_dimerization_uses_bionetgen_statistical_factor():
    from gleipnir_b200 import networks
    src = networks.dimerization().cuda_struct()
    assert "0.5 * p[0] * y[0] * y[0]" in src          # A + A -> AA: k/2! * A^2, dA/dt = -2 v
    assert "f[0] = - 2.0 * v0 + 2.0 * v1" in src


def test_plugin_builds_and_exports(tmp_path):
    from gleipnir_b200 import networks, odegen
    so = odegen.compile_plugin(networks.dimerization())
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    for sym in ("gl_plugin_walk_launch", "gl_plugin_eval_launch", "gl_plugin_last_error", "gl_plugin_info"):
        assert re.search(r" T %s\b" % sym, out), sym
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
