"""GPU parity (ODE likelihoods): the code-generated Rosenbrock kernels against the SciPy
stand-in oracle (LSODA, rtol=atol=1e-12) and the reference's golden PySB trajectory.
north_star tolerance: ODE trajectories within rtol 1e-6."""
import os

import numpy as np
import pytest

from oracle import ode
from oracle import priors as pr

pytestmark = pytest.mark.gpu

ODE_RTOL = 1e-6


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_ode.npz"))


def _simulate_on_device(network, sampled, Y, tspan, obs):
    """Observable trajectory read back through the likelihood itself: with data = 0, kind 'sse'
    and the mask on one time point, logL = -sim(t_i)^2."""
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    out = np.zeros(len(tspan))
    for i in range(len(tspan)):
        like = OdeNetworkLogLikelihood(network, sampled, tspan, {obs: (np.zeros(1), None, [i])}, kind="sse")
        out[i] = np.sqrt(-like(np.asarray(Y, dtype=float)))
    return out


def test_model_0_trajectory_matches_reference_golden(cuda, g):
    """The one genuine ScipyOdeSimulator trajectory in the reference tree."""
    from gleipnir_b200 import networks
    net = networks.model_0()
    ts = g["model_0/tspan"]
    sim = _simulate_on_device(net, ["kf", "kr"], np.log10([1.5e-5, 1.2e-4]), ts, "AB_complex")
    np.testing.assert_allclose(sim[-10:], g["model_0/AB_complex_last10"], rtol=ODE_RTOL)
    _, obs = ode.simulate(net.description(), net.param_values, ts)
    np.testing.assert_allclose(sim[1:], obs[1:, 2], rtol=1e-7)            # vs the tight-tolerance oracle


@pytest.mark.parametrize("kind", ["snlpdf", "mse", "sse"])
def test_dimerization_loglike_matches_oracle(cuda, g, kind):
    """C4 inputs: Y = log10(kf, kr) drawn from the shipped priors norm(-3,1), norm(0,1)."""
    from gleipnir_b200 import networks
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    net = networks.dimerization()
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    like = OdeNetworkLogLikelihood(net, ["kf", "kr"], ts, {"A_dimer": (data, 10.0 * sd, None)}, kind=kind)
    rng = np.random.default_rng(4)
    Y = np.stack([rng.normal(-3.0, 1.0, 64), rng.normal(0.0, 1.0, 64)], axis=1)
    got = like.evaluate(Y)
    rows = [(1, data, 10.0 * sd, np.ones(51))]
    want = np.array([ode.loglike(net.description(), [0, 1], y, ts, rows, kind) for y in Y])
    np.testing.assert_allclose(got, want, rtol=ODE_RTOL)
    # scalar host protocol + device-tensor protocol agree with the batch
    assert like(Y[0]) == got[0]
    assert np.array_equal(like.evaluate(cuda.as_tensor(Y, device="cuda")).cpu().numpy(), got)


def test_stiff_parameter_corners_do_not_fail(cuda, g):
    """10 decades of rate constants (run_jarm_multinest.py:63 spans 5e-8..1.9e3): every solve must
    finish (finite logL) and agree with the oracle."""
    from gleipnir_b200 import networks
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    net = networks.dimerization()
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    like = OdeNetworkLogLikelihood(net, ["kf", "kr"], ts, {"A_dimer": (data, 10.0 * sd, None)})
    Y = np.array([[a, b] for a in (-7.3, -3.0, 1.0, 3.28) for b in (-7.3, 0.0, 3.28)])
    got = like.evaluate(Y)
    rows = [(1, data, 10.0 * sd, np.ones(51))]
    want = np.array([ode.loglike(net.description(), [0, 1], y, ts, rows) for y in Y])
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=ODE_RTOL)


def test_nested_sampling_run_on_dimerization_matches_oracle_engine(cuda, g):
    """Whole NS run with the ODE likelihood in the walkers vs the CPU oracle engine whose
    likelihood is the SciPy stand-in: same Philox counters -> same dead slots; values to 1e-6."""
    from gleipnir_b200 import networks
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    from oracle.engine import Engine
    from oracle.likelihoods import Model
    net = networks.dimerization()
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    like = OdeNetworkLogLikelihood(net, ["kf", "kr"], ts, {"A_dimer": (data, 10.0 * sd, None)})
    rows = [(1, data, 10.0 * sd, np.ones(51))]
    desc = net.description()
    table = [(pr.GL_PRIOR_NORM, -3.0, 1.0), (pr.GL_PRIOR_NORM, 0.0, 1.0)]      # run_NS_dimerization_model.py:69-71
    N, k, n_deaths = 24, 3, 36
    o = Engine(table, Model("callable", 2, fn=lambda y: ode.loglike(desc, [0, 1], y, ts, rows)), N, batch_k=k,
               iterations=3, tuning_cycles=1, seed=11)
    want = o.run(n_deaths)
    eng = DeviceEngine(table, like, N, k, n_deaths, iterations=3, tuning_cycles=1, seed=11)
    eng.init(); eng.advance(n_deaths); eng.finalize()
    tab, idx, _ = eng.dead_table()
    st = eng.state()
    assert np.array_equal(idx[:n_deaths], want["dead_index"])
    np.testing.assert_allclose(tab[:, 0], want["dead"][:, 0], rtol=ODE_RTOL)
    np.testing.assert_allclose(tab[:, 2:], want["dead"][:, 2:], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st.log_z, want["logZ"], rtol=ODE_RTOL)


def test_reference_dimerization_script_flow(cuda, g):
    """examples/pysb_dimerization_model/run_NS_dimerization_model.py with the drop-in classes."""
    from scipy.stats import norm
    from gleipnir.nestedsampling import NestedSampling
    from gleipnir.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir.sampled_parameter import SampledParameter
    from gleipnir_b200 import networks
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    loglikelihood = OdeNetworkLogLikelihood(networks.dimerization(), ["kf", "kr"], ts,
                                            {"A_dimer": (data, 10.0 * sd, None)})
    sps = [SampledParameter('kf', norm(loc=np.log10(0.001), scale=1.)),
           SampledParameter('kr', norm(loc=np.log10(1.0), scale=1.))]
    population_size = 100
    NS = NestedSampling(sampled_parameters=sps, loglikelihood=loglikelihood,
                        sampler=MetropolisComponentWiseHardNSRejection(iterations=20, tuning_cycles=1),
                        population_size=population_size, stopping_criterion=NumberOfIterations(10 * population_size),
                        seed=1)
    logz, err = NS.run()
    assert np.isfinite(logz) and err > 0
    assert NS.dead_points.shape[1] == 4 and list(NS.dead_points.columns[2:]) == ['kf', 'kr']
    # the data were generated near kf*c = 2.5e-4, kr = 1: the best fit must beat
    # the prior-centre likelihood by a wide margin
    assert NS.max_loglikelihood() > loglikelihood(np.array([-3.0, 0.0]))
    assert len(NS.posteriors()) == 2


def _jarm_setup():
    from gleipnir_b200 import networks
    net = networks.jarm_like()
    ts = np.linspace(0.0, 60.0, 181)                       # run_jarm_multinest.py:26: 181 output times
    desc = net.description()
    _, obs = ode.simulate(desc, net.param_values, ts)
    rng = np.random.default_rng(0)
    idx = np.array([6, 12, 30, 60, 120, 180])              # sparse "experimental" time points
    data = {"pTyr_jnk3": (obs[idx, 0] * (1 + 0.05 * rng.normal(size=6)), 0.1 * obs[idx, 0] + 1.0, idx),
            "pThr_jnk3": (obs[idx, 1] * (1 + 0.05 * rng.normal(size=6)), 0.1 * obs[idx, 1] + 1.0, idx)}
    sampled = [n for n in net.param_names if n.startswith("k")][:18]     # 18 sampled rate constants
    rows = []
    for o, (dat, sd, m) in zip((0, 1), data.values()):
        fd, fs, fm = np.zeros(181), np.ones(181), np.zeros(181)
        fd[m], fs[m], fm[m] = dat, sd, 1.0
        rows.append((o, fd, fs, fm))
    return net, ts, data, sampled, rows, desc


def test_jarm_like_network_loglike_matches_oracle(cuda):
    """C5-shaped problem (synthetic: the JARM model files are absent upstream): 16 species,
    24 reactions, 18 sampled log10 rate constants, 181 output times."""
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    net, ts, data, sampled, rows, desc = _jarm_setup()
    like = OdeNetworkLogLikelihood(net, sampled, ts, data)
    nominal = np.log10(np.array([net.param_values[net.param_names.index(n)] for n in sampled]))
    rng = np.random.default_rng(1)
    Y = nominal[None, :] + rng.uniform(-1.5, 1.5, size=(12, 18))
    Y[0] = nominal
    got = like.evaluate(Y)
    sidx = [net.param_names.index(n) for n in sampled]
    want = np.array([ode.loglike(desc, sidx, y, ts, rows) for y in Y])
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=ODE_RTOL)


def test_jarm_like_nested_sampling_smoke(cuda):
    from scipy.stats import uniform
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    from gleipnir_b200.sampled_parameter import SampledParameter
    net, ts, data, sampled, rows, desc = _jarm_setup()
    like = OdeNetworkLogLikelihood(net, sampled, ts, data, rtol=1e-6)
    # NestedSampleIt default priors: uniform(log10(nominal) - 2, 4)  (nestedsample_it.py:265)
    sps = [SampledParameter(n, uniform(loc=np.log10(net.param_values[net.param_names.index(n)]) - 2.0, scale=4.0))
           for n in sampled]
    ns = NestedSampling(sps, like, 256, sampler=MetropolisComponentWiseHardNSRejection(iterations=2, tuning_cycles=0),
                        stopping_criterion=NumberOfIterations(512), batch_k=64, seed=3)
    logz, err = ns.run()
    assert np.isfinite(logz)
    dp = ns.dead_points
    assert dp.shape == (512 + 256 - 64, 20)
    ll = dp['log_l'].to_numpy()
    assert np.all(np.diff(ll[:512]) >= 0.0)                  # deaths come in ascending likelihood
    assert ll[512:].min() >= ll[511]                         # survivors are above the last wall
    # every recorded log-likelihood is what the device model returns for that row
    again = like.evaluate(dp.to_numpy()[-16:, 2:])
    np.testing.assert_allclose(again, ll[-16:], rtol=1e-12)


def _random_network(rng, ns, nr, name):
    """Random uni-/bimolecular mass-action network with duplicate reactant pairs (several terms per Jacobian entry);
    rate constants <= 1 and t <= 2 keep the mass-producing steps tame."""
    from gleipnir_b200.odegen import MassActionNetwork
    sp = ["S%d" % i for i in range(ns)]
    par, rx, pairs = {}, [], []
    for r in range(nr):
        par["k%d" % r] = float(10.0 ** rng.uniform(-2, 0))
        if pairs and rng.random() < 0.3:
            reac = list(pairs[rng.integers(len(pairs))])
        elif rng.random() < 0.5:
            reac = [sp[rng.integers(ns)]]
        else:
            reac = [sp[rng.integers(ns)], sp[rng.integers(ns)]]
        pairs.append(tuple(reac))
        prod = [sp[i] for i in rng.choice(ns, size=rng.integers(1, 3), replace=False)]
        rx.append(("k%d" % r, reac, prod))
    par["S0_0"] = 3.0
    init = {s: float(rng.uniform(0.2, 2.0)) for s in sp}
    init["S0"] = "S0_0"
    return MassActionNetwork(name, species=sp, parameters=par, reactions=rx, initials=init,
                             observables={"o": {sp[0]: 1, sp[-1]: 2}, "q": {sp[ns // 2]: 1}})


@pytest.mark.parametrize("ns,nr,seed", [(12, 30, 1), (28, 70, 4), (35, 70, 7)])
def test_cooperative_solver_on_random_networks_matches_oracle(cuda, ns, nr, seed):
    """The GENERATED code (column gather of the matrix row with several terms per entry, elimination with fill,
    level-scheduled substitutions, 16- and 32-lane groups, three leaf species) on networks that are not the shipped
    ones: log-likelihoods vs the SciPy oracle."""
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    rng = np.random.default_rng(seed)
    net = _random_network(rng, ns, nr, "rnd_gpu%d" % seed)
    desc = net.description()
    ts = np.linspace(0.0, 2.0, 21)
    _, obs = ode.simulate(desc, net.param_values, ts)
    idx = np.array([4, 10, 20])
    data = {"o": (obs[idx, 0] * 1.03, 0.1 * obs[idx, 0] + 0.1, idx), "q": (obs[idx, 1] * 0.98, 0.1 * obs[idx, 1] + 0.1, idx)}
    sampled = ["k0", "k3", "k5", "k7"]
    like = OdeNetworkLogLikelihood(net, sampled, ts, data)
    sidx = [net.param_names.index(n) for n in sampled]
    nominal = np.log10(net.param_values[sidx])
    Y = nominal[None, :] + rng.uniform(-0.7, 0.4, size=(24, 4))
    Y[0] = nominal
    rows = []
    for o_, (dat, sd, m) in zip((0, 1), data.values()):
        fd, fs, fm = np.zeros(len(ts)), np.ones(len(ts)), np.zeros(len(ts))
        fd[m], fs[m], fm[m] = dat, sd, 1.0
        rows.append((o_, fd, fs, fm))
    got = like.evaluate(Y)
    want = np.array([ode.loglike(desc, sidx, y, ts, rows) for y in Y])
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=ODE_RTOL, atol=1e-9)
