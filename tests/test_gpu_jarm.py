"""GPU parity of the JARM path (BASELINE config C5): the multi-condition ODE likelihood on the
cooperative solver (34 species: 32-row register core + 2 leaf species) against the SciPy oracle's
restatement of examples/pysb_jarm/run_jarm_multinest.py:72-143 (LSODA at rtol = atol = 1e-12), and
whole Nested Sampling runs with cooperative walkers against the CPU oracle engine.
north_star tolerance for ODE likelihoods: 1e-6 relative."""
import numpy as np
import pytest

from oracle import ode
from oracle import priors as pr

pytestmark = pytest.mark.gpu

ODE_RTOL = 1e-6


@pytest.fixture(scope="module")
def problem(cuda):
    from gleipnir_b200 import jarm, networks
    like = jarm.jarm_loglikelihood()                       # the bench's object: RODAS4 at rtol 1e-6 / atol 1e-12
    desc = like.network.description()
    nominal = np.log10(like.network.param_values[networks.JARM_CALIBRATED])
    return like, desc, like.oracle_protocol(), nominal


def _oracle(desc, proto, Y):
    return np.array([ode.protocol_loglike(desc, proto, y) for y in Y])


def test_jarm_loglike_matches_oracle_around_the_nominal_point(problem):
    like, desc, proto, nominal = problem
    rng = np.random.default_rng(1)
    Y = nominal[None, :] + rng.uniform(-0.3, 0.3, size=(16, 18))      # inside the half-normal's support mostly
    Y[0] = nominal
    got, want = like.evaluate(Y), _oracle(desc, proto, Y)
    assert np.isfinite(want[0]) and np.isfinite(want).sum() >= 4
    assert np.array_equal(np.isfinite(got), np.isfinite(want))        # -inf (half-normal support) agrees exactly
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=ODE_RTOL)
    assert like(Y[0]) == got[0]                                       # scalar host protocol == batch


def test_jarm_loglike_matches_oracle_on_prior_draws(problem):
    """10.6 decades of rate constants (run_jarm_multinest.py:63): stiff corners, mostly outside the support."""
    from gleipnir_b200 import jarm
    like, desc, proto, _ = problem
    table = jarm.jarm_prior_table()
    lo, sc = np.array([t[0] for t in table]), np.array([t[1] for t in table])
    Y = lo + sc * np.random.default_rng(7).uniform(size=(48, 18))
    got, want = like.evaluate(Y), _oracle(desc, proto, Y)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    fin = np.isfinite(want)
    assert fin.sum() >= 8
    np.testing.assert_allclose(got[fin], want[fin], rtol=ODE_RTOL)


def test_equilibrium_test_rejects_like_the_oracle(cuda):
    """eq_tol: the likelihood is -inf when the pre-equilibration has not reached max|dy/dt| <= eq_tol at t_end --
    the stand-in for the reference's pre_equilibration raising (run_jarm_multinest.py:115-119)."""
    from gleipnir_b200 import jarm, networks
    like = jarm.jarm_loglikelihood(eq_tol=1e-7)
    desc, proto = like.network.description(), like.oracle_protocol()
    nominal = np.log10(like.network.param_values[networks.JARM_CALIBRATED])
    rng = np.random.default_rng(11)
    Y = nominal[None, :] + rng.uniform(-0.2, 0.2, size=(8, 18))
    Y[1, 0] = np.log10(120.0)
    Y[2:5, 1] = np.log10(5e-8) + rng.uniform(0, 2, 3)                  # very slow exchange: no equilibrium by t = 100
    got, want = like.evaluate(Y), _oracle(desc, proto, Y)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    assert (~np.isfinite(want)).any() and np.isfinite(want).any()
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=ODE_RTOL)


def test_work_counters_count_what_ran(problem):
    like, desc, proto, nominal = problem
    like.counters(reset=True)
    like.evaluate(np.tile(nominal, (5, 1)))
    c = like.counters(reset=True)
    assert c["evals"] == 5 and c["solves"] == 20                      # 2 conditions x (pre-equilibration + main)
    assert 200 * 5 < c["steps"] < 5000 * 5
    assert like.counters()["evals"] == 0


def test_nested_sampling_on_jarm_matches_oracle_engine(problem):
    """Whole NS run, walkers = walk_coop_kernel with one warp per condition, vs the CPU oracle engine whose
    likelihood is the SciPy protocol: same Philox counters -> same dead slots, values to 1e-6."""
    from gleipnir_b200 import jarm
    from gleipnir_b200.engine import DeviceEngine
    from oracle.engine import Engine
    from oracle.likelihoods import Model
    like, desc, proto, nominal = problem
    # a box around the nominal point, inside the script's prior box: every point has a finite likelihood chance
    table = [(pr.GL_PRIOR_UNIFORM, max(lo, y - 0.25), min(lo + w, y + 0.25) - max(lo, y - 0.25))
             for y, (lo, w) in zip(nominal, jarm.jarm_prior_table())]
    N, k, n_deaths = 10, 2, 6
    o = Engine(table, Model("callable", 18, fn=lambda y: ode.protocol_loglike(desc, proto, y)), N, batch_k=k,
               iterations=1, burn_in=1, tuning_cycles=0, seed=21)
    want = o.run(n_deaths)
    eng = DeviceEngine(table, like, N, k, n_deaths, iterations=1, burn_in=1, tuning_cycles=0, seed=21)
    eng.init(); eng.advance(n_deaths); eng.finalize()
    tab, idx, _ = eng.dead_table()
    st = eng.state()
    assert np.array_equal(idx[:n_deaths], want["dead_index"])
    fin = np.isfinite(want["dead"][:, 0])
    assert np.array_equal(np.isfinite(tab[:, 0]), fin)
    np.testing.assert_allclose(tab[fin, 0], want["dead"][fin, 0], rtol=ODE_RTOL)
    np.testing.assert_allclose(tab[:, 2:], want["dead"][:, 2:], rtol=1e-9, atol=1e-12)
    if np.isfinite(want["logZ"]):
        np.testing.assert_allclose(st.log_z, want["logZ"], rtol=ODE_RTOL)


def test_cooperative_walker_whole_run_matches_oracle_engine(cuda):
    """the cooperative ODE walker (single condition, 16-lane groups, tuning cycle with pooled
    width adaptation) against the oracle engine on the 16-species network -- slots identical, values to 1e-6."""
    from gleipnir_b200 import networks
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.odegen import OdeNetworkLogLikelihood
    from oracle.engine import Engine
    from oracle.likelihoods import Model
    net = networks.jarm_like()
    ts = np.linspace(0.0, 60.0, 181)
    desc = net.description()
    _, obs = ode.simulate(desc, net.param_values, ts)
    idx = np.array([6, 12, 30, 60, 120, 180])
    rng = np.random.default_rng(0)
    data = {"pTyr_jnk3": (obs[idx, 0] * (1 + 0.05 * rng.normal(size=6)), 0.1 * obs[idx, 0] + 1.0, idx),
            "pThr_jnk3": (obs[idx, 1] * (1 + 0.05 * rng.normal(size=6)), 0.1 * obs[idx, 1] + 1.0, idx)}
    sampled = [n for n in net.param_names if n.startswith("k")][:4]
    rows = []
    for o_, (dat, sd, m) in zip((0, 1), data.values()):
        fd, fs, fm = np.zeros(181), np.ones(181), np.zeros(181)
        fd[m], fs[m], fm[m] = dat, sd, 1.0
        rows.append((o_, fd, fs, fm))
    like = OdeNetworkLogLikelihood(net, sampled, ts, data)
    sidx = [net.param_names.index(n) for n in sampled]
    nominal = np.log10(net.param_values[sidx])
    table = [(pr.GL_PRIOR_UNIFORM, y - 1.0, 2.0) for y in nominal]
    N, k, n_deaths = 16, 2, 8
    o = Engine(table, Model("callable", 4, fn=lambda y: ode.loglike(desc, sidx, y, ts, rows)), N, batch_k=k,
               iterations=2, tuning_cycles=1, seed=5)
    want = o.run(n_deaths)
    eng = DeviceEngine(table, like, N, k, n_deaths, iterations=2, tuning_cycles=1, seed=5)
    eng.init(); eng.advance(n_deaths); eng.finalize()
    tab, didx, _ = eng.dead_table()
    st = eng.state()
    assert np.array_equal(didx[:n_deaths], want["dead_index"])
    np.testing.assert_allclose(tab[:, 0], want["dead"][:, 0], rtol=ODE_RTOL)
    np.testing.assert_allclose(tab[:, 2:], want["dead"][:, 2:], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st.log_z, want["logZ"], rtol=ODE_RTOL)
    np.testing.assert_allclose(np.array(st.widths[:4]), want["widths"], rtol=1e-12)


def test_reference_jarm_script_flow(cuda):
    """examples/pysb_jarm/run_jarm_multinest.py's __main__ with the built-in sampler instead of MultiNest
    (population_size=100 as in the script, :148)."""
    from gleipnir.nestedsampling import NestedSampling
    from gleipnir.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir_b200 import jarm
    sampled_parameters, loglikelihood = jarm.jarm_problem()
    population_size = 100
    NS = NestedSampling(sampled_parameters=sampled_parameters, loglikelihood=loglikelihood,
                        population_size=population_size,
                        sampler=MetropolisComponentWiseHardNSRejection(iterations=1, burn_in=0, tuning_cycles=0),
                        stopping_criterion=NumberOfIterations(60), batch_k=20, seed=4)
    log_evidence, log_evidence_error = NS.run(verbose=False)
    dp = NS.dead_points
    assert dp.shape == (60 + 100 - 20, 20)
    ll = dp["log_l"].to_numpy()
    assert np.all(np.diff(ll[:60][np.isfinite(ll[:60])]) >= 0.0)
    fin = np.isfinite(ll)
    again = loglikelihood.evaluate(dp.to_numpy()[fin][-8:, 2:])
    np.testing.assert_allclose(again, ll[fin][-8:], rtol=1e-12)
    assert len(NS.posteriors()) == 18


def test_walker_trace_records_every_block(cuda):
    """The diagnostic trace of the cooperative walker (tools/walker_trace.py): one record per walker block and launch
    phase, SM ids inside the device, end after start, and the blocks' ODE steps add up to the plugin's counter."""
    from gleipnir_b200 import _lib as L
    from gleipnir_b200 import jarm
    from gleipnir_b200.engine import DeviceEngine
    like = jarm.jarm_loglikelihood()
    table = [(L.GL_PRIOR_UNIFORM, a, b) for a, b in jarm.jarm_prior_table()]
    N, k = 64, 8
    eng = DeviceEngine(table, like, N, k, 4 * k, iterations=1, burn_in=0, tuning_cycles=1, seed=3)
    eng.init()
    eng.kill(k)
    like.counters(reset=True)
    like.trace(True)
    eng.refill()
    eng.kill(k)
    tr = like.trace()
    like.trace(False)
    c = like.counters(reset=True)
    n_sm = L.device_info()[0]
    assert len(tr) == 2 * (k // 2)                      # two walkers per block, tuning + production launch
    assert set((tr[:, 0] >> 32).tolist()) == {0, 1}
    assert np.all((tr[:, 0] & 0xffffffff) < n_sm) and np.all(tr[:, 2] >= tr[:, 1])
    assert int(tr[:, 3].sum()) == c["steps"]
    like.trace(True)                                     # switching it on again clears it
    assert len(like.trace()) == 0
    like.trace(False)
