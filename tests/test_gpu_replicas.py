"""Concurrent independent runs on one GPU (gleipnir_b200/replicas.py, gl_ns_advance_many) and the model-selection
sweep on top of them (pysb_utilities.ModelSelector -- hyp_selector.py:236-271)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _egg(seed, n=600, k=60, deaths=2400):
    from scipy.stats import uniform
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations
    from gleipnir_b200.sampled_parameter import SampledParameter
    sps = [SampledParameter(i, uniform(loc=0.0, scale=10.0 * np.pi)) for i in range(2)]
    return NestedSampling(sps, EggCartonLogLikelihood(2), n, sampler=MetropolisComponentWiseHardNSRejection(
        iterations=5, tuning_cycles=1), stopping_criterion=NumberOfIterations(deaths), batch_k=k, seed=seed)


def _gauss(seed, d=5, n=300, deaths=1500):
    from scipy.stats import norm
    from gleipnir_b200.likelihoods import IIDGaussianLogLikelihood
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import RemainingPriorMass
    from gleipnir_b200.sampled_parameter import SampledParameter
    sps = [SampledParameter(i, norm(0.5, 2.0)) for i in range(d)]
    return NestedSampling(sps, IIDGaussianLogLikelihood(d), n, sampler=MetropolisComponentWiseHardNSRejection(
        iterations=4, tuning_cycles=1, proposal="normal"), stopping_criterion=RemainingPriorMass(1e-2), batch_k=10, seed=seed)


def test_concurrent_runs_equal_the_same_runs_alone(cuda):
    """Different problems, populations, batch sizes and stopping rules side by side: every run's outputs are
    bit-identical to running it alone."""
    from gleipnir_b200.replicas import run_many
    make = [lambda: _egg(1), lambda: _egg(2), lambda: _gauss(3), lambda: _egg(4, n=900, k=1, deaths=300), lambda: _gauss(5, d=3)]
    alone = [m() for m in make]
    want = [ns.run() for ns in alone]
    together = [m() for m in make]
    got = run_many(together, max_concurrent=3)              # two groups: 3 + 2
    assert got == want
    for a, b in zip(alone, together):
        assert np.array_equal(a.dead_points.to_numpy(), b.dead_points.to_numpy())
        assert a.information == b.information and a._n_evals == b._n_evals
        pa, pb = a.posteriors(), b.posteriors()
        assert all(np.array_equal(pa[key][0], pb[key][0]) for key in pa)


def test_run_replicas_gives_the_seed_scatter(cuda):
    from gleipnir_b200.replicas import evidence_statistics
    runs = _egg(0, n=1000, k=100, deaths=14000).run_replicas(range(10))
    mean, scatter, reported = evidence_statistics(runs)
    assert [r.seed for r in runs] == list(range(10))
    assert abs(mean - 235.856) < 3.0 * max(scatter, reported) / np.sqrt(10) + 0.05      # exact ln Z = 235.85594
    assert 0.3 * reported < scatter < 3.0 * reported
    single = _egg(7, n=1000, k=100, deaths=14000)
    assert single.run() == (runs[7].log_evidence, runs[7].log_evidence_error)


def test_user_defined_criterion_is_refused(cuda):
    from gleipnir_b200.replicas import run_many
    ns = _egg(1)
    ns.stopping_criterion = lambda sampler: sampler._n_iterations >= 100
    with pytest.raises(TypeError):
        run_many([ns])


def test_model_selection_sweep(cuda, golden_dir):
    """ModelSelector on two candidate networks for the dimerization data (hyp_selector.py:134-312 flow): the
    reversible dimerization that generated the data against an irreversible variant."""
    from gleipnir_b200 import networks
    from gleipnir_b200.odegen import MassActionNetwork
    from gleipnir_b200.pysb_utilities import ModelSelector
    g = np.load(os.path.join(golden_dir, "ref_ode.npz"))
    data, sd, ts = g["dimerization/data"], g["dimerization/sd"], g["dimerization/tspan"]
    irreversible = MassActionNetwork("dimerization_irreversible", species=["A", "AA"],
                                     parameters={"kf": 0.001, "A_0": 2000.0}, reactions=[("kf", ["A", "A"], ["AA"])],
                                     initials={"A": "A_0"}, observables={"A_free": {"A": 1}, "A_dimer": {"AA": 1}})
    sel = ModelSelector([networks.dimerization(), irreversible])
    sel.gen_nested_samplers(ts, {"A_dimer": (data, 10.0 * sd, None)}, ns_population_size=100, ns_kwargs=dict(seed=5))
    assert sel.number_of_models() == 2 and len(sel.nested_samplers) == 2
    assert [len(ns.sampled_parameters) for ns in sel.nested_samplers] == [2, 1]
    table = sel.run_nested_sampling()
    assert list(table.columns) == ["model", "log_evidence", "log_evidence_error"]
    assert table["log_evidence"].iloc[0] >= table["log_evidence"].iloc[1]
    assert np.all(np.isfinite(table["log_evidence"])) and np.all(table["log_evidence_error"] > 0)
    bf = sel.bayes_factors()
    lz = [ns.log_evidence for ns in sel.nested_samplers]
    assert bf.loc["model_0", "model_1"] == pytest.approx(np.exp(lz[1] - lz[0]))
    assert list(sel.akaike_ic().columns) == ["model", "AIC"] and len(sel.bayesian_ic()) == 2 and len(sel.deviance_ic()) == 2
    # each run equals the same NestedSampleIt run on its own
    from gleipnir_b200.pysb_utilities import NestedSampleIt
    alone = NestedSampleIt(networks.dimerization(), {"A_dimer": (data, 10.0 * sd, None)}, ts)(
        'built-in', ns_population_size=100, ns_kwargs=dict(seed=5))
    assert alone.run() == (sel.nested_samplers[0].log_evidence, sel.nested_samplers[0].log_evidence_error)


def test_graph_recorded_run_equals_the_directly_launched_run(cuda):
    """gl_ns_advance_graph: the whole run recorded into a CUDA graph (programmatic-dependent-launch edges included)
    and launched as one unit gives the bit-identical run."""
    import torch
    from gleipnir_b200 import _lib as L
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.likelihoods import EggCartonLogLikelihood
    lib = L.load()
    table = [(L.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * 2
    like = EggCartonLogLikelihood(2)
    out = []
    for graph in (False, True):
        eng = DeviceEngine(table, like, 2000, 200, 12000, iterations=8, tuning_cycles=2, seed=6, stream=torch.cuda.Stream())
        torch.cuda.synchronize()
        eng.init()
        if graph:
            L.check(lib.gl_ns_advance_graph(eng._h, 12000, eng._stream()))
        else:
            eng.advance(12000)
        eng.finalize()
        st = eng.state()
        out.append((st.log_z, st.information, st.n_evals, eng.dead_table()[0]))
    assert out[0][:3] == out[1][:3] and np.array_equal(out[0][3], out[1][3])
