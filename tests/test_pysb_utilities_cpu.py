"""Host side of the pysb_utilities front end (no GPU, no PySB): a duck-typed stand-in for a generated PySB model
is converted to the same network as the hand transcription, NestIt / NestedSampleIt reproduce upstream's
defaults (nestedsample_it.py:93-99,257-266,416-427) and the selector's tables follow hyp_selector.py:250-312."""
import numpy as np
import pytest
from scipy.stats import norm, uniform

from gleipnir_b200 import networks
from gleipnir_b200.pysb_utilities import ModelSelector, NestedSampleIt, NestIt, network_from_pysb


class _Par(object):
    def __init__(self, name, value):
        self.name, self.value = name, value


class _Rule(object):
    def __init__(self, fwd, rev=None):
        self.rate_forward, self.rate_reverse = fwd, rev


class _Obs(object):
    def __init__(self, name, species, coefficients):
        self.name, self.species, self.coefficients = name, species, coefficients


class _Ini(object):
    def __init__(self, pattern, value):
        self.pattern, self.value = pattern, value


class FakeDimerizationModel(object):
    """What examples/pysb_dimerization_model/dimerization_model.py looks like after generate_equations: species
    A(d=None) and A(d=1) % A(d=1); reactions 2 A -> AA with rate kf*s0**2/2 (BioNetGen's symmetry factor) and
    AA -> 2 A with rate kr*s1."""

    def __init__(self):
        import sympy
        kf, kr, a0 = _Par("kf", 0.001), _Par("kr", 1.0), _Par("A_0", 2000.0)
        s0, s1 = sympy.Symbol("__s0"), sympy.Symbol("__s1")
        self.name = "dimerization"
        self.parameters = [kf, kr, a0]
        self.species = ["A", "AA"]
        self.rules = [_Rule(kf, kr)]
        self.reactions = [dict(reactants=(0, 0), products=(1,), rate=sympy.Symbol("kf") * s0 ** 2 / 2),
                          dict(reactants=(1,), products=(0, 0), rate=sympy.Symbol("kr") * s1)]
        self.observables = [_Obs("A_free", [0], [1]), _Obs("A_dimer", [1], [1])]
        self.initials = [_Ini("A", a0)]


def test_network_from_a_pysb_like_model_equals_the_hand_transcription():
    got = network_from_pysb(FakeDimerizationModel()).description()
    want = networks.dimerization().description()
    assert got["species"] == want["species"] and got["param_names"] == want["param_names"]
    assert np.array_equal(got["param_values"], want["param_values"])
    assert got["reactions"] == want["reactions"]                  # incl. the 1/2! statistical factor
    assert got["initials"] == want["initials"] and got["observables"] == want["observables"]
    # the generated CUDA is therefore the same translation unit
    assert network_from_pysb(FakeDimerizationModel()).plugin_source() == networks.dimerization().plugin_source()


def test_nestit_defaults_and_registry():
    ni = NestIt()
    kf, kr = _Par("kf", 0.001), _Par("kr", 1.0)
    ni(kf)
    ni.default_to_norm_prior()
    ni(kr)
    ni += _Par("kc", 10.0)
    assert ni.names() == ["kf", "kr", "kc"] and "kf" in ni
    assert ni["kf"].dist.name == "uniform" and np.allclose(ni["kf"].args or list(ni["kf"].kwds.values()), [-5.0, 4.0])
    assert ni["kr"].dist.name == "norm" and np.allclose(list(ni["kr"].kwds.values()), [0.0, 2.0])
    ni -= "kc"
    assert ni.mask([kf, _Par("x", 1.0), kr]) == [True, False, True]
    sps = ni.sampled_parameters()
    assert [sp.name for sp in sps] == ["kf", "kr"]
    ni["kf"] = norm(-3.0, 1.0)
    assert ni.priors()[0].dist.name == "norm"


def test_nestedsampleit_builds_the_reference_configuration():
    ts = np.linspace(0, 1, 51)
    data = {"A_dimer": (np.linspace(0, 500, 51), np.full(51, 10.0), None)}
    it = NestedSampleIt(FakeDimerizationModel(), data, ts)
    # default: every rule's rate constants, uniform(log10(value) - 2, 4)   (nestedsample_it.py:257-266)
    assert [sp.name for sp in it._sampled_parameters] == ["kf", "kr"]
    pr = it._sampled_parameters[0].prior_dist
    assert pr.dist.name == "uniform" and np.allclose(list(pr.kwds.values()), [-5.0, 4.0])
    ns = it('built-in', ns_population_size=200, ns_kwargs=dict(seed=3))
    assert ns.population_size == 200 and ns.stopping_criterion.n_iterations == 2000          # 10 * N  (:421)
    assert (ns.sampler.iterations, ns.sampler.burn_in, ns.sampler.tuning_cycles) == (10, 10, 1)   # :416-418
    assert ns.loglikelihood.kind == "snlpdf" and ns.loglikelihood.ndim == 2 and ns.seed == 3
    assert it('built-in', log_likelihood_type='mse').loglikelihood.kind == "mse"
    with pytest.raises(NotImplementedError):
        it('multinest')
    with pytest.raises(TypeError):
        it(custom_loglikelihood=lambda model, sim: 0.0)
    # explicit NestIt: only the registered parameters, with their priors
    ni = NestIt()
    ni(_Par("kr", 1.0), norm(0.0, 1.0))
    it2 = NestedSampleIt(networks.dimerization(), data, ts, nest_it=ni)
    assert [sp.name for sp in it2._sampled_parameters] == ["kr"] and it2._rate_mask == [False, True, False]


class _Done(object):
    def __init__(self, lz, err, aic):
        self.log_evidence, self.log_evidence_error, self._aic = lz, err, aic

    def akaike_ic(self):
        return self._aic

    def bayesian_ic(self, n):
        return self._aic + n

    def deviance_ic(self):
        return -self._aic


def test_selector_tables_follow_the_reference():
    sel = ModelSelector(["m0", "m1", "m2"])
    sel.nested_samplers = [_Done(-10.0, 0.1, 30.0), _Done(-8.0, 0.2, 20.0), _Done(-12.0, 0.3, 25.0)]

    class _It(object):
        observable_data = {"a": (np.zeros(7), None, None), "b": (np.zeros(5), None, None)}
    sel._nested_sample_its = [_It()]
    bf = sel.bayes_factors()
    assert list(bf.columns) == ["model_0", "model_1", "model_2"]
    assert bf.loc["model_0", "model_1"] == pytest.approx(np.exp(-8.0 + 10.0))       # [j, i] = Z_i / Z_j
    assert bf.loc["model_1", "model_0"] == pytest.approx(np.exp(-2.0)) and bf.loc["model_2", "model_2"] == 1.0
    assert list(sel.akaike_ic()["model"]) == ["model_1", "model_2", "model_0"]
    assert list(sel.bayesian_ic()["BIC"]) == [32.0, 37.0, 42.0]                      # n_data = 12
    assert list(sel.deviance_ic()["model"]) == ["model_0", "model_2", "model_1"]
    assert sel.number_of_models() == 3
