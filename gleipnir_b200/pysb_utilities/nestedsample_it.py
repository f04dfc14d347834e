"""``NestIt`` and ``NestedSampleIt`` for the B200 engine.

Same front end as ``gleipnir/pysb_utilities/nestedsample_it.py:65-452``: a model plus
``observable_data = {observable: (data, sd or None, time indices or None)}`` and a timespan in, a ready
``NestedSampling`` object out (``ns_version='built-in'``: sampler(iterations=10, burn_in=10,
tuning_cycles=1), ``NumberOfIterations(10*population)``, :416-427).  What changes is what stands behind the
likelihood: instead of a Python closure over ``ScipyOdeSimulator.run`` (:284-339) the model's reaction network
is code-generated into a CUDA integrator (gleipnir_b200/odegen.py) with the misfit fused in.

``model`` may be
  * a ``MassActionNetwork``, or
  * a PySB model whose reaction network has been generated (``pysb.bng.generate_equations(model)`` is called
    when PySB is importable), or any object with the same attributes -- see ``network_from_pysb``.
"""
import warnings

import numpy as np
from scipy.stats import norm, uniform

from ..odegen import MassActionNetwork, OdeNetworkLogLikelihood
from ..sampled_parameter import SampledParameter


class NestIt(object):
    """Registry of the parameters to sample and their (log10-space) priors -- nestedsample_it.py:65-186."""

    def __init__(self):
        self.parms = dict()
        self._default_prior = 'uniform'

    def __call__(self, parameter, prior=None):
        if prior is None:
            if self._default_prior == 'norm':
                prior = norm(loc=np.log10(parameter.value), scale=2.0)                  # :93-95
            else:
                prior = uniform(loc=np.log10(parameter.value) - 2.0, scale=4.0)         # :96-98
        self.parms[parameter.name] = prior
        return parameter

    def __getitem__(self, key):
        return self.parms[key]

    def __setitem__(self, key, prior):
        self.parms[key] = prior

    def __delitem__(self, key):
        del self.parms[key]

    def __contains__(self, key):
        return key in self.names()

    def __iadd__(self, parm):
        self.__call__(parm)
        return self

    def __isub__(self, parm):
        self.__delitem__(getattr(parm, "name", parm))
        return self

    def names(self):
        return list(self.parms.keys())

    def keys(self):
        return self.parms.keys()

    def mask(self, model_parameters):
        names = self.names()
        return [(parm.name in names) for parm in model_parameters]

    def priors(self):
        return [self.parms[name] for name in self.parms.keys()]

    def sampled_parameters(self):
        return [SampledParameter(name, self.parms[name]) for name in self.keys()]

    def default_to_norm_prior(self):
        self._default_prior = 'norm'

    def default_to_uniform_prior(self):
        self._default_prior = 'uniform'

    def add_all_kinetic_params(self, pysb_model):
        for rule in pysb_model.rules:
            for rate in (rule.rate_forward, getattr(rule, "rate_reverse", None)):
                if rate:
                    try:
                        self.__call__(rate)
                    except Exception:
                        warnings.warn('Ignoring kinetic rate defined as a PySB Expression.')

    def add_all_nonkinetic_params(self, pysb_model):
        kinetic = []
        for rule in pysb_model.rules:
            for rate in (rule.rate_forward, getattr(rule, "rate_reverse", None)):
                if rate:
                    kinetic.append(rate)
        for param in pysb_model.parameters:
            if param not in kinetic:
                self.__call__(param)

    def add_by_name(self, pysb_model, name_or_list):
        names = name_or_list if isinstance(name_or_list, (list, tuple)) else [name_or_list]
        for param in pysb_model.parameters:
            if param.name in names:
                self.__call__(param)


class _P(object):
    def __init__(self, name, value):
        self.name, self.value = name, float(value)


def _rate_of(reaction, param_names, species_symbols):
    """(rate parameter name, statistical factor or None) of one entry of ``model.reactions``.  PySB stores the rate
    as a sympy expression ``k * __s0 * __s1`` (times 1/m! for m identical reactants); a plain description may give
    ``rate_param`` (and ``stat_factor``) directly."""
    if "rate_param" in reaction:
        return reaction["rate_param"], reaction.get("stat_factor")
    rate = reaction["rate"]
    if hasattr(rate, "name") and rate.name in param_names:           # a bare Parameter
        return rate.name, None
    free = getattr(rate, "free_symbols", None)
    if free is None:
        raise TypeError("cannot read the rate law of reaction %r" % (reaction,))
    ks = [s for s in free if str(s) in param_names]
    if len(ks) != 1:
        raise ValueError("only mass-action rate laws with one rate constant are supported: %s" % rate)
    expr = rate / ks[0]
    for idx in reaction["reactants"]:
        expr = expr / species_symbols(idx)
    factor = float(expr)                                              # what is left is the statistical factor
    return str(ks[0]), factor


def network_from_pysb(model, name=None):
    """``MassActionNetwork`` of a PySB model (or a look-alike).  Reads, after network generation,
    ``model.species`` (their ``str`` is the species name), ``model.parameters`` (``.name``, ``.value``),
    ``model.reactions`` (dicts with ``reactants`` / ``products`` as species indices and ``rate``),
    ``model.observables`` (``.name``, ``.species`` indices, ``.coefficients``) and ``model.initials``
    (``.pattern`` / ``.value``; older PySB: ``model.initial_conditions`` pairs)."""
    if isinstance(model, MassActionNetwork):
        return model
    if not getattr(model, "species", None):
        try:
            from pysb.bng import generate_equations
            generate_equations(model)
        except ImportError:
            raise ValueError("the model has no generated species/reactions and PySB (BioNetGen) is not available "
                             "to generate them")
    species = [str(s) for s in model.species]
    params = {p.name: p.value for p in model.parameters if hasattr(p, "value")}
    try:
        import sympy
        sym = lambda i: sympy.Symbol("__s%d" % i)
    except ImportError:          # pragma: no cover
        sym = None
    reactions = []
    for r in model.reactions:
        k, sf = _rate_of(r, params, sym)
        reac = [species[i] for i in r["reactants"]]
        prod = [species[i] for i in r["products"]]
        reactions.append((k, reac, prod) if sf is None else (k, reac, prod, sf))
    initials = {}
    if getattr(model, "initials", None):
        for ini in model.initials:
            initials[str(ini.pattern)] = ini.value.name if hasattr(ini.value, "name") else float(ini.value)
    else:
        for pattern, value in getattr(model, "initial_conditions", []):
            initials[str(pattern)] = value.name if hasattr(value, "name") else float(value)
    observables = {o.name: {species[i]: c for i, c in zip(o.species, o.coefficients)} for o in model.observables}
    return MassActionNetwork(name or getattr(model, "name", "pysb_model").replace(".", "_"), species, params, reactions,
                             initials, observables)


class NestedSampleIt(object):
    """Builds ``NestedSampling`` objects for reaction-network models -- nestedsample_it.py:188-452.

    Args as upstream: ``model``, ``observable_data``, ``timespan``, ``solver`` / ``solver_kwargs`` (the CUDA
    integrator takes ``rtol`` / ``atol`` / ``max_steps`` from ``solver_kwargs``; a ``solver`` class is ignored),
    ``nest_it`` or ``builder`` naming the sampled parameters and their priors; by default every rule's rate
    constants with ``uniform(log10(value) - 2, 4)`` (:257-266)."""

    def __init__(self, model, observable_data, timespan, solver=None, solver_kwargs=None, nest_it=None, builder=None):
        self.model = model
        self.observable_data = observable_data
        self.timespan = np.asarray(timespan, dtype=np.float64)
        self.solver = solver
        self.solver_kwargs = dict(solver_kwargs or {})
        self._ns_kwargs = None
        self.network = network_from_pysb(model)
        net = self.network
        parameters = [_P(n, v) for n, v in zip(net.param_names, net.param_values)]
        if nest_it is not None:
            mask = nest_it.mask(parameters)
            self._sampled_parameters = [SampledParameter(p.name, nest_it[p.name]) for i, p in enumerate(parameters) if mask[i]]
        elif builder is not None:
            pnames = [p.name for p in builder.estimate_params]
            self._sampled_parameters = [SampledParameter(p.name, builder.priors[pnames.index(p.name)])
                                        for p in parameters if p.name in pnames]
        else:
            if hasattr(model, "rules"):
                kinetic = []
                for rule in model.rules:
                    for rate in (rule.rate_forward, getattr(rule, "rate_reverse", None)):
                        if rate:
                            kinetic.append(rate.name)
            else:                                        # a bare network: the rate constants of its reactions
                kinetic = [r[0] for r in net.reactions]
            self._sampled_parameters = [SampledParameter(p.name, uniform(loc=np.log10(p.value) - 2.0, scale=4.0))
                                        for p in parameters if p.name in kinetic]
        self._rate_mask = [p.name in {sp.name for sp in self._sampled_parameters} for p in parameters]
        self._param_values = net.param_values.copy()

    def loglikelihood(self, log_likelihood_type='snlpdf'):
        """Device likelihood object for 'snlpdf' | 'mse' | 'sse' (:273-339)."""
        kw = {k: self.solver_kwargs[k] for k in ("rtol", "atol", "max_steps") if k in self.solver_kwargs}
        return OdeNetworkLogLikelihood(self.network, [sp.name for sp in self._sampled_parameters], self.timespan,
                                       self.observable_data, kind=log_likelihood_type, **kw)

    def __call__(self, ns_version='built-in', ns_population_size=1000, ns_kwargs=None, log_likelihood_type='snlpdf',
                 custom_loglikelihood=None):
        if custom_loglikelihood is not None:
            raise TypeError("a custom Python loglikelihood cannot run inside a CUDA kernel: pass "
                            "log_likelihood_type='snlpdf' | 'mse' | 'sse' (gleipnir_b200 has no CPU fallback)")
        if ns_version != 'built-in':
            raise NotImplementedError("gleipnir_b200 implements Gleipnir's built-in Nested Sampling; the third-party "
                                      "engines (%r) are out of its scope" % (ns_version,))
        if log_likelihood_type not in ('snlpdf', 'mse', 'sse'):
            log_likelihood_type = 'snlpdf'                               # upstream's else-branch (:409-410)
        self._ns_kwargs = dict(ns_kwargs or {})
        from ..nestedsampling import NestedSampling
        from ..nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
        from ..nestedsampling.stopping_criterion import NumberOfIterations
        population_size = ns_population_size
        sampler = MetropolisComponentWiseHardNSRejection(iterations=10, burn_in=10, tuning_cycles=1)     # :416-418
        stopping_criterion = NumberOfIterations(10 * population_size)                                      # :421
        return NestedSampling(sampled_parameters=self._sampled_parameters,
                              loglikelihood=self.loglikelihood(log_likelihood_type), sampler=sampler,
                              population_size=population_size, stopping_criterion=stopping_criterion,
                              **self._ns_kwargs)
