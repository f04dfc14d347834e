"""Model selection by Nested Sampling: ``ModelSelector`` and ``HypSelector``.

Same surface as ``gleipnir/pysb_utilities/hyp_selector.py:134-312`` and ``model_selector.py``:
``gen_nested_samplers`` builds one ``NestedSampling`` object per model through ``NestedSampleIt``,
``run_nested_sampling`` runs them and returns the evidence table (sorted by ln Z), ``bayes_factors`` the matrix
of evidence ratios, and the information-criterion tables.  Upstream runs the models one after another
(:236-248, with a commented-out process pool); here they run CONCURRENTLY on the GPU (gleipnir_b200/replicas.py):
one engine and stream per model, kernels overlapping.
"""
import warnings

import numpy as np
import pandas as pd

from ..replicas import run_many
from .nestedsample_it import NestedSampleIt


class ModelSelector(object):
    """``ModelSelector(models)`` -- model_selector.py:28-53 -- with HypSelector's methods."""

    def __init__(self, models):
        self.nested_samplers = None
        self._nested_sample_its = None
        self.selection = None
        self.models = list(models)

    def load_models(self):
        pass

    def append_to_models(self, line):
        pass

    def number_of_models(self):
        return len(self.models)

    def gen_nested_samplers(self, timespan, observable_data, solver=None, solver_kwargs=None, ns_version='built-in',
                            ns_population_size=1000, ns_kwargs=None, log_likelihood_type='snlpdf'):
        """hyp_selector.py:134-234."""
        if self.models is None:
            self.load_models()
        its, samplers = [], []
        for model in self.models:
            sample_it = NestedSampleIt(model, observable_data, timespan, solver=solver, solver_kwargs=solver_kwargs)
            samplers.append(sample_it(ns_version, ns_population_size=ns_population_size, ns_kwargs=ns_kwargs,
                                      log_likelihood_type=log_likelihood_type))
            its.append(sample_it)
        self._nested_sample_its = its
        self.nested_samplers = samplers

    def run_nested_sampling(self, max_concurrent=16):
        """hyp_selector.py:236-271; the models run concurrently on the device."""
        if self.nested_samplers is None:
            warnings.warn("Unable to run. Must call the 'gen_nested_samplers' function first!")
            return
        run_many(self.nested_samplers, max_concurrent=max_concurrent)
        frame = [dict(model="model_{}".format(i), log_evidence=ns.log_evidence, log_evidence_error=ns.log_evidence_error)
                 for i, ns in enumerate(self.nested_samplers)]
        selection = pd.DataFrame(frame)
        selection.sort_values(by=['log_evidence'], ascending=False, inplace=True)
        self.selection = selection
        return selection.reset_index(drop=True)

    def bayes_factors(self):
        """Symmetric table of evidence ratios: entry [model_j, model_i] = Z_i / Z_j (:250-271)."""
        n = self.number_of_models()
        names = ["model_{}".format(i) for i in range(n)]
        lz = np.array([ns.log_evidence for ns in self.nested_samplers])
        with np.errstate(over="ignore"):
            bf = np.exp(lz[None, :] - lz[:, None])
        np.fill_diagonal(bf, 1.0)
        return pd.DataFrame(bf, index=names, columns=names)

    def _ic_frame(self, column, values):
        frame = pd.DataFrame([{"model": "model_{}".format(i), column: v} for i, v in enumerate(values)])
        frame.sort_values(by=[column], ascending=True, inplace=True)
        return frame.reset_index(drop=True)

    def akaike_ic(self):
        return self._ic_frame('AIC', [ns.akaike_ic() for ns in self.nested_samplers])

    def _n_data(self):
        obs = self._nested_sample_its[0].observable_data
        return sum(len(obs[item][0]) for item in obs)

    def bayesian_ic(self):
        n_data = self._n_data()
        return self._ic_frame('BIC', [ns.bayesian_ic(n_data) for ns in self.nested_samplers])

    def deviance_ic(self):
        return self._ic_frame('DIC', [ns.deviance_ic() for ns in self.nested_samplers])


class HypSelector(ModelSelector):
    """``HypSelector(model_csv, hb_library=None)``: the models are enumerated by HypBuilder from a hypothesis table
    (hyp_selector.py:57-99).  HypBuilder is a separate package; without it the constructor raises, as importing
    upstream's module does."""

    def __init__(self, model_csv, hb_library=None):
        try:
            from hypbuilder import ModelAssembler            # noqa: F401
        except ImportError:
            raise ImportError("HypSelector needs HypBuilder (https://github.com/LoLab-VU/HypBuilder) to enumerate the "
                              "models; with an explicit list of models use ModelSelector")
        raise NotImplementedError("HypBuilder-driven enumeration is outside the Nested Sampling hot path; build the "
                                  "models with HypBuilder and pass them to ModelSelector")
