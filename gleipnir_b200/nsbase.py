"""``NestedSamplingBase`` -- host mirror of the abstract surface in ``gleipnir/nsbase.py``
plus its concrete helpers (posterior_moments, akaike_ic, bayesian_ic, best_fit_posterior;
nsbase.py:66-85,93-130,157-175).  These are thin post-processing of ``posteriors()`` /
``max_loglikelihood()`` results and run on the host.
"""
from abc import ABC, abstractmethod

import numpy as np
import scipy.stats


class NestedSamplingBase(ABC):

    @abstractmethod
    def __init__(self, sampled_parameters, loglikelihood, population_size):
        self.sampled_parameters = sampled_parameters
        self.loglikelihood = loglikelihood
        self.population_size = population_size
        self._post_eval = False
        self._posteriors = None

    @abstractmethod
    def run(self, verbose=False):
        pass

    @property
    @abstractmethod
    def evidence(self):
        pass

    @property
    @abstractmethod
    def evidence_error(self):
        pass

    @property
    @abstractmethod
    def log_evidence(self):
        pass

    @property
    @abstractmethod
    def log_evidence_error(self):
        pass

    @property
    @abstractmethod
    def information(self):
        pass

    @abstractmethod
    def posteriors(self, nbins=None):
        pass

    @abstractmethod
    def max_loglikelihood(self):
        pass

    @abstractmethod
    def deviance_ic(self):
        pass

    @abstractmethod
    def best_fit_likelihood(self):
        pass

    def posterior_moments(self, n_resample=1000000):
        """(mean, var, skew, kurtosis) of each marginal, by resampling its histogram."""
        out = dict()
        for name, (marginal, edges, centers) in self.posteriors().items():
            width = edges[1] - edges[0]
            draws = np.random.choice(centers, size=n_resample, p=marginal / marginal.sum())
            draws = draws + width * (np.random.random(n_resample) - 0.5)
            out[name] = (np.mean(draws), np.var(draws), scipy.stats.skew(draws), scipy.stats.kurtosis(draws))
        return out

    def akaike_ic(self):
        """AIC = 2k - 2*max logL."""
        return 2.0 * len(self.sampled_parameters) - 2.0 * self.max_loglikelihood()

    def bayesian_ic(self, n_data):
        """BIC = ln(n_data)*k - 2*max logL."""
        return np.log(n_data) * len(self.sampled_parameters) - 2.0 * self.max_loglikelihood()

    def best_fit_posterior(self):
        """Centre of the highest marginal-posterior bin per parameter, +- half a bin."""
        best, half_widths = [], []
        for marginal, edges, centers in self.posteriors().values():
            best.append(centers[np.argmax(marginal)])
            half_widths.append(0.5 * (edges[1] - edges[0]))
        return np.array(best), np.array(half_widths)
