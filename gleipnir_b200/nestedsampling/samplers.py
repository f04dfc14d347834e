"""Samplers -- host mirror of ``gleipnir/nestedsampling/samplers.py``.

``MetropolisComponentWiseHardNSRejection`` keeps upstream's constructor, public
attributes (``iterations``, ``burn_in``, ``tuning_cycles``, ``proposal``) and private
state names (``_first``, ``_widths``, ``_ndim``).  On the B200 path the object is a bag of
hyper-parameters for the walker kernel (``csrc/walker.cu``): thousands of independent
constrained Metropolis walkers per step, each applying upstream's tuning-loop move
(samplers.py:95-114) -- propose one component, accept iff ``u < prior ratio`` and
``logL > L*`` -- during tuning and production alike.

Differences from upstream, all deliberate (DESIGN.md "Sampler semantics"):
  Q1  the proposed value IS written into the trial point (upstream's production chain
      evaluates prior and likelihood at the current point, samplers.py:131-149);
  Q2  widths adapt once per 20-sweep tuning cycle from the acceptance pooled over the
      step's walkers (upstream shrinks x0.66 on every sweep, samplers.py:115-121);
  RNG Philox4x32-10 counters instead of numpy's global MT19937.
"""
import numpy as np


class MetropolisComponentWiseHardNSRejection(object):
    """Component-wise Metropolis with a hard nested-sampling likelihood wall."""

    def __init__(self, iterations=100, burn_in=0, tuning_cycles=0, proposal='uniform'):
        self.iterations = iterations
        self.burn_in = burn_in
        self.tuning_cycles = tuning_cycles
        self.proposal = proposal
        self._first = True
        self._widths = list()
        self._ndim = None

    def device_config(self):
        """Hyper-parameters as the engine wants them."""
        if self.proposal not in ('uniform', 'normal'):
            raise ValueError("proposal must be 'uniform' or 'normal'")
        return dict(iterations=int(self.iterations), burn_in=int(self.burn_in),
                    tuning_cycles=int(self.tuning_cycles), proposal=0 if self.proposal == 'uniform' else 1)

    def __call__(self, sampled_parameters, loglikelihood, start_param_vec, ns_boundary, **kwargs):
        """Upstream's per-call protocol (samplers.py:56): one walker from ``start_param_vec``
        under the wall ``ns_boundary``; returns ``(point, logL)``.  Runs the same CUDA walker
        kernel with a batch of one (population = the start point), so it needs a device
        log-likelihood object."""
        from .single_walker import run_single_walker
        point, logl, widths = run_single_walker(self, sampled_parameters, loglikelihood,
                                                np.asarray(start_param_vec, dtype=np.float64),
                                                float(ns_boundary), seed=kwargs.get('seed'))
        self._ndim = len(sampled_parameters)
        self._widths = widths
        self._first = False
        return point, logl
