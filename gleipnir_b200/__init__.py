"""gleipnir_b200 -- B200-native (sm_100a) implementation of Gleipnir's built-in classic
Nested Sampling hot path, behind the reference's own Python interface.

    from gleipnir_b200.sampled_parameter import SampledParameter
    from gleipnir_b200.nestedsampling import NestedSampling
    from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection
    from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations, RemainingPriorMass
    from gleipnir_b200.likelihoods import IIDGaussianLogLikelihood, EggCartonLogLikelihood, ...

(the ``gleipnir`` package at the repo root re-exports the same names under the reference's
module paths, so existing scripts import unchanged).  All compute runs in hand-written
CUDA kernels reached through the C ABI in ``include/gleipnir_b200.h``; there is no CPU
fallback.
"""
__version__ = "0.1.0"
