"""Drop-in alias: the reference's module paths (``gleipnir.sampled_parameter``,
``gleipnir.nestedsampling{,.samplers,.stopping_criterion}``, ``gleipnir.nsbase``) served by
the B200-native implementation in ``gleipnir_b200``.  Existing scripts keep their imports;
only the log-likelihood has to become a device log-likelihood object
(``gleipnir.likelihoods`` / ``gleipnir_b200.likelihoods``)."""
from gleipnir_b200 import __version__  # noqa: F401
