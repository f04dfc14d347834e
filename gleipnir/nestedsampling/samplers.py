from gleipnir_b200.nestedsampling.samplers import MetropolisComponentWiseHardNSRejection  # noqa: F401
