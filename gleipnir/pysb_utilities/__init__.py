from gleipnir_b200.pysb_utilities import HypSelector, ModelSelector, NestedSampleIt, NestIt  # noqa: F401
