from gleipnir_b200.nsbase import NestedSamplingBase  # noqa: F401
