from gleipnir_b200.nestedsampling.nested_sampling import NestedSampling  # noqa: F401
