from gleipnir_b200.nestedsampling import NestedSampling  # noqa: F401
