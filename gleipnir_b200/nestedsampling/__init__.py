from .nested_sampling import NestedSampling

__all__ = ["NestedSampling"]
