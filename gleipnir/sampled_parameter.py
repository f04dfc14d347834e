from gleipnir_b200.sampled_parameter import *  # noqa: F401,F403
from gleipnir_b200.sampled_parameter import SampledParameter  # noqa: F401
