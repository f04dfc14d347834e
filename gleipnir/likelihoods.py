from gleipnir_b200.likelihoods import *  # noqa: F401,F403
