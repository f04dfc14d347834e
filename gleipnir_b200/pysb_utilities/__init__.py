"""B200 counterparts of ``gleipnir.pysb_utilities``: ``NestIt`` / ``NestedSampleIt``
(nestedsample_it.py:65-452) and the model-selection sweeps ``ModelSelector`` / ``HypSelector``
(hyp_selector.py:134-312, model_selector.py).  PySB itself is optional: a model is anything that exposes the
generated reaction network the way a ``pysb.Model`` does after ``generate_equations`` (see
``network_from_pysb``), or a ``MassActionNetwork`` directly."""
from .nestedsample_it import NestedSampleIt, NestIt, network_from_pysb  # noqa: F401
from .hyp_selector import HypSelector, ModelSelector  # noqa: F401
