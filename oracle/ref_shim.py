"""TEST INFRASTRUCTURE ONLY -- harness for running the *unmodified* upstream
Gleipnir reference (read-only at /root/reference) under pandas >= 3.

Nothing in the product path imports this file.  It exists only so that
``oracle/gen_golden.py`` can record golden vectors from the real reference in
the build container (the reference tree does not travel to the GPU box).

Why a shim is needed: ``nested_sampling.py:154`` writes through
``DataFrame.values``, which is a read-only view under pandas-3 copy-on-write.
We do not edit the reference; instead the ``pd`` name inside that one module is
replaced by a proxy whose *first* ``DataFrame(...)`` call (the live-points
frame, ``nested_sampling.py:103``) returns a subclass with a cached, writable
``.values`` array -- restoring the pandas-0.2x semantics the code was written
for.  Later ``DataFrame`` calls (the dead-points frame, ``:211``) are real.
"""
import sys
import warnings
import importlib

import numpy as np
import pandas as _real_pd

REFERENCE_ROOT = "/root/reference"


class _WritableValuesFrame(_real_pd.DataFrame):
    _metadata = ["_wv"]

    @property
    def _constructor(self):
        return _WritableValuesFrame

    @property
    def values(self):
        wv = getattr(self, "_wv", None)
        if wv is None:
            wv = np.array(_real_pd.DataFrame.to_numpy(self), dtype=float, copy=True)
            object.__setattr__(self, "_wv", wv)
        return wv


class _PandasProxy(object):
    def __init__(self):
        self._first = True

    def DataFrame(self, *a, **kw):
        if self._first:
            self._first = False
            return _WritableValuesFrame(*a, **kw)
        return _real_pd.DataFrame(*a, **kw)

    def __getattr__(self, name):
        return getattr(_real_pd, name)


def load_reference():
    """Import the upstream package from REFERENCE_ROOT and return its modules."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in list(sys.modules):
        if name == "gleipnir" or name.startswith("gleipnir."):
            mod = sys.modules[name]
            if not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
                del sys.modules[name]
    nsmod = importlib.import_module("gleipnir.nestedsampling.nested_sampling")
    assert nsmod.__file__.startswith(REFERENCE_ROOT), nsmod.__file__
    return {
        "nested_sampling": nsmod,
        "samplers": importlib.import_module("gleipnir.nestedsampling.samplers"),
        "stopping_criterion": importlib.import_module("gleipnir.nestedsampling.stopping_criterion"),
        "sampled_parameter": importlib.import_module("gleipnir.sampled_parameter"),
    }


def run_reference(ns_object, nsmod, verbose=False):
    """``ns_object.run()`` with a fresh pandas proxy installed."""
    nsmod.pd = _PandasProxy()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            return ns_object.run(verbose=verbose)
    finally:
        nsmod.pd = _real_pd
