"""Import the UNMODIFIED reference (authoring container only; TEST INFRASTRUCTURE).

``itr.metricmodule.evaluation`` pulls in nltk and pycocotools through
``itr.datamodule`` (itr/datamodule/data_loader.py:5,11); neither is installed
and neither is used by the arithmetic, so empty stub modules are registered
first.  /root/reference is read-only: bytecode
writing is disabled.
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _root():
    """/root/reference where it exists (authoring container), else the byte-identical copy oracle/build_ref.py
    staged under oracle/_ref (what travels to the GPU box)."""
    env = os.environ.get("ITR_REFERENCE_ROOT", "/root/reference")
    for cand in (env, _STAGED):
        if os.path.isdir(os.path.join(cand, "itr", "modalmodule")):
            return cand
    return env


REFERENCE_ROOT = _root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "itr", "modalmodule"))


def load():
    """Returns (Objectives, evaluation) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in ("nltk", "pycocotools", "pycocotools.coco", "tensorboard_logger"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["pycocotools.coco"], "COCO"):
        sys.modules["pycocotools.coco"].COCO = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from itr.modalmodule import Objectives
        from itr.metricmodule import evaluation
    return Objectives, evaluation


class _AnyDecorator:
    """Stand-in for sacred.Experiment: every attribute is an identity decorator (itr/config.py only decorates)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda f=None, *a, **k: f


def load_utils():
    """The reference's ``itr.utils`` (validate_step, utils.py:144-186).  It imports sacred (through itr/config.py) and
    tensorboard_logger, neither installed nor used by the arithmetic: inert stubs are registered first."""
    load()
    if "sacred" not in sys.modules:
        m = types.ModuleType("sacred")
        m.Experiment = _AnyDecorator
        sys.modules["sacred"] = m
    tb = sys.modules.get("tensorboard_logger")
    if tb is None or not hasattr(tb, "log_value"):
        tb = types.ModuleType("tensorboard_logger")
        sys.modules["tensorboard_logger"] = tb
    if not hasattr(tb, "log_value"):
        tb.logged = {}
        tb.log_value = lambda name, value, step=None: tb.logged.__setitem__(name, value)
        tb.configure = lambda *a, **k: None
    import importlib
    return importlib.import_module("itr.utils")
