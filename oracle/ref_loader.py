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

REFERENCE_ROOT = os.environ.get("ITR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "itr", "modalmodule"))


def load():
    """Returns (Objectives, evaluation) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in ("nltk", "pycocotools", "pycocotools.coco"):
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
