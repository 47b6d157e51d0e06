"""Import shim: ``import tcl_b200`` == the package in ``gan-based-video-style-transfer_b200/``.

The package directory keeps the name the build contract prescribes, which is not a Python
identifier; this module loads it with importlib and re-exports it under an importable name.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("gan-based-video-style-transfer_b200")
sys.modules[__name__] = _pkg
