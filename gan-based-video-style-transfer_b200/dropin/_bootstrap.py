"""Locate and import the B200 package when only this directory is on sys.path."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
pkg = importlib.import_module("gan-based-video-style-transfer_b200")
