"""Drop-in for methods/optimization-based/flowtools.py (occlusion test disabled, :35,41-42,45,55):
``import flowtools_obst as flowtools`` in obst_eval.py."""
from _bootstrap import pkg as _pkg

gradient = _pkg.gradient
warp = _pkg.warp
fbcCheckTorch = _pkg.fbcCheckTorch_mob
