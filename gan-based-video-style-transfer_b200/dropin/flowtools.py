"""Drop-in for the reference's ``flowtools`` module (utils/flowtools.py and its 7 vendored copies).

Put this directory first on ``sys.path`` (or copy the three files next to a method's ``main.py``) and
``from flowtools import fbcCheckTorch, warp`` (utils/sintel_eval.py:29, MoGAN/models/cycle_gan_model.py:14,
fast_style_transfer.py:29) resolves to the B200 kernels.  Same names, argument meaning and results.
"""
from _bootstrap import pkg as _pkg

gradient = _pkg.gradient
warp = _pkg.warp
fbcCheckTorch = _pkg.fbcCheckTorch
