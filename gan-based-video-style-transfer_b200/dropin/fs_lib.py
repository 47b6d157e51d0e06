"""Drop-in for methods/learning-based/fs_lib.py: ``from fs_lib import warp`` (fs_ruder.py:4, fs_huang.py:4, fs_reconet.py:4)."""
from _bootstrap import pkg as _pkg

warp = _pkg.fs_warp
