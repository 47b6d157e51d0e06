"""Drop-in for the hot-path names of utils/sintel_eval.py: ``from sintel_eval import computeTCL, save_dict_as_json``
(StarGANv2AdvCon/core/solver.py:36).  RAFT setup, datasets and image saving stay in the caller's code base."""
from _bootstrap import pkg as _pkg

computeTCL = _pkg.computeTCL
save_dict_as_json = _pkg.save_dict_as_json
