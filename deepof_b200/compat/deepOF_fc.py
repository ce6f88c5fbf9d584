"""Import shim: ``import deepOF_fc`` (reference: deepOF_fc.py) resolves to ``deepof_b200.deepOF_fc`` when deepof_b200.compat.PATH is on sys.path."""
from deepof_b200.deepOF_fc import *  # noqa: F401,F403
from deepof_b200 import deepOF_fc as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
globals().update({n: getattr(_impl, n) for n in __all__})
