"""Import shim: ``import flyingChairsTrain_vgg`` (reference: flyingChairsTrain_vgg.py) resolves to ``deepof_b200.flyingChairsTrain_vgg`` when deepof_b200.compat.PATH is on sys.path."""
from deepof_b200.flyingChairsTrain_vgg import *  # noqa: F401,F403
from deepof_b200 import flyingChairsTrain_vgg as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
globals().update({n: getattr(_impl, n) for n in __all__})
