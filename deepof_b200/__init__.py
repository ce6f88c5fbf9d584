"""deepof_b200 -- B200-native (sm_100a) implementation of the deepOF training hot path.

Host side in Python (the reference is Python), binding hand-written CUDA through the C ABI in
``include/deepof_b200.h`` with ctypes.  PyTorch is used for device memory, streams and
``torch.distributed`` only.  Public surface (mirrors the reference's call sites):

* ``deepof_b200.flyingChairsWrapFlow.flowNet / loss_interp``  (flyingChairsWrapFlow.py:5,752)
* ``deepof_b200.warpflow.loss_interp``                         (version1/model/warpflow.py:4)
* ``deepof_b200.flyingChairsTrain.train`` / ``TrainStep``      (flyingChairsTrain.py:94-213)
"""
from ._lib import DeepOFError, load as load_library  # noqa: F401

__all__ = ["DeepOFError", "load_library"]
