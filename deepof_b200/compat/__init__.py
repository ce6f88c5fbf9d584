"""Reference module names for the drop-in surface.

The reference's scripts import their siblings by bare name (``import flyingChairsWrapFlow``, ``from flyingChairsTrain_vgg import train``,
``import utils as utils`` ...).  Putting ``deepof_b200.compat.PATH`` in front of ``sys.path`` (or ``PYTHONPATH``) makes those imports
resolve to the B200 implementations without touching the scripts:

    import sys, deepof_b200.compat as compat
    sys.path.insert(0, compat.PATH)
    import flyingChairsWrapFlow          # -> deepof_b200.flyingChairsWrapFlow
"""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
MODULES = ("flyingChairsWrapFlow", "flyingChairsWrapFlow_vgg", "flyingChairsTrain", "flyingChairsTrain_vgg", "flyingChairsLoader",
           "sintelWrapFlow", "deepOF_fc", "warpflow", "Flownet", "utils")
