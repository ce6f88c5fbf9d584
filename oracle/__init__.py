"""CPU oracle for the deepOF hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch/NumPy (CPU, fp32 or fp64) restatement of the
reference's algorithm for the FlowNet guided-optical-flow training step:

    pre-processing -> FlowNetS conv tower / refinement -> per-scale bilinear
    flow-warp -> Charbonnier photometric + smoothness loss -> backward -> Adam

It exists to CHECK the CUDA product path (deepof_b200/).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``deepof_b200/`` imports it, and
the product path raises if the CUDA library is missing instead of falling back
to this code.

PARITY UNPINNED.  The reference (bryanyzhu/deepOF @ 7d6c12b) is Python-2 /
TensorFlow-0.1x graph code; neither TensorFlow nor python2 exists in this image
and the repository ships no tests, golden vectors or fixtures.
The arithmetic lives in an un-vendored dependency (TensorFlow ~0.11/0.12 +
tf.contrib.slim, version not pinned by the reference).  Every function here
cites the reference file:line it follows and re-states the published TF op
semantics by hand; it is protected by analytic known-answer tests
(tests/test_oracle_*.py) and by two independent restatements of the warp
(vectorised torch in loss_interp.py vs. a literal per-pixel loop in
warp_literal.py).
"""
from . import tf_ops, loss_interp, flownet_s, adam, metrics  # noqa: F401
