"""Synthetic FlyingChairs-shaped inputs: re-export of deepof_b200.synth so that tests
written against the oracle and the product draw the same batches."""
from deepof_b200.synth import make_pairs  # noqa: F401
