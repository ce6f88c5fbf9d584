"""Pins the oracle's warp to reference-held code: /root/reference/check_loss.py:61-135 (the author's NumPy restatement of the warp).

The fixture tests/golden/check_loss_warp.npz is produced by tests/golden/make_check_loss_golden.py, which EXECUTES those reference lines
(read from the checkout at run time, not copied).  On interior pixels (all four bilinear corners inside the image) the script's
flat-index clamp and the TF graph's per-axis clamp (flyingChairsWrapFlow.py:815-818) coincide, so there the oracle must agree to
float64 rounding.  The out-of-image behaviour (per-axis clamp) is covered by the analytic tests in test_oracle_known_answers.py."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import loss_interp as li
from oracle.warp_literal import warp_literal

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ("s6_6x8", "s4_24x32", "s3_48x64")


def _load(name):
    z = np.load(GOLDEN / "check_loss_warp.npz")
    flow, target, recon = z[name + "_flow"], z[name + "_target"], z[name + "_recon"]
    return flow, target, recon.reshape(target.shape)


@pytest.mark.parametrize("name", CASES)
def test_oracle_warp_equals_reference_check_loss(name):
    flow, target, want = _load(name)
    got = li.warp(torch.from_numpy(flow)[None], torch.from_numpy(target)[None])[0].numpy()
    assert got.dtype == np.float64
    assert np.abs(got - want).max() < 1e-13
    got32 = li.warp(torch.from_numpy(flow).float()[None], torch.from_numpy(target).float()[None])[0].numpy()
    assert np.abs(got32 - want).max() < 2e-6


@pytest.mark.parametrize("name", CASES[:2])
def test_literal_loop_equals_reference_check_loss(name):
    flow, target, want = _load(name)
    got = warp_literal(flow[None].astype(np.float32), target[None].astype(np.float32))[0]
    assert np.abs(got - want).max() < 2e-6


def test_loss_interp_reconstruction_is_the_reference_warp():
    """The full op (oracle.loss_interp, both variants) returns that same reconstruction: flow_scale multiplies first (:783)."""
    flow, target, want = _load("s4_24x32")
    src = torch.from_numpy(target).flip(1)[None]
    for variant in ("A", "B"):
        _ld, recon = li.loss_interp(torch.from_numpy(flow)[None] / 1.25, src, torch.from_numpy(target)[None], 1e-4, 0.25, 0.37, 1.0, 1.25,
                                    variant=variant)
        assert np.abs(recon[0].numpy() - want).max() < 1e-12


@pytest.mark.skipif(not Path("/root/reference/check_loss.py").exists(), reason="reference checkout not present (GPU box)")
def test_fixture_is_what_the_reference_code_produces():
    spec = importlib.util.spec_from_file_location("make_check_loss_golden", GOLDEN / "make_check_loss_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    src = mod.reference_warp_source()
    assert "np.clip(idx_a, 0, height*width-1)" in src and "xrange(channels)" in src          # really the reference's lines
    for name, flow, target in mod.cases():
        z_flow, z_target, z_recon = _load(name)
        assert np.array_equal(flow, z_flow) and np.array_equal(target, z_target)
        assert np.array_equal(mod.run_reference_warp(flow, target).reshape(target.shape), z_recon)
