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


# ---------------------------------------------------------------------------------------------------------------------
# Host-side pieces the reference holds as plain NumPy / cv2 code: utils.readFlow / flow_ee, train.load_deconv_weights, the evaluation
# recipe and flyingChairsLoader.hookTrainData.  tests/golden/make_reference_host_golden.py EXECUTES those reference lines (read at run
# time) and stores inputs + outputs in tests/golden/reference_host.npz.
# ---------------------------------------------------------------------------------------------------------------------
HOST = GOLDEN / "reference_host.npz"


def _host():
    return np.load(HOST)


def test_readflow_and_flow_ee_equal_reference_utils(tmp_path):
    from deepof_b200 import utils as U
    z = _host()
    for fid in z["ids"]:
        p = tmp_path / f"{fid}_flow.flo"
        p.write_bytes(z[f"flo_{fid}"].tobytes())
        got = U.readFlow(str(p))
        want = z[f"readflow_{fid}"]
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)
        # writeFlow -> readFlow round trip reproduces the reference parser's array (the reference's own writeFlow names an undefined TAG_CHAR)
        q = tmp_path / f"{fid}_rt.flo"
        U.writeFlow(str(q), want)
        assert q.read_bytes() == p.read_bytes()
    assert abs(float(U.flow_ee(z["ee_f1"], z["ee_f2"])) - float(z["ee_aee"])) < 1e-12
    from oracle import metrics
    got = float(metrics.flow_ee(torch.from_numpy(z["ee_f1"]).double(), torch.from_numpy(z["ee_f2"]).double()))
    assert abs(got - float(z["ee_aee"])) < 1e-6          # (the reference reduces in float32)


def test_bilinear_deconv_init_equals_reference_load_deconv_weights():
    """The bilinear filters the reference assigns to every 'up*' variable (flyingChairsTrain.py:78-92, executed with a stand-in session)
    against the engine's initialiser (deepof_b200.flownet.bilinear_deconv, also behind train.load_deconv_weights) and the oracle's."""
    from deepof_b200.flownet import bilinear_deconv
    from oracle import tf_ops
    z = _host()
    keys = [k for k in z.files if k.startswith("bilinear_")]
    assert len(keys) == 3
    for key in keys:
        shape = tuple(int(v) for v in key.split("_")[1].split("x"))
        want = z[key]
        assert want.shape == shape and want.max() == 0.5625
        assert np.abs(bilinear_deconv(shape).double().numpy() - want).max() < 1e-7
        assert np.abs(tf_ops.bilinear_deconv_weights(shape, dtype=torch.float64).numpy() - want).max() < 1e-12


def test_eval_recipe_equals_reference_lines():
    """flows_all[0] * 2 -> clip -> cv2.resize -> utils.flow_ee (flyingChairsTrain.py:263-267, 294-296) against the oracle's eval_flow / flow_ee."""
    from oracle import metrics
    z = _host()
    pr1, gt = torch.from_numpy(z["eval_pr1"]), torch.from_numpy(z["eval_gt"])
    up = metrics.eval_flow(pr1, gt.shape[1], gt.shape[2])
    assert np.abs(up.numpy() - z["eval_up"]).max() < 2e-4          # float32 bilinear weights, cv2 vs torch summation order
    assert abs(float(metrics.flow_ee(up, gt)) - float(z["eval_aee"])) < 1e-4


@pytest.mark.skipif(not Path("/root/reference/utils.py").exists(), reason="reference checkout not present (GPU box)")
def test_reference_host_fixture_is_what_the_reference_produces():
    spec = importlib.util.spec_from_file_location("make_reference_host_golden", GOLDEN / "make_reference_host_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh, z = mod.build(), _host()
    assert sorted(fresh) == sorted(z.files)
    for k in z.files:
        assert np.array_equal(np.asarray(fresh[k]), z[k]), k
