"""Precision evidence for the tensor-core math modes.

* forward flows of every mode against the fp32 CPU ORACLE at all six scales (per-pixel end-point distance, stated tolerances of
  deepof_b200/precision.py), on init weights and on weights after a few hundred optimiser steps;
* training trajectories: fp32, tf32 and bf16 engines trained from the same initialisation on the same batches stay together."""
import pytest
import torch

from oracle import flownet_s as ofs

pytestmark = pytest.mark.gpu
H, W, B = 192, 256, 4


def _batches(n, seed0=100):
    from deepof_b200.synth import make_pairs
    out = []
    for i in range(n):
        s, t, _ = make_pairs(B, H, W, seed=seed0 + i)
        out.append((s.cuda(), t.cuda()))
    return out


@pytest.fixture(scope="module")
def trained_params():
    """Weights after 300 bf16 training steps at a learning rate 10x the reference's (non-trivial flows in a short run)."""
    from deepof_b200.flownet import FlowNetS
    eng = FlowNetS(B, H, W, math_mode="bf16", seed=1, tc_wgrad=True)
    data = _batches(4)
    for i in range(300):
        eng.train_step(*data[i % 4], lr=1.6e-4)
    torch.cuda.synchronize()
    return eng.export_params()


@pytest.mark.parametrize("which", ["init", "trained"])
@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_forward_flows_against_the_cpu_oracle(mode, which, trained_params):
    from deepof_b200.flownet import FlowNetS, FLOW_SCALES
    from deepof_b200 import precision
    from deepof_b200.synth import make_pairs
    params = ofs.init_params(1) if which == "init" else trained_params
    src, tgt, _ = make_pairs(2, H, W, seed=1234)
    with torch.no_grad():
        _l, flows_ref, _p, _t = ofs.forward(params, src, tgt)
    eng = FlowNetS(2, H, W, math_mode=mode, seed=None, tc_wgrad=mode != "fp32")
    eng.load_params(params)
    eng.forward(src.cuda(), tgt.cuda(), with_grad=False)
    flows = [eng.pr[s] * FLOW_SCALES[s] for s in range(1, 7)]
    stats = precision.end_point_distance(flows, flows_ref)
    mag = precision.flow_magnitude(flows_ref)
    assert precision.within(stats, mode), (mode, which, stats, mag)
    if which == "trained":
        assert max(mag) > 0.05          # the trained network predicts non-trivial flows: the comparison is not 0 vs 0


def test_training_trajectories_of_the_math_modes_stay_together():
    """50 steps at the reference's learning rate from the same initialisation on the same batches: the loss trajectories of tf32 / bf16
    track fp32 (measured max relative gap: tf32 2.8e-3, bf16 1.2e-3) and all three descend by the same amount."""
    from deepof_b200.flownet import FlowNetS
    data = _batches(5, seed0=300)
    traj = {}
    for mode in ("fp32", "tf32", "bf16"):
        eng = FlowNetS(B, H, W, math_mode=mode, seed=1, tc_wgrad=mode != "fp32")
        losses = []
        for i in range(50):
            eng.train_step(*data[i % 5], lr=1.6e-5)
            losses.append(eng.total_loss().reshape(1).clone())
        traj[mode] = torch.cat(losses).cpu().double()
    ref = traj["fp32"]
    assert float(ref[-5:].mean()) < float(ref[:5].mean())                    # it trains
    for mode in ("tf32", "bf16"):
        gap = ((traj[mode] - ref).abs() / ref.abs()).max()
        assert float(gap) < 8e-3, (mode, float(gap))
        d_ref, d = float(ref[:5].mean() - ref[-5:].mean()), float(traj[mode][:5].mean() - traj[mode][-5:].mean())
        assert abs(d - d_ref) < 0.05 * abs(d_ref), (mode, d, d_ref)
