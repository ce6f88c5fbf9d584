"""FlowNetC (siamese tower + correlation cost volume) on the device vs its oracle (oracle/flownet_c.py).

The reference has no FlowNetC: this pins the CUDA path to OUR specification of it."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flownet_c as fc, synth, metrics  # noqa: E402

KEYS = ("total", "Charbonnier_reconstruct", "U_loss", "V_loss")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def case():
    from deepof_b200.flownet import FlowNetC
    B, H, W = 2, 192, 256
    src, tgt, gt = synth.make_pairs(B, H, W, seed=9)
    params = fc.init_params(seed=1)
    total, grads, losses, flows_all, prev1 = fc.loss_and_grads(params, src, tgt)
    _t, grads64, *_ = fc.loss_and_grads({k: v.double() for k, v in params.items()}, src.double(), tgt.double())
    eng = FlowNetC(B, H, W, seed=None)
    eng.load_params(params)
    eng.forward(src.cuda(), tgt.cuda(), fc.LOSS_WEIGHTS, with_grad=True)
    eng.backward()
    torch.cuda.synchronize()
    return dict(eng=eng, params=params, total=total, grads=grads, grads64=grads64, losses=losses, flows_all=flows_all, prev1=prev1,
                src=src, tgt=tgt, gt=gt)


def test_flownetc_forward(case):
    eng = case["eng"]
    want = torch.tensor([[l[k].item() for k in KEYS] for l in case["losses"]])
    assert torch.allclose(eng.loss4.cpu(), want, rtol=5e-5, atol=1e-6), (eng.loss4.cpu(), want)
    _l, flows_all, prev1 = eng.outputs()
    for s in range(6):
        assert (flows_all[s].cpu() - case["flows_all"][s].detach()).abs().max() < 2e-4, s
    assert (prev1.cpu() - case["prev1"].detach()).abs().max() < 1e-4


def test_flownetc_gradients(case):
    eng = case["eng"]
    assert len(eng.grads) == 54
    for name, g32 in case["grads"].items():
        g64 = case["grads64"][name]
        e_dev, e_cpu = rel(eng.grads[name], g64), rel(g32, g64)
        assert e_dev < 3.0 * e_cpu + 1e-3, (name, e_dev, e_cpu)       # same bar as FlowNetS (ill-conditioned Charbonnier loss)


def test_flownetc_tf32_matches_fp32(case):
    from deepof_b200.flownet import FlowNetC
    B, H, W = 2, 192, 256
    etf = FlowNetC(B, H, W, seed=None, math_mode="tf32", tc_wgrad=True)
    etf.load_params(case["params"])
    etf.forward(case["src"].cuda(), case["tgt"].cuda())
    etf.backward()
    e32 = case["eng"]
    d = (etf.pr[1] - e32.pr[1]).abs() * 10.0
    epe32 = metrics.flow_ee(metrics.eval_flow(e32.pr[1].cpu() * 10.0, H, W), case["gt"]).item()
    epetf = metrics.flow_ee(metrics.eval_flow(etf.pr[1].cpu() * 10.0, H, W), case["gt"]).item()
    print(f"FlowNetC tf32 vs fp32: mean|dflow1|={d.mean().item():.3e} max={d.max().item():.3e} EPE {epe32:.6f} / {epetf:.6f}")
    assert abs(epetf - epe32) < 1e-3
    assert torch.allclose(etf.loss4, e32.loss4, rtol=5e-3, atol=1e-4)
    cos = torch.nn.functional.cosine_similarity(etf.grad.double(), e32.grad.double(), dim=0).item()
    print("gradient cosine tf32 vs fp32:", cos)
    assert cos > 0.9          # (FlowNetS reaches > 0.98; the correlation layer adds a second ill-conditioned stage)


def test_flownetc_train_steps_reduce_loss():
    from deepof_b200.flownet import FlowNetC
    src, tgt, _ = synth.make_pairs(1, 192, 256, seed=4)
    eng = FlowNetC(1, 192, 256, seed=1, math_mode="tf32", tc_wgrad=True)
    eng.forward(src.cuda(), tgt.cuda(), with_grad=False)
    l0 = eng.total_loss().item()
    for _ in range(3):
        eng.train_step(src.cuda(), tgt.cuda(), lr=1.6e-5)
    eng.forward(src.cuda(), tgt.cuda(), with_grad=False)
    assert eng.total_loss().item() < l0


def test_flownetc_bf16_tracks_fp32(case):
    """bf16 tensor-core mode (bf16 shadows, CTA pairs, merged stride phases, batched packs, bf16 siamese pre-processing) on FlowNetC."""
    from deepof_b200.flownet import FlowNetC
    B, H, W = 2, 192, 256
    eb = FlowNetC(B, H, W, seed=None, math_mode="bf16", tc_wgrad=True)
    eb.load_params(case["params"])
    eb.forward(case["src"].cuda(), case["tgt"].cuda())
    eb.backward()
    e32 = case["eng"]
    epe32 = metrics.flow_ee(metrics.eval_flow(e32.pr[1].cpu() * 10.0, H, W), case["gt"]).item()
    epeb = metrics.flow_ee(metrics.eval_flow(eb.pr[1].cpu() * 10.0, H, W), case["gt"]).item()
    print(f"FlowNetC bf16 vs fp32: EPE {epe32:.6f} / {epeb:.6f}")
    assert abs(epeb - epe32) < 1e-3                                   # north_star tolerance
    assert torch.allclose(eb.loss4, e32.loss4, rtol=2e-2, atol=1e-3)
    assert torch.isfinite(eb.grad).all()
    cos = torch.nn.functional.cosine_similarity(eb.grad.double(), e32.grad.double(), dim=0).item()
    print("gradient cosine bf16 vs fp32:", cos)
    assert cos > 0.8
    for _ in range(2):
        eb.train_step(case["src"].cuda(), case["tgt"].cuda(), lr=1.6e-5)
    assert torch.isfinite(eb.theta).all()
