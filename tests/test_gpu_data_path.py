"""Device-side FlyingChairs data path: .ppm / .flo decode and the evaluation recipe against the reference's own
host tools -- cv2.imread / cv2.resize (flyingChairsLoader.py:70-78) and utils.readFlow / utils.flow_ee (utils.py:4-21,64-68)."""
import os

import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_dataset(root, n, hw=(48, 64), comment=False):
    from deepof_b200 import utils
    data = os.path.join(root, "data")
    os.makedirs(data, exist_ok=True)
    rng = np.random.RandomState(5)
    H, W = hw
    with open(os.path.join(root, "FlyingChairs_train_val.txt"), "w") as f:
        for i in range(n):
            f.write("2\n" if i % 4 == 3 else "1\n")
    flows = []
    for i in range(n):
        fid = "%05d" % (i + 1)
        for k in (1, 2):
            img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)       # RGB as stored in a P6 file
            with open(os.path.join(data, f"{fid}_img{k}.ppm"), "wb") as f:
                f.write(b"P6\n" + (b"# made by a test\n" if comment else b"") + b"%d %d\n255\n" % (W, H) + img.tobytes())
        fl = (rng.randn(H, W, 2) * 5).astype(np.float32)
        utils.writeFlow(os.path.join(data, fid + "_flow.flo"), fl)
        flows.append(fl)
    return data, flows


@pytest.mark.parametrize("image_size,comment", [((48, 64), False), ((40, 56), True), ((24, 32), False)])
def test_loader_decodes_like_cv2(tmp_path, image_size, comment):
    from deepof_b200.flyingChairsLoader import flyingChairsLoader
    from deepof_b200 import utils
    data, flows = _write_dataset(str(tmp_path), 8, comment=comment)
    ld = flyingChairsLoader(str(tmp_path), image_size, split_file=os.path.join(str(tmp_path), "FlyingChairs_train_val.txt"))
    assert len(ld.trainList) == 6 and len(ld.valList) == 2 and ld.trainList[0] == "00001" and ld.valList[0] == "00004"
    src, tgt, flow = ld.sampleTrain(3, 2)                                 # samples 3..5 of the train list
    assert src.is_cuda and tuple(src.shape) == (3, image_size[0], image_size[1], 3) and tuple(flow.shape) == (3, 48, 64, 2)
    for j, fid in enumerate(ld.trainList[3:6]):
        for k, got in ((1, src), (2, tgt)):
            ref = cv2.imread(os.path.join(data, f"{fid}_img{k}.ppm"), cv2.IMREAD_COLOR)
            ref = cv2.resize(ref, (image_size[1], image_size[0]))
            assert np.array_equal(got[j].cpu().numpy(), ref.astype(np.float32)), (fid, k)      # bit exact (8-bit fixed-point bilinear)
        assert np.array_equal(flow[j].cpu().numpy(), utils.readFlow(os.path.join(data, fid + "_flow.flo")))
    (vs, vt, vf), idxs = ld.sampleVal(2, 1)
    assert list(idxs) == [0, 1] and tuple(vs.shape) == (2, image_size[0], image_size[1], 3)


def test_bad_files_are_loud(tmp_path):
    from deepof_b200.flyingChairsLoader import flyingChairsLoader
    from deepof_b200 import DeepOFError
    data, _ = _write_dataset(str(tmp_path), 4)
    ld = flyingChairsLoader(str(tmp_path), (48, 64), split_file=os.path.join(str(tmp_path), "FlyingChairs_train_val.txt"))
    with open(os.path.join(data, "00001_flow.flo"), "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(DeepOFError, match="Magic number"):
        ld.sampleTrain(2, 1)
    with open(os.path.join(data, "00002_img1.ppm"), "r+b") as f:
        f.write(b"P5")
    with pytest.raises(DeepOFError, match="P6"):
        ld.hookTrainData([1])


@pytest.mark.parametrize("hw,HW", [((24, 32), (48, 64)), ((20, 28), (48, 64)), ((192, 256), (384, 512))])
def test_eval_recipe_matches_numpy_cv2(hw, HW):
    """flows_all[0] * 2 -> clip -> cv2.resize -> utils.flow_ee (flyingChairsTrain.py:264-266,294-296)."""
    from deepof_b200.flyingChairsLoader import evaluate_aee
    from deepof_b200 import utils
    g = torch.Generator().manual_seed(hw[0])
    B = 3
    flow1 = torch.randn(B, hw[0], hw[1], 2, generator=g) * 80.0          # some values beyond the clip range
    gt = torch.randn(B, HW[0], HW[1], 2, generator=g) * 20.0
    ups = []
    for b in range(B):
        fi = np.clip(flow1[b].numpy() * 2, -300.0, 250.0)
        ups.append(cv2.resize(fi, (HW[1], HW[0]))[None])
    want = utils.flow_ee(np.concatenate(ups, axis=0), gt.numpy())
    got = evaluate_aee(flow1.cuda(), gt.cuda())
    assert abs(got - want) < 1e-5 * max(1.0, want)
    assert abs(utils.flow_ee(torch.from_numpy(np.concatenate(ups, axis=0)).cuda(), gt.cuda()) - want) < 1e-5 * max(1.0, want)


def test_train_class_runs_from_a_dataset_directory(tmp_path):
    """deepOF_fc.deepOF(data_path) end to end on a miniature data set: loader -> pre-scaling -> 4-feed VGG16 step (deepOF_fc.py:5-7)."""
    from deepof_b200 import deepOF_fc
    _write_dataset(str(tmp_path), 8, hw=(128, 160))
    os.chdir(str(tmp_path))                               # the reference looks the split file up in the working directory
    deepOF_fc.IMAGE_SIZE[:] = [128, 160]     # (at 64 x 96 the 2 x 3 map of scale 5 has an empty border mask: NaN by the reference's own formula)
    try:
        t = deepOF_fc.deepOF(str(tmp_path), batch_size=2, max_iters=2, math_mode="fp32", display=1)
    finally:
        deepOF_fc.IMAGE_SIZE[:] = [320, 448]
    assert t.step.engine.t == 2 and np.isfinite(t.step.last_loss())


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode", ["fp32", "bf16"])
def test_uint8_feeds_equal_their_float32_casts(math_mode, monkeypatch):
    """TrainStep.run / fetch with uint8 images (what flyingChairsLoader.hookTrainData returns) == the same images fed as float32: the cast
    happens in the pre-processing kernel and (float)u8 is exact -> identical network inputs, pyramids, losses and flows."""
    import torch
    from deepof_b200 import ops
    from deepof_b200.flyingChairsTrain import TrainStep, WEIGHT_L
    monkeypatch.setenv("DOFB_SPLITK", "0")        # bit-for-bit comparison of two engines: no atomically summed K ranges
    B, H, W = 2, 192, 256                          # (smaller maps have an empty border mask at scale 6: NaN losses)
    g = torch.Generator().manual_seed(3)
    src8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    tgt8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    outs = []
    for feed in ((src8.numpy(), tgt8.numpy()), (src8.float().numpy(), tgt8.float().numpy())):
        step = TrainStep(B, (H, W), math_mode=math_mode, seed=1)
        losses, flows, total = step.fetch({"source_img": feed[0], "target_img": feed[1], "loss_weight": WEIGHT_L})
        step.run({"source_img": feed[0], "target_img": feed[1], "loss_weight": WEIGHT_L, "learning_rate": 1e-5})
        outs.append((total, flows, step.last_loss(), step.engine.pyr_src[1].clone(), step.engine.pyr_tgt[2].clone()))
    a, b = outs
    assert np.isfinite(a[0]) and a[0] == b[0] and a[2] == b[2]
    for fa, fb in zip(a[1], b[1]):
        assert (fa == fb).all()
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    # odd width: scalar path of the kernel
    s8 = torch.randint(0, 256, (1, 8, 10, 3), generator=g, dtype=torch.uint8).cuda()
    t8 = torch.randint(0, 256, (1, 8, 10, 3), generator=g, dtype=torch.uint8).cuda()
    xa, xb = torch.zeros(1, 8, 10, 8, device="cuda"), torch.zeros(1, 8, 10, 8, device="cuda")
    pa = [torch.zeros(1, 4, 5, 3, device="cuda") for _ in range(2)]
    pb = [torch.zeros(1, 4, 5, 3, device="cuda") for _ in range(2)]
    ops.preprocess(s8, t8, (100.0, 110.0, 120.0), xa, pa[:1], pa[1:])
    ops.preprocess(s8.float(), t8.float(), (100.0, 110.0, 120.0), xb, pb[:1], pb[1:])
    torch.cuda.synchronize()
    assert torch.equal(xa, xb) and torch.equal(pa[0], pb[0]) and torch.equal(pa[1], pb[1])


def _ref_host():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "reference_host.npz")


def test_device_loader_equals_reference_hook_train_data(tmp_path):
    """flyingChairsLoader on the device (file bytes -> dofb_decode_ppm incl. the cv2.resize restatement, dofb_decode_flo) against what the
    REFERENCE's hookTrainData (flyingChairsLoader.py:64-82, executed by tests/golden/make_reference_host_golden.py) returned for the same files."""
    from deepof_b200.flyingChairsLoader import flyingChairsLoader
    z = _ref_host()
    ids = [str(i) for i in z["ids"]]
    data = tmp_path / "data"
    data.mkdir()
    for fid in ids:
        for k in (1, 2):
            (data / f"{fid}_img{k}.ppm").write_bytes(z[f"ppm_{fid}_{k}"].tobytes())
        (data / f"{fid}_flow.flo").write_bytes(z[f"flo_{fid}"].tobytes())
    (tmp_path / "FlyingChairs_train_val.txt").write_text("1\n" * len(ids))
    ld = flyingChairsLoader(str(tmp_path), [32, 48], split_file=str(tmp_path / "FlyingChairs_train_val.txt"))
    src, tgt, flow = ld.sampleTrain(len(ids), 1)
    torch.cuda.synchronize()
    assert torch.equal(src.cpu(), torch.from_numpy(z["loader_source"].astype(np.float32)))      # bit-exact cv2.resize (shrinking), BGR order
    assert torch.equal(tgt.cpu(), torch.from_numpy(z["loader_target"].astype(np.float32)))
    assert torch.equal(flow.cpu(), torch.from_numpy(z["loader_flow"]))


def test_device_eval_recipe_and_epe_equal_reference_lines():
    """dofb_eval_flow_aee_sum (x2, clip, cv2.resize, AEE in one pass) and dofb_epe_sum against the numbers the reference's own lines
    (flyingChairsTrain.py:263-267,294-296; utils.py:64-68) produced."""
    from deepof_b200.flyingChairsLoader import evaluate_aee
    from deepof_b200 import utils as U
    z = _ref_host()
    got = evaluate_aee(torch.from_numpy(z["eval_pr1"]).cuda(), torch.from_numpy(z["eval_gt"]).cuda())
    assert abs(got - float(z["eval_aee"])) < 2e-5 * max(1.0, float(z["eval_aee"]))
    ee = U.flow_ee(torch.from_numpy(z["ee_f1"]).cuda(), torch.from_numpy(z["ee_f2"]).cuda())
    assert abs(ee - float(z["ee_aee"])) < 1e-5 * float(z["ee_aee"])
