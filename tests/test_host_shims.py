"""Host-side drop-in surface: reference module names, signatures, schedules and file parsing -- no GPU needed."""
import importlib
import inspect
import struct
import sys

import numpy as np
import pytest


def test_reference_module_names_resolve_through_the_compat_path(monkeypatch):
    import deepof_b200.compat as compat
    monkeypatch.syspath_prepend(compat.PATH)
    for name in compat.MODULES:
        sys.modules.pop(name, None)
        m = importlib.import_module(name)
        assert m.__file__.startswith(compat.PATH)
    import flyingChairsWrapFlow, flyingChairsWrapFlow_vgg, flyingChairsTrain, flyingChairsTrain_vgg, deepOF_fc, warpflow, Flownet, utils  # noqa: E401
    import sintelWrapFlow, flyingChairsLoader
    # the call sites the reference scripts use
    assert list(inspect.signature(flyingChairsWrapFlow.flowNet).parameters)[:3] == ["inputs", "outputs", "loss_weight"]          # flyingChairsWrapFlow.py:5
    assert list(inspect.signature(flyingChairsWrapFlow.loss_interp).parameters) == [
        "flows", "inputs", "outputs", "epsilon", "alpha_c", "alpha_s", "lambda_smooth", "flow_scale", "deltaWeights"]            # :752
    assert list(inspect.signature(warpflow.loss_interp).parameters) == list(inspect.signature(flyingChairsWrapFlow.loss_interp).parameters)
    assert list(inspect.signature(sintelWrapFlow.loss_interp_multi).parameters) == [
        "flows", "inputs", "epsilon", "alpha_c", "alpha_s", "lambda_smooth", "flow_scale", "deltaWeights"]                      # sintelWrapFlow.py:492
    assert list(inspect.signature(flyingChairsWrapFlow_vgg.VGG16).parameters)[:5] == [
        "photo_source", "photo_target", "geo_source", "geo_target", "loss_weight"]                                             # _vgg.py:7
    assert list(inspect.signature(Flownet.model).parameters)[:6] == [
        "source_imgs", "target_imgs", "sample_mean", "loss_weight", "hyper_params", "is_training"]                              # Flownet.py:22
    assert list(inspect.signature(deepOF_fc.deepOF).parameters)[0] == "data_path" and deepOF_fc.IMAGE_SIZE == [320, 448]        # deepOF_fc.py:5-6
    assert hasattr(flyingChairsTrain.train, "trainNet") and hasattr(flyingChairsTrain.train, "load_deconv_weights")
    assert flyingChairsTrain_vgg.WEIGHT_L == [16, 8, 4, 2, 1] and flyingChairsTrain.WEIGHT_L == [16, 8, 4, 2, 1, 1]             # :171 / :165
    assert flyingChairsTrain_vgg.TrainStep.FEEDS == ("photo_source", "photo_target", "geo_source", "geo_target")
    assert callable(utils.readFlow) and callable(utils.writeFlow) and callable(utils.flow_ee) and callable(flyingChairsLoader.flyingChairsLoader)


def test_learning_rate_schedule():
    from deepof_b200.flyingChairsTrain import learning_rate_at, LEARNING_RATE
    assert LEARNING_RATE == 0.000016
    # flyingChairsTrain.py:208-209: `if epoch % 18 == 0: lr *= 0.5` AFTER the epoch -> epochs 1..18 at lr, 19..36 at lr/2, ...
    lr, want = LEARNING_RATE, []
    for epoch in range(1, 60):
        want.append(lr)
        if epoch % 18 == 0:
            lr *= 0.5
    assert [learning_rate_at(e) for e in range(1, 60)] == want


def test_hyper_parameter_orders():
    from deepof_b200 import Flownet
    assert Flownet.hyper_from_list([1.0, 1e-4, 0.25, 0.37]) == dict(lambda_smooth=1.0, epsilon=1e-4, alpha_c=0.25, alpha_s=0.37)   # Flownet.py:65-68
    assert Flownet.hyper_from_sintel_list([1e-4, 0.3, 0.3, 0]) == dict(lambda_smooth=0.0, epsilon=1e-4, alpha_c=0.3, alpha_s=0.3)  # sintelTrain.py:181
    assert Flownet.LOSS_WEIGHT_V1 == [9, 7, 5, 3, 3, 1]


def test_bilinear_deconv_init_is_the_reference_formula():
    import torch
    from deepof_b200.flyingChairsTrain import load_deconv_weights
    # flyingChairsTrain.py:78-92 evaluated by hand for k = 4 (py2: f = ceil(4/2.0) = 2.0, c = (2*2 - 1 - 0) / 4 = 0.75)
    ax = np.array([1 - abs(i / 2.0 - 0.75) for i in range(4)])
    w = torch.full((4, 4, 3, 3), 7.0)
    load_deconv_weights.__wrapped__(w) if hasattr(load_deconv_weights, "__wrapped__") else None
    from deepof_b200.flownet import bilinear_deconv
    got = bilinear_deconv((4, 4, 3, 3)).numpy()
    for i in range(3):
        for j in range(3):
            assert np.allclose(got[:, :, i, j], np.outer(ax, ax) if i == j else 0.0)
    assert np.allclose(ax, [0.25, 0.75, 0.75, 0.25])


def test_ppm_header_and_flo_io(tmp_path):
    from deepof_b200.flyingChairsLoader import _ppm_header
    from deepof_b200 import utils, DeepOFError
    assert _ppm_header(memoryview(b"P6\n64 48\n255\n" + bytes(10))) == (64, 48, 255, 13)
    assert _ppm_header(memoryview(b"P6 # a comment\n# another\n512 384 255\n" + bytes(10)))[:3] == (512, 384, 255)
    with pytest.raises(DeepOFError):
        _ppm_header(memoryview(b"P5\n4 4\n255\n"))
    with pytest.raises(DeepOFError):
        _ppm_header(memoryview(b"P6\n4 4\n65535\n"))
    fl = (np.random.RandomState(0).randn(5, 7, 2) * 3).astype(np.float32)
    fn = str(tmp_path / "a.flo")
    utils.writeFlow(fn, fl)
    raw = open(fn, "rb").read()
    assert raw[:4] == b"PIEH" and struct.unpack("<f", raw[:4])[0] == 202021.25 and struct.unpack("<ii", raw[4:12]) == (7, 5)   # utils.py:12
    assert np.array_equal(utils.readFlow(fn), fl)
    open(fn, "r+b").write(b"XXXX")
    assert utils.readFlow(fn) is None                                                                      # :13-14
    f1, f2 = np.zeros((2, 3, 4, 2), np.float32), np.ones((2, 3, 4, 2), np.float32)
    assert abs(utils.flow_ee(f1, f2) - np.sqrt(2.0)) < 1e-7                                                # :64-68
