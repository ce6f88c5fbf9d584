"""The oracle's dense ops against their literal-loop second restatements (oracle/literal.py): TF-SAME conv with every asymmetric padding the
net uses, conv2d_transpose, and the FlowNetC correlation (which has NO reference symbol: two independent restatements are all there is)."""
import numpy as np
import pytest
import torch

from oracle import tf_ops, literal, flownet_c


@pytest.mark.parametrize("H,W,ci,co,k,s", [(7, 9, 3, 4, 3, 1), (8, 10, 2, 5, 5, 2), (9, 7, 3, 2, 7, 2), (6, 6, 4, 3, 3, 2), (12, 16, 6, 4, 7, 2),
                                           (5, 5, 2, 2, 1, 1)])
def test_same_conv(H, W, ci, co, k, s):
    rng = np.random.RandomState(H * W + k)
    x, w, b = rng.randn(2, H, W, ci), rng.randn(k, k, ci, co), rng.randn(co)
    got = tf_ops.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), s).numpy()
    assert np.abs(got - literal.conv2d_same_literal(x, w, b, s)).max() < 1e-12
    oh, pb, pa = tf_ops.same_pad(H, k, s)
    assert got.shape[1] == oh == -(-H // s) and pb <= pa                     # TF puts the extra padding AFTER


@pytest.mark.parametrize("h,w,ci,co", [(3, 4, 5, 2), (5, 3, 2, 3), (6, 8, 2, 2)])
def test_transposed_conv(h, w, ci, co):
    rng = np.random.RandomState(h * w + ci)
    x, wt, b = rng.randn(2, h, w, ci), rng.randn(4, 4, co, ci), rng.randn(co)
    got = tf_ops.conv2d_transpose_same(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(b), 2).numpy()
    assert got.shape == (2, 2 * h, 2 * w, co)
    assert np.abs(got - literal.conv2d_transpose_same_literal(x, wt, b, 2)).max() < 1e-12


@pytest.mark.parametrize("h,w,md,s2", [(6, 9, 4, 2), (5, 7, 3, 1), (4, 12, 20, 2)])
def test_correlation(h, w, md, s2):
    rng = np.random.RandomState(h + w + md)
    f1, f2 = rng.randn(2, h, w, 8), rng.randn(2, h, w, 8)
    got = flownet_c.correlation(torch.from_numpy(f1), torch.from_numpy(f2), md, s2).numpy()
    assert np.abs(got - literal.correlation_literal(f1, f2, md, s2)).max() < 1e-12
