"""CPU pins of the folded tangent, against the oracle's restatement of the
reference's modulation (oracle/net.py::modulate <- style_layers_vel.py:62-105).

The kernels replace  x*dW + dx*W  by  (dx + a (.) x)*W + beta (.) (x*W).  That is only legitimate if the
reference's tangent weights really have the structure  dW = W (.) (a_i + beta_o)  -- checked here in fp64
for every layer kind, including the first-layer + W/Dz term -- and if the host-side factorisation used for
premodulated trees (csrc/nbe_api.cu::factor_premod, restated below in numpy) recovers a and beta from
(W, dW) alone.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle.net import init_params, layer_table, modulate

P = init_params(42)
S = torch.tensor([[(0.25 - 0.3) * 5.0, 0.61 - 1.0]], dtype=torch.float64)      # (Om, Dz) = (0.25, 0.61)


def fold_vectors(lp, s, first, eps=1e-8):
    """a_i and beta_o in closed form from the style parameters (what modulate_kernel computes)."""
    W = torch.as_tensor(lp["weight"]).double()
    SW = torch.as_tensor(lp["style_weight"]).double()
    sb = torch.as_tensor(lp["style_bias"]).double()
    m = s[0] @ SW.T + sb                                           # (I,)
    w = W * m[None, :, None, None, None]
    dws = W * SW[:, 1][None, :, None, None, None]
    n2 = (w * w).sum(dim=(1, 2, 3, 4)) + eps
    a = SW[:, 1] / m
    beta = -(w * dws).sum(dim=(1, 2, 3, 4)) / n2
    if first:
        beta = beta + 1.0 / (s[0, 1] + 1.0)
    return a, beta


def factor_premod(W, dW):
    """numpy restatement of csrc/nbe_api.cu::factor_premod."""
    num = (dW * W).sum(axis=(2, 3, 4))
    den = (W * W).sum(axis=(2, 3, 4))
    R = num / den
    a = R.mean(axis=0)
    beta = R.mean(axis=1) - R.mean()
    resid = np.abs(dW - (a[None, :, None, None, None] + beta[:, None, None, None, None]) * W).max()
    return a, beta, resid / np.abs(dW).max()


def test_reference_tangent_weights_have_the_fold_structure():
    checked = 0
    for block, layer, co, ci, k in layer_table():
        lp = P["params"][block][layer]
        first = block == "conv_l00" and layer in ("conv_0", "skip")
        wn, dwn = modulate(lp, S, first, True, dtype=torch.float64)
        a, beta = fold_vectors(lp, S, first)
        want = wn[0] * (a[None, :, None, None, None] + beta[:, None, None, None, None])
        err = float((dwn[0] - want).abs().max() / dwn[0].abs().max())
        assert err < 1e-12, (block, layer, err)
        checked += 1
    assert checked == 33


def test_folded_layer_equals_the_reference_tangent():
    """x*dW + dx*W == (dx + a x)*W + beta (x*W) on a 3^3 conv, and the residual form for a second reader of the
    same tensor (a 1^3 skip whose own a differs from the fold vector stored with the tensor)."""
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.standard_normal((1, 64, 7, 8, 9)))
    dx = torch.as_tensor(rng.standard_normal((1, 64, 7, 8, 9)))
    lp, ls = P["params"]["conv_l01"]["conv_0"], P["params"]["conv_l01"]["skip"]
    wn, dwn = modulate(lp, S, False, True, dtype=torch.float64)
    a, beta = fold_vectors(lp, S, False)
    ref = F.conv3d(x, dwn[0]) + F.conv3d(dx, wn[0])
    dxp = dx + a[None, :, None, None, None] * x                    # what the producer stores
    y = F.conv3d(x, wn[0])
    got = F.conv3d(dxp, wn[0]) + beta[None, :, None, None, None] * y
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-12
    # the skip conv of the same block reads the same tensor: its tangent rows are dW_s - a (.) W_s
    ws, dws = modulate(ls, S, False, True, dtype=torch.float64)
    ref_s = F.conv3d(x, dws[0]) + F.conv3d(dx, ws[0])
    res = dws[0] - a[None, :, None, None, None] * ws[0]
    got_s = F.conv3d(x, res) + F.conv3d(dxp, ws[0])
    assert float((got_s - ref_s).abs().max() / ref_s.abs().max()) < 1e-12


def test_premodulated_factorisation_recovers_a_and_beta_up_to_the_gauge():
    lp = P["params"]["conv_r00"]["conv_0"]                          # 128 -> 128
    wn, dwn = modulate(lp, S, False, True, dtype=torch.float32)
    W, dW = wn[0].numpy().astype(np.float64), dwn[0].numpy().astype(np.float64)
    a, beta, resid = factor_premod(W, dW)
    assert resid < 1e-4                                             # the acceptance test of nbe_set_params
    a0, b0 = fold_vectors(lp, S, False)
    shift = float(np.mean(a - a0.numpy()))                          # (a + c, beta - c) is the same fold
    assert np.abs(a - a0.numpy() - shift).max() < 1e-4 and np.abs(beta - b0.numpy() + shift).max() < 1e-4
    # a dweight without the structure is rejected
    bad = dW + 0.05 * np.abs(dW).max() * np.random.default_rng(1).standard_normal(dW.shape)
    assert factor_premod(W, bad)[2] > 1e-2
