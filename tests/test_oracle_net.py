"""Pin the network oracle (oracle/net.py) against what the reference's tests assert:
LeakyReLU known answers, shapes, the algebra of the velocity branch, Style == premodulated,
up-sampling semantics, and a finite-difference check of the manual JVP.  Also checks the
committed golden fixtures were produced by this oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import cosmology as oc
from oracle.net import Net, init_params, layer_table, modulate, modulate_emulator_parameters, rel_l2, _up_fast, _up_literal

torch.set_num_threads(max(1, os.cpu_count() or 1))
P = init_params(42)


def test_leaky_relu_known_answers():       # tests/test_layers.py:147-171, test_layers_vel.py:268-334
    net = Net()
    x = torch.tensor([-2.0, -1.0, 0.0, 1.0, 2.0], dtype=torch.float64)
    dx = torch.ones_like(x)
    y, dy = net.act(x, dx)
    assert torch.allclose(y, torch.tensor([-0.02, -0.01, 0.0, 1.0, 2.0], dtype=torch.float64))
    # tangent at x == 0 takes the slope branch (strict x > 0)
    assert torch.allclose(dy, torch.tensor([0.01, 0.01, 0.01, 1.0, 1.0], dtype=torch.float64))


def test_layer_table_matches_reference_tree():   # tests/test_style_nbody_emulator_vel_core.py:391-446
    rows = layer_table()
    assert len(rows) == 33
    blocks = [r[0] for r in rows]
    assert sorted(set(blocks)) == sorted(["conv_l00", "conv_l01", "down_l0", "conv_l1", "down_l1", "conv_l2",
                                          "down_l2", "conv_c", "up_r2", "conv_r2", "up_r1", "conv_r1", "up_r0",
                                          "conv_r00", "conv_r01"])
    n = sum(co * ci * k ** 3 + co + 2 * ci + ci for _, _, co, ci, k in rows)
    assert n == 3354776
    d = {(b, l): (co, ci, k) for b, l, co, ci, k in rows}
    assert d[("conv_l00", "conv_0")] == (64, 3, 3) and d[("conv_l00", "skip")] == (64, 3, 1)
    assert d[("conv_r2", "conv_0")] == (128, 128, 3) and d[("conv_r2", "conv_1")] == (64, 128, 3)
    assert d[("conv_r01", "conv_1")] == (3, 64, 3) and d[("down_l0", "conv_0")] == (64, 64, 2)
    for (b, l), (co, ci, k) in d.items():
        lp = P["params"][b][l]
        assert lp["weight"].shape == (co, ci, k, k, k) and lp["bias"].shape == (co,)
        assert lp["style_weight"].shape == (ci, 2) and lp["style_bias"].shape == (ci,)


def test_upsample_is_lhs_dilated_conv():
    x = torch.randn(1, 4, 3, 4, 5, dtype=torch.float64)
    w = torch.randn(6, 4, 2, 2, 2, dtype=torch.float64)
    b = torch.randn(6, dtype=torch.float64)
    a, c = _up_literal(x, w, b), _up_fast(x, w, b)
    assert a.shape == (1, 6, 6, 8, 10) and torch.allclose(a, c, atol=1e-12)
    # parity (a,b,c) of the output voxel selects tap (1-a,1-b,1-c)
    o = a[0, :, 1, 2, 5] - b
    assert torch.allclose(o, w[:, :, 0, 1, 0] @ x[0, :, 0, 1, 2], atol=1e-12)


def test_first_layer_tangent_has_extra_term():   # tests/test_style_layers_vel.py:624-651
    lp = P["params"]["conv_l00"]["conv_0"]
    s = torch.tensor([[0.1, -0.2]], dtype=torch.float64)
    wn, dw_first = modulate(lp, s, True, True, dtype=torch.float64)
    _, dw_later = modulate(lp, s, False, True, dtype=torch.float64)
    assert torch.allclose(dw_first - dw_later, wn / 0.8, atol=1e-12)
    assert torch.allclose((wn ** 2).sum(dim=(2, 3, 4, 5)), torch.ones(1, 64, dtype=torch.float64), atol=1e-6)


def test_modulation_tangent_is_derivative():
    lp = P["params"]["conv_l01"]["conv_1"]
    h = 1e-6
    s = torch.tensor([[0.3, -0.25]], dtype=torch.float64)
    wp, _ = modulate(lp, s + torch.tensor([[0, h]]), False, True, dtype=torch.float64)
    wm, _ = modulate(lp, s - torch.tensor([[0, h]]), False, True, dtype=torch.float64)
    _, dw = modulate(lp, s, False, True, dtype=torch.float64)
    assert rel_l2((wp - wm) / (2 * h), dw) < 1e-7


@pytest.fixture(scope="module")
def run104():
    x = np.random.default_rng(1234).standard_normal((1, 3, 104, 104, 104), dtype=np.float32)
    Dz = float(np.float32(oc.growth_factor(0.5, 0.3)))
    vf = float(np.float32(oc.vel_norm(0.5, 0.3)))
    Om = float(np.float32(0.3))
    d, v = Net(True, True, torch.float32).forward(P, x, Om, Dz, vf)
    return x, Om, Dz, vf, d.numpy(), v.numpy()


def test_shape_law_and_golden(run104, golden_dir):     # out = in - 96
    x, Om, Dz, vf, d, v = run104
    assert d.shape == (1, 3, 8, 8, 8) and v.shape == (1, 3, 8, 8, 8)
    assert np.isfinite(d).all() and np.isfinite(v).all()
    g = np.load(os.path.join(golden_dir, "n104.npz"))
    assert rel_l2(d, g["disp"]) < 5e-6 and rel_l2(v, g["vel"]) < 5e-5     # fp32 run vs committed fp64 fixture


def test_velocity_algebra(run104):          # tests/test_nbody_emulator_vel_core.py:189-221, 575-591
    x, Om, Dz, vf, d, v = run104
    net = Net(True, True, torch.float32)
    d2, v2 = [t.numpy() for t in net.forward(P, x, Om, Dz, 2 * vf)]
    assert np.array_equal(d2, d)                        # displacement independent of vel_fac
    assert np.allclose(v2, 2 * v, rtol=1e-5, atol=1e-6)
    _, v0 = net.forward(P, x, Om, Dz, 0.0)
    assert np.all(v0.numpy() == 0)
    dn = Net(True, False, torch.float32).forward(P, x, Om, Dz).numpy()
    assert np.allclose(dn, d, rtol=1e-5, atol=1e-6)     # primal identical with / without the tangent


def test_style_equals_premodulated(run104):
    x, Om, Dz, vf, d, v = run104
    pm = modulate_emulator_parameters(P, Dz, Om, True)
    assert set(pm["params"]["conv_l00"]["conv_0"]) == {"weight", "dweight", "bias"}
    d2, v2 = [t.numpy() for t in Net(False, True, torch.float32).forward(pm, x, None, Dz, vf)]
    assert rel_l2(d2, d) < 1e-5 and rel_l2(v2, v) < 1e-4


def test_resblock_jvp_matches_finite_difference():
    """Manual forward-mode rule of one style ResNet block vs central differences in Dz (fp64)."""
    net = Net(True, True, torch.float64)
    bp = P["params"]["conv_l01"]
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((1, 64, 9, 9, 9)))
    g = torch.from_numpy(rng.standard_normal((1, 64, 9, 9, 9)))      # x depends on Dz: x(Dz) = x + (Dz-D0) g
    D0, h = 0.8, 1e-6

    def f(Dz):
        s = torch.tensor([[0.0, Dz - 1.0]], dtype=torch.float64)
        return net.res_block(bp, "CACA", x + (Dz - D0) * g, g, s)

    y, dy = f(D0)
    yp, _ = f(D0 + h)
    ym, _ = f(D0 - h)
    assert rel_l2((yp - ym) / (2 * h), dy) < 1e-6


def test_batch_equals_singles():
    x = np.random.default_rng(5).standard_normal((2, 3, 104, 104, 104), dtype=np.float32)
    Om, Dz, vf = [0.1, 0.5], [1.0, 0.42], [30.0, 60.0]
    net = Net(True, True, torch.float32)
    d, v = [t.numpy() for t in net.forward(P, x, Om, Dz, vf)]
    for b in range(2):
        db, vb = [t.numpy() for t in net.forward(P, x[b:b + 1], Om[b], Dz[b], vf[b])]
        assert rel_l2(d[b:b + 1], db) < 1e-6 and rel_l2(v[b:b + 1], vb) < 1e-5


def test_cosmology_oracle_known_values():   # README.md:178-180; tests/test_cosmology.py:18-38, 173-194
    assert abs(oc.growth_factor(0.0, 0.3) - 1.0) < 1e-12
    assert abs(oc.hubble_rate(0.0, 0.3) - 100.0) < 1e-12
    assert abs(oc.growth_factor(0.5, 0.3) - 0.77318) < 5e-6
    assert abs(oc.growth_rate(0.5, 0.3) - 0.74922) < 5e-6
    assert abs(oc.hubble_rate(0.5, 0.3) - 130.86) < 5e-3
    assert abs(oc.vel_norm(0.5, 0.3) - 50.538) < 5e-4
    # EdS limit and f ~ Om(z)^0.55
    assert abs(oc.growth_factor(1.0, 1.0 - 1e-9) - 0.5) < 1e-6 and abs(oc.growth_rate(1.0, 1.0 - 1e-9) - 1.0) < 1e-6
    z, Om = 3.0, 0.3
    Omz = Om * (1 + z) ** 3 / (Om * (1 + z) ** 3 + 1 - Om)
    assert abs(oc.growth_rate(z, Om) / Omz ** 0.55 - 1) < 0.01
    h = 1e-5
    fd = -(np.log(oc.growth_factor(0.5 + h, 0.3)) - np.log(oc.growth_factor(0.5 - h, 0.3))) / (2 * h) * 1.5
    assert abs(fd - oc.growth_rate(0.5, 0.3)) < 1e-6


def _literal_layer(lp, x, dx, s, kind):
    """One Style layer restated with explicit index arithmetic (no torch conv): the formulas of
    SURVEY App. B / style_layers_vel.py:62-147, 235-255, in float64."""
    W = np.asarray(lp["weight"], np.float64); SW = np.asarray(lp["style_weight"], np.float64)
    sb = np.asarray(lp["style_bias"], np.float64); b = np.asarray(lp["bias"], np.float64)
    O, I, k = W.shape[0], W.shape[1], W.shape[2]
    s_mod = SW @ s + sb                                             # (I,)
    w = W * s_mod[None, :, None, None, None]
    norm = np.sqrt((w ** 2).sum(axis=(1, 2, 3, 4), keepdims=True) + 1e-8)
    wn = w / norm
    dws = W * SW[:, 1][None, :, None, None, None]
    dwn = dws / norm - w * (w * dws).sum(axis=(1, 2, 3, 4), keepdims=True) / norm ** 3
    if dx is None:
        dwn = dwn + wn / (s[1] + 1.0)

    def corr(inp, wt):
        n = inp.shape[1:]
        if kind == "up":                                            # out[2i+a] = sum_ci w[o,ci,1-a,1-b,1-c] x[ci,i,j,k]
            out = np.zeros((O,) + tuple(2 * m for m in n))
            for a in range(2):
                for bb in range(2):
                    for c in range(2):
                        out[:, a::2, bb::2, c::2] = np.einsum("oi,idhw->odhw", wt[:, :, 1 - a, 1 - bb, 1 - c], inp)
            return out
        st = 2 if kind == "down" else 1
        m = tuple((d - k) // st + 1 for d in n)
        out = np.zeros((O,) + m)
        for a in range(k):
            for bb in range(k):
                for c in range(k):
                    win = inp[:, a:a + st * (m[0] - 1) + 1:st, bb:bb + st * (m[1] - 1) + 1:st, c:c + st * (m[2] - 1) + 1:st]
                    out += np.einsum("oi,idhw->odhw", wt[:, :, a, bb, c], win)
        return out

    y = corr(x, wn) + b[:, None, None, None]
    dy = corr(x, dwn) + (corr(dx, wn) if dx is not None else 0.0)
    return y, dy


@pytest.mark.parametrize("kind,k", [("conv", 3), ("skip", 1), ("down", 2), ("up", 2)])
@pytest.mark.parametrize("first", [True, False])
def test_style_layer_matches_literal_restatement(kind, k, first):
    rng = np.random.default_rng(k * 10 + first)
    I, O, n = 3, 4, 6
    lp = {"weight": rng.standard_normal((O, I, k, k, k)), "bias": rng.standard_normal(O),
          "style_weight": rng.standard_normal((I, 2)), "style_bias": 1 + 0.1 * rng.standard_normal(I)}
    x = rng.standard_normal((I, n, n, n)); dx = None if first else rng.standard_normal((I, n, n, n))
    s = np.array([(0.31 - 0.3) * 5, 0.77 - 1.0])
    net = Net(True, True, torch.float64)
    y, dy = net.layer(lp, torch.from_numpy(x)[None], None if first else torch.from_numpy(dx)[None], torch.from_numpy(s)[None],
                      k, stride=2 if kind == "down" else 1, up=kind == "up")
    ry, rdy = _literal_layer(lp, x, dx, s, kind)
    assert y.shape[1:] == ry.shape
    assert np.allclose(y[0].numpy(), ry, rtol=1e-11, atol=1e-12)
    assert np.allclose(dy[0].numpy(), rdy, rtol=1e-10, atol=1e-11)
