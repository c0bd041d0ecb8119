"""Pins of the oracle against outputs of the REFERENCE SOURCES themselves.

tests/golden/ref_*.npz were produced by tools/make_reference_golden.py: the unmodified reference package
(/root/reference/src/jax_nbody_emulator) imported over oracle/jaxshim, a numpy/torch stand-in for the jax / flax
primitives it calls (neither is installable in this image).  Every formula, index, crop and loop that produced those
arrays is the reference's own Python; oracle/jaxshim/README.md lists the primitives that are ours and states what this
does and does not pin (the algorithm, not XLA's last-bit floating point).

Three kinds of check:
  * the oracle's layers / blocks / modulation / cosmology, evaluated here, against the reference's arrays (fp64: 1e-12);
  * the committed ORACLE fixtures that the GPU parity tests use (n104, batch2, noncubic, n128, box), file against file
    with the reference fixture of the same input: they are the same numbers to fp64 rounding;
  * when /root/reference is present (build container), a live run of reference layers and SubboxConfig tables.
"""
import os
import sys

import numpy as np
import pytest
import torch

import jax_nbody_emulator_with_dj_b200 as nb
from oracle import cosmology as oc
from oracle import subbox as osb
from oracle.net import Net, init_params, modulate_emulator_parameters, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"


def load(golden_dir, name):
    p = os.path.join(golden_dir, name + ".npz")
    assert os.path.exists(p), f"{p} missing: run tools/make_reference_golden.py in the build container"
    return np.load(p)


# ------------------------------------------------------------------------------------------------ cosmology
def test_cosmology_oracle_and_product_match_reference_functions(golden_dir):
    """cosmology.py:33-155 evaluated by the reference code (fp32) on a 6 x 5 (z, Om) grid."""
    g = load(golden_dir, "ref_cosmology")
    z, Om = g["z"].astype(np.float64), g["Om"].astype(np.float64)
    for name in ("growth_factor", "hubble_rate", "growth_rate", "dlogH_dloga", "vel_norm", "acc_norm"):
        want = g[name].astype(np.float64)
        assert np.allclose(getattr(oc, name)(z, Om), want, rtol=3e-6, atol=0), name             # oracle (quadrature)
        got = np.asarray(getattr(nb, name)(g["z"], g["Om"]), np.float64)                         # product host code
        assert np.allclose(got, want, rtol=3e-6, atol=0), name


# ------------------------------------------------------------------------------------------------ layers, blocks
def _lp(g, prefix):
    return {k: g[f"{prefix}:{k}"] for k in ("weight", "bias", "style_weight", "style_bias")}


@pytest.mark.parametrize("literal_up", [False, True])
def test_oracle_layers_match_reference_layers(golden_dir, literal_up):
    """style_layers_vel.py:35-147 (conv, skip, down), :165-275 (up), with an input tangent and as a first layer
    (dx=None: the + W/Dz rule), in fp64."""
    g = load(golden_dir, "ref_layers")
    net = Net(True, True, torch.float64, literal_up=literal_up)
    x, dx, s = (torch.from_numpy(g[k]) for k in ("x", "dx", "s"))
    for name, k, stride, up in (("conv", 3, 1, False), ("skip", 1, 1, False), ("down", 2, 2, False), ("up", 2, 1, True)):
        y, dy = net.layer(_lp(g, name), x, dx, s, k, stride=stride, up=up)
        assert y.shape == g[f"{name}:y"].shape
        assert rel_l2(y.numpy(), g[f"{name}:y"]) < 1e-13 and rel_l2(dy.numpy(), g[f"{name}:dy"]) < 1e-13, name
        y, dy = net.layer(_lp(g, name), x, None, s, k, stride=stride, up=up)
        assert rel_l2(y.numpy(), g[f"{name}:y_first"]) < 1e-13 and rel_l2(dy.numpy(), g[f"{name}:dy_first"]) < 1e-13, name
        # the first-layer tangent differs from the plain weight tangent by exactly conv(x, Wn / Dz)
        assert rel_l2(g[f"{name}:dy_first"], g[f"{name}:dy"]) > 1e-3


def test_oracle_blocks_match_reference_blocks(golden_dir):
    """style_blocks_vel.py:31-166: ResNet block (cropped 1^3 skip, with and without the final activation),
    down / up resample blocks; layers_vel.py:178-186 LeakyReLUVel."""
    g = load(golden_dir, "ref_layers")
    net = Net(True, True, torch.float64)
    x, dx, s = (torch.from_numpy(g[k]) for k in ("x", "dx", "s"))
    bp = {l: _lp(g, f"res:{l}") for l in ("conv_0", "conv_1", "skip")}
    for seq in ("CACA", "CAC"):
        y, dy = net.res_block(bp, seq, x, dx, s)
        assert y.shape == g[f"res:{seq}:y"].shape
        assert rel_l2(y.numpy(), g[f"res:{seq}:y"]) < 1e-13 and rel_l2(dy.numpy(), g[f"res:{seq}:dy"]) < 1e-13, seq
    for seq in ("DA", "UA"):
        y, dy = net.resample_block({"conv_0": _lp(g, f"{seq}:conv_0")}, seq, x, dx, s)
        assert y.shape == g[f"{seq}:y"].shape
        assert rel_l2(y.numpy(), g[f"{seq}:y"]) < 1e-13 and rel_l2(dy.numpy(), g[f"{seq}:dy"]) < 1e-13, seq
    y, dy = net.act(x, dx)
    assert np.array_equal(y.numpy(), g["act:y"]) and np.array_equal(dy.numpy(), g["act:dy"])


# ------------------------------------------------------------------------------------------------ premodulation
def test_oracle_premodulation_matches_reference(golden_dir):
    """nbody_emulator.py:131-264 (both entry points) on the fixed-seed tree at two cosmologies: six layers in full,
    all 33 through their sums.  fp32 on both sides."""
    g = load(golden_dir, "ref_modulation")
    P = init_params(42)
    for tag in ("a", "b"):
        z, Om = g[f"{tag}:zOm"]
        Dz = np.float32(oc.growth_factor(z, Om))
        got = modulate_emulator_parameters(P, Dz, np.float32(Om), True)["params"]
        names = [str(n) for n in g[f"{tag}:names"]]
        assert len(names) == 33
        for n, st in zip(names, g[f"{tag}:stats"]):
            b, l = n.split("/")
            w, dw = got[b][l]["weight"].astype(np.float64), got[b][l]["dweight"].astype(np.float64)
            mine = np.array([w.sum(), dw.sum(), np.abs(w).sum(), np.abs(dw).sum()])
            assert np.allclose(mine, st, rtol=0, atol=2e-5 * st[2:].max()), (tag, n, mine, st)
        full = sorted({k.split(":")[1] for k in g.files if k.startswith(tag + ":") and k.endswith(":weight")})
        assert len(full) == 6
        for n in full:
            b, l = n.split("/")
            assert rel_l2(got[b][l]["weight"], g[f"{tag}:{n}:weight"]) < 2e-6, (tag, n)
            assert rel_l2(got[b][l]["dweight"], g[f"{tag}:{n}:dweight"]) < 2e-6, (tag, n)


# ------------------------------------------------------------------------------------------------ whole network
@pytest.mark.parametrize("ours,theirs", [("n104", "ref_n104"), ("batch2", "ref_batch2"), ("noncubic", "ref_noncubic"),
                                         ("n128", "ref_n128")])
def test_committed_oracle_fixtures_equal_reference_outputs(golden_dir, ours, theirs):
    """The fp64 oracle fixtures the GPU parity tests are gated on ARE the reference's outputs: same seeded input, same
    parameters, same (Om, Dz, vel_fac); StyleNBodyEmulatorVelCore.apply of the reference in fp64 against the file
    written by tools/make_golden.py in round 1."""
    g, r = load(golden_dir, ours), load(golden_dir, theirs)
    assert int(g["seed"]) == int(r["seed"]) and np.allclose(g["z"], r["z"]) and np.allclose(g["Om"], r["Om"])
    assert r["disp"].dtype == np.float64 and g["disp"].shape == r["disp"].shape
    ed, ev = rel_l2(g["disp"], r["disp"]), rel_l2(g["vel"], r["vel"])
    assert ed < 1e-12 and ev < 1e-12, (ed, ev)
    # the reference's default fp32 arithmetic sits where the oracle's conditioning number says plain fp32 sits
    ed32, ev32 = rel_l2(r["disp32"], r["disp"]), rel_l2(r["vel32"], r["vel"])
    assert ed32 < 5e-6 and ev32 < 3e-3, (ed32, ev32)


def test_production_geometry_truth_blocks_equal_reference_outputs(golden_dir):
    """The fp64 truth the 224^3 GPU gate uses (n224_block_225.npz: 32^3 output blocks computed from 128^3 input windows,
    stored as float32) against the reference run on the same windows."""
    g, r = load(golden_dir, "n224_block_225"), load(golden_dir, "ref_n224_blocks")
    assert int(g["seed"]) == int(r["seed"]) == 225 and np.array_equal(g["offsets"], r["offsets"])
    for b in range(2):
        assert np.array_equal(g["disp"][b], r["disp"][b].astype(np.float32)) or rel_l2(g["disp"][b], r["disp"][b]) < 1e-7
        assert rel_l2(g["vel"][b], r["vel"][b]) < 1e-7


def test_reference_variants_agree_with_each_other(golden_dir):
    """Style, premodulated+vel and premodulated models of the reference on the n104 input (fp32): the same
    displacement / velocity as StyleNBodyEmulatorVelCore (not tested anywhere in the reference's own suite)."""
    r = load(golden_dir, "ref_n104")
    assert rel_l2(r["style_disp32"], r["disp"]) < 5e-6
    assert rel_l2(r["premod_disp32"], r["disp"]) < 5e-6 and rel_l2(r["premod_novel_disp32"], r["disp"]) < 5e-6
    assert rel_l2(r["premod_vel32"], r["vel"]) < 3e-3


def test_oracle_fp32_network_matches_reference_fp32(golden_dir):
    """One live oracle run (fp32, 104^3) against the reference's fp32 output: displacement to fp32 rounding."""
    r = load(golden_dir, "ref_n104")
    x = np.random.default_rng(int(r["seed"])).standard_normal((1, 3, 104, 104, 104), dtype=np.float32)
    d, v = [t.numpy() for t in Net(True, True, torch.float32).forward(
        init_params(42), x, np.float64(np.float32(0.3)), r["Dz"].astype(np.float64), r["vel_fac"].astype(np.float64))]
    assert rel_l2(d, r["disp32"]) < 5e-6 and rel_l2(d, r["disp"]) < 5e-6
    assert rel_l2(v, r["vel"]) < 3e-3


def test_committed_box_fixture_equals_reference_process_box(golden_dir):
    """SubboxProcessor.process_box of the reference (subbox.py:139-233, periodic gather, per-subbox apply, paste) on
    the 8x8x16 box against the oracle fixture box.npz.  The reference derives Dz / vel_fac itself (cosmology.py, fp64
    under jax_enable_x64), the oracle fixture used their fp32 roundings: agreement to 1e-6, not 1e-12."""
    g, r = load(golden_dir, "box"), load(golden_dir, "ref_box")
    assert tuple(g["size"]) == tuple(r["size"]) and tuple(g["ndiv"]) == tuple(r["ndiv"]) and int(g["seed"]) == int(r["seed"])
    assert r["disp"].shape == (3, 8, 8, 16) and r["disp"].dtype == np.float64
    assert rel_l2(g["disp"], r["disp"]) < 1e-6 and rel_l2(g["vel"], r["vel"]) < 1e-6
    assert rel_l2(r["disp32"], r["disp"]) < 5e-6 and rel_l2(r["vel32"], r["vel"]) < 3e-3


# ------------------------------------------------------------------------------------------------ live (build container)
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference checkout not present (GPU box)")


@pytest.fixture(scope="module")
def ref_pkg():
    """The reference package imported over the stand-in, in a way that leaves sys.modules clean for other tests."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("jax", "flax", "jax_nbody_emulator")}
    for k in saved:
        del sys.modules[k]
    added = [os.path.join(ROOT, "oracle", "jaxshim"), REF_SRC]
    sys.path[:0] = added
    try:
        import jax
        assert jax.__version__.endswith("shim")
        import jax_nbody_emulator as ref
        yield ref
    finally:
        for a in added:
            if a in sys.path:
                sys.path.remove(a)
        for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "flax", "jax_nbody_emulator")]:
            del sys.modules[k]
        sys.modules.update(saved)


@needs_ref
@pytest.mark.parametrize("size,ndiv", [((64, 64, 64), (2, 2, 2)), ((8, 8, 16), (1, 1, 2)), ((16, 32, 48), (2, 4, 3)),
                                       ((512, 512, 512), (4, 4, 4)), ((24, 8, 40), (3, 1, 5))])
def test_live_reference_subbox_tables(ref_pkg, size, ndiv):
    """SubboxConfig of the reference (subbox.py:45-97), of the oracle and of the product: identical integer tables."""
    rc = ref_pkg.SubboxConfig(size=size, ndiv=ndiv)
    mine = nb.SubboxConfig(size=size, ndiv=ndiv)
    assert int(rc.n_subboxes) == int(mine.n_subboxes) == osb.n_subboxes(ndiv)
    assert tuple(rc.crop_size) == tuple(mine.crop_size) == tuple(osb.crop_size(size, ndiv))
    step = max(1, int(rc.n_subboxes) // 16)
    for idx in list(range(0, int(rc.n_subboxes), step)) + [int(rc.n_subboxes) - 1]:
        for kind, fn in (("all_crop_inds", osb.crop_inds), ("all_add_inds", osb.add_inds)):
            a, b, c = getattr(rc, kind)[idx], getattr(mine, kind)[idx], fn(idx, size, ndiv)
            assert a[0] == slice(None) and b[0] == slice(None)
            for d in range(1, 4):
                assert np.array_equal(np.asarray(a[d]), np.asarray(b[d])) and np.asarray(a[d]).shape == np.asarray(b[d]).shape
                assert np.array_equal(np.asarray(a[d]).ravel(), np.asarray(c[d]).ravel()), (kind, idx, d)


@needs_ref
def test_live_reference_first_block_fp64(ref_pkg):
    """The reference's first block live (StyleResNetBlock3DVel('CACA', 2, 3, 64) with dx=None: BOTH conv_0 and the skip
    take the first-layer + W/Dz rule, style_layers_vel.py:94-101), then a second block fed with its tangent, fp64."""
    import jax
    import jax.numpy as jnp
    from jax_nbody_emulator.style_blocks_vel import StyleResNetBlock3DVel
    P = init_params(42)["params"]
    x = np.random.default_rng(5).standard_normal((2, 3, 12, 11, 10))
    s = np.array([[0.05, -0.2], [-0.4, 0.3]])
    jax.config.update("jax_enable_x64", True)
    try:
        t64 = lambda bp: {"params": jax.tree_util.tree_map(lambda a: np.asarray(a, np.float64), bp)}
        y, dy = StyleResNetBlock3DVel("CACA", 2, 3, 64).apply(t64(P["conv_l00"]), jnp.asarray(x), jnp.asarray(s), None)
        y2, dy2 = StyleResNetBlock3DVel("CACA", 2, 64, 64).apply(t64(P["conv_l01"]), y, jnp.asarray(s), dy)
    finally:
        jax.config.update("jax_enable_x64", False)
    net = Net(True, True, torch.float64)
    oy, ody = net.res_block(P["conv_l00"], "CACA", torch.from_numpy(x), None, torch.from_numpy(s))
    oy2, ody2 = net.res_block(P["conv_l01"], "CACA", oy, ody, torch.from_numpy(s))
    assert y2.shape == (2, 64, 4, 3, 2)
    # s is cast to fp32 inside the reference MODEL only; the block takes it as given, so this is fp64 rounding
    for a, b in ((oy, y), (ody, dy), (oy2, y2), (ody2, dy2)):
        assert rel_l2(a.numpy(), np.asarray(b)) < 1e-13
