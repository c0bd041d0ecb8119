"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference sources
(/root/reference/src/jax_nbody_emulator) over the numpy/torch stand-in for jax + flax in
oracle/jaxshim (jax and flax themselves are not installable in this image).

    python tools/make_reference_golden.py [name ...]

Runs in the build container only (the GPU box has no /root/reference); the fixtures travel.
Every array under a ``ref_`` file was produced by the reference's own Python: its models'
``apply``, ``modulate_emulator_parameters(_vel)``, ``SubboxConfig`` and
``SubboxProcessor.process_box``; only the primitives underneath (conv_general_dilated, jnp
functions, hyp2f1) are the stand-in's — see oracle/jaxshim/README.md for what that pins and
what it does not.  Inputs are regenerated from seeds inside the tests.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NBE_REFERENCE_SRC", "/root/reference/src")
sys.path[:0] = [os.path.join(ROOT, "oracle", "jaxshim"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402

assert jax.__version__.endswith("shim")
import jax_nbody_emulator as ref  # noqa: E402  (the reference package, unmodified)
from jax_nbody_emulator.nbody_emulator import modulate_emulator_parameters, modulate_emulator_parameters_vel  # noqa: E402

from oracle import cosmology as oc  # noqa: E402
from oracle.net import init_params  # noqa: E402  (fixed-seed parameter tree, same format as the reference's)

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(os.cpu_count())
P = init_params(42)


def field(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def tree64(t):
    return jax.tree_util.tree_map(lambda a: np.asarray(a, np.float64), t)


class x64:
    """Run the reference in double precision: jax_enable_x64 plus fp64 inputs and parameters
    (the reference computes in the dtype of ``x`` and of the parameter leaves)."""

    def __enter__(self):
        jax.config.update("jax_enable_x64", True)

    def __exit__(self, *a):
        jax.config.update("jax_enable_x64", False)


def scalars(z, Om):
    """(Dz, vel_fac, Om) handed to ``model.apply``: the fp32-rounded values tools/make_golden.py used for the
    oracle fixtures of the same names, so that reference and oracle outputs are comparable to rounding.  (The
    reference's own cosmology functions are pinned separately, ref_cosmology; they differ from these by <= 1 ulp.)"""
    z, Om = np.atleast_1d(np.asarray(z, np.float64)), np.atleast_1d(np.asarray(Om, np.float64))
    return oc.growth_factor(z, Om).astype(np.float32), oc.vel_norm(z, Om).astype(np.float32), Om.astype(np.float32)


def run_style_vel(x, z, Om, double):
    Dz, vf, Om32 = scalars(z, Om)
    m = ref.StyleNBodyEmulatorVelCore()
    if double:
        with x64():
            d, v = m.apply(tree64(P), jnp.asarray(x, dtype=jnp.float64), jnp.asarray(Om32, dtype=jnp.float64),
                           jnp.asarray(Dz, dtype=jnp.float64), jnp.asarray(vf, dtype=jnp.float64))
    else:
        d, v = m.apply(P, jnp.asarray(x), jnp.asarray(Om32), jnp.asarray(Dz), jnp.asarray(vf))
    return np.asarray(d), np.asarray(v)


def g_ref_n104():
    """All four reference models on one 104^3 input (seed 1234, z = 0.5, Om = 0.3): fp64 and fp32."""
    x = field((1, 3, 104, 104, 104), 1234)
    z, Om = 0.5, 0.3
    Dz, vf, Om32 = scalars(z, Om)
    out = dict(seed=1234, N=104, z=z, Om=Om, Dz=Dz, vel_fac=vf)
    out["disp"], out["vel"] = run_style_vel(x, z, Om, True)
    out["disp32"], out["vel32"] = run_style_vel(x, z, Om, False)
    xj = jnp.asarray(x)
    out["style_disp32"] = np.asarray(ref.StyleNBodyEmulatorCore().apply(P, xj, jnp.asarray(Om32), jnp.asarray(Dz)))
    pmv = modulate_emulator_parameters_vel(P, z, Om)
    d, v = ref.NBodyEmulatorVelCore().apply(pmv, xj, jnp.asarray(Dz), jnp.asarray(vf))
    out["premod_disp32"], out["premod_vel32"] = np.asarray(d), np.asarray(v)
    pm = modulate_emulator_parameters(P, z, Om)
    out["premod_novel_disp32"] = np.asarray(ref.NBodyEmulatorCore().apply(pm, xj, jnp.asarray(Dz)))
    return out


def g_ref_batch2():
    """Per-sample cosmologies through the reference's vmap branch (style_layers_vel.py:129-141)."""
    x = field((2, 3, 104, 104, 104), 77)
    z, Om = [0.0, 2.0], [0.1, 0.5]          # python floats, as tools/make_golden.py passes them
    out = dict(seed=77, z=z, Om=Om)
    out["disp"], out["vel"] = run_style_vel(x, z, Om, True)
    out["disp32"], out["vel32"] = run_style_vel(x, z, Om, False)
    return out


def g_ref_noncubic():
    shape = (104, 112, 120)
    x = field((1, 3) + shape, 6)
    out = dict(seed=6, shape=np.array(shape), z=1.0, Om=0.25)
    out["disp"], out["vel"] = run_style_vel(x, 1.0, 0.25, True)
    out["disp32"], out["vel32"] = run_style_vel(x, 1.0, 0.25, False)
    return out


def g_ref_n128():
    """The reference's native subbox geometry (128^3 -> 32^3), fp64 and fp32."""
    x = field((1, 3, 128, 128, 128), 9)
    out = dict(seed=9, N=128, z=0.5, Om=0.3)
    out["disp"], out["vel"] = run_style_vel(x, 0.5, 0.3, True)
    out["disp32"], out["vel32"] = run_style_vel(x, 0.5, 0.3, False)
    return out


def g_ref_n224_blocks():
    """Two 32^3 blocks of the 224^3 -> 128^3 PRODUCTION subbox (seed 225: one corner, one centre) from the reference in
    fp64 on the 128^3 input windows behind them — the construction of the oracle fixture n224_block_225.npz (VALID
    convolutions are translation-consistent; a whole 224^3 fp64 run does not fit 62 GB)."""
    x = field((1, 3, 224, 224, 224), 225)
    offsets = [(0, 0, 0), (48, 48, 48)]
    ds, vs = [], []
    for (a, b, c) in offsets:
        d, v = run_style_vel(x[:, :, a:a + 128, b:b + 128, c:c + 128], 0.5, 0.3, True)
        ds.append(d[0]), vs.append(v[0])
    return dict(disp=np.stack(ds), vel=np.stack(vs), offsets=np.array(offsets, np.int32), seed=225, N=224, z=0.5, Om=0.3)


KEEP_FULL = (("conv_l00", "conv_0"), ("conv_l00", "skip"), ("down_l0", "conv_0"), ("up_r2", "conv_0"),
             ("conv_r2", "skip"), ("conv_r01", "conv_1"))


def g_ref_modulation():
    """modulate_emulator_parameters(_vel) of the reference (nbody_emulator.py:131-264) at two cosmologies:
    six layers in full (first layer with its + W/Dz term, skip, down, up, 128-channel skip, last layer), and
    for every one of the 33 layers the sums of weight, dweight, |weight|, |dweight| in fp64."""
    out = {}
    for tag, (z, Om) in dict(a=(0.5, 0.3), b=(2.0, 0.2)).items():
        pmv = modulate_emulator_parameters_vel(P, z, Om)["params"]
        pm = modulate_emulator_parameters(P, z, Om)["params"]
        names, stats = [], []
        for b in sorted(pmv):
            for l in sorted(pmv[b]):
                w, dw = np.asarray(pmv[b][l]["weight"], np.float64), np.asarray(pmv[b][l]["dweight"], np.float64)
                assert set(pm[b][l]) == {"weight", "bias"} and set(pmv[b][l]) == {"weight", "dweight", "bias"}
                assert np.array_equal(np.asarray(pm[b][l]["weight"]), np.asarray(pmv[b][l]["weight"]))
                names.append(f"{b}/{l}")
                stats.append([w.sum(), dw.sum(), np.abs(w).sum(), np.abs(dw).sum()])
                if (b, l) in KEEP_FULL:
                    out[f"{tag}:{b}/{l}:weight"] = np.asarray(pmv[b][l]["weight"])
                    out[f"{tag}:{b}/{l}:dweight"] = np.asarray(pmv[b][l]["dweight"])
        out[f"{tag}:names"], out[f"{tag}:stats"], out[f"{tag}:zOm"] = np.array(names), np.array(stats), np.array([z, Om])
    return out


def _ref_box(size, ndiv, seed, z, Om, double):
    """The reference's SubboxProcessor.process_box loop (subbox.py:139-233) through create_emulator, Style+vel."""
    box = field((3,) + size, seed)
    if double:
        with x64():
            cfg = ref.SubboxConfig(size=size, ndiv=ndiv, dtype=jnp.float64, output_dtype=np.float64)
            emu = ref.create_emulator(compute_vel=True, load_params=False, processor_config=cfg)
            emu.params = emu.processor.params = tree64(P)
            return emu.process_box(box, z=z, Om=Om, show_progress=False)
    cfg = ref.SubboxConfig(size=size, ndiv=ndiv)
    emu = ref.create_emulator(compute_vel=True, load_params=False, processor_config=cfg)
    emu.params = emu.processor.params = P
    return emu.process_box(box, z=z, Om=Om, show_progress=False)


def g_ref_box():
    """Same box as the oracle fixture box.npz: 8x8x16, ndiv (1,1,2), seed 31 (each 104^3 window wraps the box
    13 times), in fp64 (jax_enable_x64, fp64 parameters, SubboxConfig dtype float64) and in the default fp32."""
    out = dict(size=np.array([8, 8, 16]), ndiv=np.array([1, 1, 2]), seed=31, z=0.5, Om=0.3)
    out["disp"], out["vel"] = _ref_box((8, 8, 16), (1, 1, 2), 31, 0.5, 0.3, True)
    out["disp32"], out["vel32"] = _ref_box((8, 8, 16), (1, 1, 2), 31, 0.5, 0.3, False)
    return out


def g_ref_box16():
    """A 16^3 box cut into 2x2x2 subboxes and a non-cubic 8x16x24 box with ndiv (1,2,3), fp64."""
    out = {}
    for tag, size, ndiv, seed, z, Om in (("a", (16, 16, 16), (2, 2, 2), 16, 0.5, 0.3),
                                         ("b", (8, 16, 24), (1, 2, 3), 17, 1.0, 0.27)):
        out[f"{tag}:disp"], out[f"{tag}:vel"] = _ref_box(size, ndiv, seed, z, Om, True)
        out[f"{tag}:meta"] = np.array(list(size) + list(ndiv) + [seed])
        out[f"{tag}:zOm"] = np.array([z, Om])
    return out


def g_ref_cosmology():
    """The reference's cosmology functions (cosmology.py) on a (z, Om) grid.  hyp2f1 is scipy's here, jax's own
    fp32 series there: pins the formulas (argument transformation for x < 0, normalisation, the log-derivatives)."""
    z, Om = np.meshgrid(np.array([0.0, 0.25, 0.5, 1.0, 2.0, 3.0], np.float32), np.array([0.1, 0.2, 0.3, 0.4, 0.5], np.float32))
    z, Om = jnp.asarray(z.ravel()), jnp.asarray(Om.ravel())
    return dict(z=np.asarray(z), Om=np.asarray(Om), growth_factor=np.asarray(ref.growth_factor(z, Om)),
                hubble_rate=np.asarray(ref.hubble_rate(z, Om)), growth_rate=np.asarray(ref.growth_rate(z, Om)),
                dlogH_dloga=np.asarray(ref.dlogH_dloga(z, Om)), vel_norm=np.asarray(ref.vel_norm(z, Om)),
                acc_norm=np.asarray(ref.acc_norm(z, Om)))


def g_ref_layers():
    """Single reference layers and blocks on small inputs (fp64): every layer kind with and without an input
    tangent, a ResNet block with its cropped skip, both resample blocks, LeakyReLUVel.  The parameter trees are
    drawn here (seeded) and stored with the outputs so that the tests can feed any implementation."""
    from jax_nbody_emulator.style_layers_vel import StyleConv3DVel, StyleSkip3DVel, StyleDownSample3DVel, StyleUpSample3DVel
    from jax_nbody_emulator.style_blocks_vel import StyleResNetBlock3DVel, StyleResampleBlock3DVel
    from jax_nbody_emulator.layers_vel import LeakyReLUVel
    rng = np.random.default_rng(2024)
    out = {}
    with x64():
        def lp(cin, cout, k):
            return dict(weight=rng.standard_normal((cout, cin, k, k, k)) / np.sqrt(cin * k ** 3), bias=rng.standard_normal(cout) * 0.1,
                        style_weight=rng.standard_normal((cin, 2)) * 0.3, style_bias=1.0 + 0.1 * rng.standard_normal(cin))
        s = jnp.asarray(np.array([[0.15, -0.23]]))
        x = jnp.asarray(rng.standard_normal((1, 5, 8, 9, 10)))
        dx = jnp.asarray(rng.standard_normal((1, 5, 8, 9, 10)))
        out["s"], out["x"], out["dx"] = np.asarray(s), np.asarray(x), np.asarray(dx)
        for name, cls, k in (("conv", StyleConv3DVel, 3), ("skip", StyleSkip3DVel, 1), ("down", StyleDownSample3DVel, 2),
                             ("up", StyleUpSample3DVel, 2)):
            p = lp(5, 7, k)
            for kk, vv in p.items():
                out[f"{name}:{kk}"] = vv
            m = cls(in_chan=5, out_chan=7)
            y, dy = m.apply({"params": p}, x, s, dx)
            out[f"{name}:y"], out[f"{name}:dy"] = np.asarray(y), np.asarray(dy)
            y, dy = m.apply({"params": p}, x, s, None)          # first-layer rule: + W/Dz
            out[f"{name}:y_first"], out[f"{name}:dy_first"] = np.asarray(y), np.asarray(dy)
        bp = dict(conv_0=lp(5, 7, 3), conv_1=lp(7, 7, 3), skip=lp(5, 7, 1))
        for l, p in bp.items():
            for kk, vv in p.items():
                out[f"res:{l}:{kk}"] = vv
        for seq in ("CACA", "CAC"):
            y, dy = StyleResNetBlock3DVel(seq, 2, 5, 7).apply({"params": bp}, x, s, dx)
            out[f"res:{seq}:y"], out[f"res:{seq}:dy"] = np.asarray(y), np.asarray(dy)
        for seq, k in (("DA", 2), ("UA", 2)):
            p = dict(conv_0=lp(5, 7, k))
            for kk, vv in p["conv_0"].items():
                out[f"{seq}:conv_0:{kk}"] = vv
            y, dy = StyleResampleBlock3DVel(seq, 2, 5, 7).apply({"params": p}, x, s, dx)
            out[f"{seq}:y"], out[f"{seq}:dy"] = np.asarray(y), np.asarray(dy)
        y, dy = LeakyReLUVel().apply({}, x, dx)
        out["act:y"], out["act:dy"] = np.asarray(y), np.asarray(dy)
    return out


ALL = dict(ref_cosmology=g_ref_cosmology, ref_layers=g_ref_layers, ref_modulation=g_ref_modulation, ref_n104=g_ref_n104,
           ref_batch2=g_ref_batch2, ref_noncubic=g_ref_noncubic, ref_box=g_ref_box, ref_n128=g_ref_n128, ref_box16=g_ref_box16,
           ref_n224_blocks=g_ref_n224_blocks)

if __name__ == "__main__":
    for name in (sys.argv[1:] or ALL):
        t = time.time()
        res = ALL[name]()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
        print(f"{name}: {time.time() - t:.1f} s, {os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e6:.2f} MB", flush=True)
