"""GPU tests of the folded tangent (csrc/conv_mma.cuh FOLD instances).

The reference's modulated tangent weights factor as dW = W (.) (a_i + beta_o)
(style_layers_vel.py:86-93), so a layer's tangent x*dW + dx*W equals (dx + a (.) x)*W + beta (.) (x*W)
and the 3^3 velocity launches need 4 tensor-core products instead of 5.  These tests pin WHEN the folded
kernels run (Style and premodulated trees that have the structure), that the 5-product kernels take
over when they cannot (a modulation close to zero, a hand-made dweight), and that both arithmetics
agree with the oracle fixtures.
"""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
from oracle.net import Net, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3
P = nb.init_params(42)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def field(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def fold_active():
    eng = Engine.get()
    return int(eng.lib.nbe_fold_active(eng.h))


def test_fold_is_active_for_style_and_premodulated_trees(golden_dir):
    g = np.load(os.path.join(golden_dir, "n104.npz"))
    x = field((1, 3, 104, 104, 104), int(g["seed"]))
    Dz, vf = nb.growth_factor(0.5, 0.3), nb.vel_norm(0.5, 0.3)
    d, v = nb.StyleNBodyEmulatorVelCore().apply(P, x, 0.3, Dz, vf)
    assert fold_active() == 1
    assert rel_l2(d, g["disp"]) <= TOL and rel_l2(v, g["vel"]) <= TOL
    # premodulated (weight, dweight): a and beta are recovered from the ratio dW / W on the host
    pmv = nb.modulate_emulator_parameters_vel(P, 0.5, 0.3)
    d2, v2 = nb.NBodyEmulatorVelCore().apply(pmv, x, Dz, vf)
    assert fold_active() == 1
    assert rel_l2(d2, g["disp"]) <= TOL and rel_l2(v2, g["vel"]) <= TOL
    # same arithmetic up to the fp32 rounding of a and beta
    assert rel_l2(d2, d) < 1e-5 and rel_l2(v2, v) < 1e-3
    # displacement-only and single-product models never fold
    nb.StyleNBodyEmulatorCore().apply(P, x, 0.3, Dz)
    assert fold_active() == 0


def test_hand_made_dweight_keeps_the_five_product_kernels(golden_dir):
    """A dweight without the a_i + beta_o structure cannot be folded: the factorisation check rejects it and
    the result still follows the (perturbed) tangent weights exactly as the oracle does."""
    x = field((1, 3, 104, 104, 104), 77)
    Dz, vf = float(nb.growth_factor(0.5, 0.3)), float(nb.vel_norm(0.5, 0.3))
    pmv = copy.deepcopy(nb.modulate_emulator_parameters_vel(P, 0.5, 0.3))
    dw = pmv["params"]["conv_l01"]["conv_0"]["dweight"]
    rng = np.random.default_rng(5)
    pmv["params"]["conv_l01"]["conv_0"]["dweight"] = (dw + 0.05 * np.abs(dw).max() * rng.standard_normal(dw.shape)).astype(np.float32)
    d, v = nb.NBodyEmulatorVelCore().apply(pmv, x, Dz, vf)
    assert fold_active() == 0
    net = Net(False, True, torch.float64)
    with torch.no_grad():
        rd, rv = net.forward(pmv, x, None, Dz, vf)
    assert rel_l2(d, rd.numpy()) <= TOL and rel_l2(v, rv.numpy()) <= 2 * TOL
    # and the unperturbed tree folds again
    nb.NBodyEmulatorVelCore().apply(nb.modulate_emulator_parameters_vel(P, 0.5, 0.3), x, Dz, vf)
    assert fold_active() == 1


def test_modulation_close_to_zero_falls_back(golden_dir):
    """a_i = SW[i,1] / m_i: a style bias that drives one m_i to ~0 at this cosmology makes the fold vector
    useless (fp16 overflow of dx + a x); nbe_modulate notices on the host and rebuilds the 5-product layout."""
    x = field((1, 3, 104, 104, 104), 78)
    z, Om = 0.5, 0.3
    Dz, vf = float(nb.growth_factor(z, Om)), float(nb.vel_norm(z, Om))
    Pz = copy.deepcopy(P)
    lay = Pz["params"]["conv_l01"]["conv_1"]
    s0, s1 = np.float32((Om - 0.3) * 5.0), np.float32(Dz - 1.0)
    sw = lay["style_weight"]
    lay["style_bias"] = lay["style_bias"].copy()
    lay["style_bias"][7] = -(s0 * sw[7, 0] + s1 * sw[7, 1]) + 1e-4 * abs(sw[7, 1])     # m_7 = 1e-4 |SW|: a_7 = 1e4
    d, v = nb.StyleNBodyEmulatorVelCore().apply(Pz, x, Om, Dz, vf)
    assert fold_active() == 0
    net = Net(True, True, torch.float64)
    with torch.no_grad():
        rd, rv = net.forward(Pz, x, Om, Dz, vf)
    assert rel_l2(d, rd.numpy()) <= TOL and rel_l2(v, rv.numpy()) <= 2 * TOL
    # another cosmology of the same tree is fine again
    Dz2, vf2 = float(nb.growth_factor(1.0, 0.3)), float(nb.vel_norm(1.0, 0.3))
    nb.StyleNBodyEmulatorVelCore().apply(Pz, x, Om, Dz2, vf2)
    assert fold_active() == 1


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import jax_nbody_emulator_with_dj_b200 as nb
x = np.random.default_rng(1234).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
P = nb.init_params(42)
d, v = nb.StyleNBodyEmulatorVelCore().apply(P, x, 0.25, nb.growth_factor(1.0, 0.25), nb.vel_norm(1.0, 0.25))
np.savez(sys.argv[1], d=d, v=v)
"""


def test_folded_and_unfolded_kernels_agree(tmp_path):
    """Same input through NBE_FOLD=0 (5 products) and the default (4): the displacement differs only through
    the item order of the accumulations, the velocity by the fp16 rounding of (dx + a x) vs (dx, dW)."""
    outs = {}
    for fold in ("0", "1"):
        f = str(tmp_path / ("fold%s.npz" % fold))
        env = dict(os.environ, NBE_FOLD=fold)
        subprocess.run([sys.executable, "-c", _CHILD % ROOT, f], check=True, env=env, timeout=600)
        outs[fold] = np.load(f)
    ed = rel_l2(outs["1"]["d"], outs["0"]["d"])
    ev = rel_l2(outs["1"]["v"], outs["0"]["v"])
    print(f"[fold] folded vs five-product kernels: disp rel-L2 {ed:.2e}, vel rel-L2 {ev:.2e}")
    assert ed < 1e-5 and ev < 2e-3, (ed, ev)
