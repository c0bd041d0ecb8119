"""CPU oracle of the emulator network (test infrastructure, see oracle/__init__.py).

A restatement, with torch-CPU convolutions, of the four reference models

  StyleNBodyEmulatorVelCore   style_nbody_emulator_vel_core.py:105-195
  StyleNBodyEmulatorCore      style_nbody_emulator_core.py:100-175
  NBodyEmulatorVelCore        nbody_emulator_vel_core.py:103-183
  NBodyEmulatorCore           nbody_emulator_core.py

their blocks (style_blocks_vel.py:31-166, blocks_vel.py:30-159) and their layers
(style_layers_vel.py:35-147 conv, :165-275 up-sampling; layers_vel.py:33-96, :112-176,
:178-186 LeakyReLUVel; style_layers.py:35-105).  One generic implementation covers the
four variants through the flags ``style`` (modulate on the fly vs. premodulated
``weight``/``dweight``) and ``vel`` (carry the Dz-tangent ``dx`` or not).

Pinned against the reference sources executed over oracle/jaxshim (fp64 agreement 1e-13 per
layer / block, 1e-12 for the whole network: tests/test_oracle_vs_reference.py); see the
package docstring for what that does and does not cover.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# (block name, kind, seq, in_chan, out_chan) in execution order of the reference's setup()
# style_nbody_emulator_vel_core.py:46-103
BLOCKS = (
    ("conv_l00", "res", "CACA", 3, 64),
    ("conv_l01", "res", "CACA", 64, 64),
    ("down_l0", "resample", "DA", 64, 64),
    ("conv_l1", "res", "CACA", 64, 64),
    ("down_l1", "resample", "DA", 64, 64),
    ("conv_l2", "res", "CACA", 64, 64),
    ("down_l2", "resample", "DA", 64, 64),
    ("conv_c", "res", "CACA", 64, 64),
    ("up_r2", "resample", "UA", 64, 64),
    ("conv_r2", "res", "CACA", 128, 64),
    ("up_r1", "resample", "UA", 64, 64),
    ("conv_r1", "res", "CACA", 128, 64),
    ("up_r0", "resample", "UA", 64, 64),
    ("conv_r00", "res", "CACA", 128, 64),
    ("conv_r01", "res", "CAC", 64, 3),
)
NEG_SLOPE = 0.01
EPS = 1e-8


def layer_table(in_chan=3, out_chan=3, mid_chan=64):
    """[(block, layer, Cout, Cin, k)] for all 33 conv layers."""
    rows = []
    m1, m2 = mid_chan, 2 * mid_chan
    spec = {
        "conv_l00": (in_chan, m1), "conv_l01": (m1, m1), "down_l0": (m1, m1),
        "conv_l1": (m1, m1), "down_l1": (m1, m1), "conv_l2": (m1, m1), "down_l2": (m1, m1),
        "conv_c": (m1, m1), "up_r2": (m1, m1), "conv_r2": (m2, m1), "up_r1": (m1, m1),
        "conv_r1": (m2, m1), "up_r0": (m1, m1), "conv_r00": (m2, m1), "conv_r01": (m1, out_chan),
    }
    for name, kind, seq, _, _ in BLOCKS:
        cin, cout = spec[name]
        if kind == "res":
            mid = max(cin, cout)          # style_blocks_vel.py:126
            rows.append((name, "skip", cout, cin, 1))
            rows.append((name, "conv_0", mid, cin, 3))
            rows.append((name, "conv_1", cout, mid, 3))
        else:
            rows.append((name, "conv_0", cout, cin, 2))
    return rows


def _t(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.as_tensor(np.asarray(a)).to(dtype)


# ----------------------------------------------------------------------------------------
# weight modulation  (style_layers_vel.py:62-105, nbody_emulator.py:131-148, :189-219)
# ----------------------------------------------------------------------------------------
def modulate(lp, s, first, vel, eps=EPS, dtype=torch.float32):
    """Return (Wn, dWn|None): per-sample modulated + demodulated weights, shape (B,O,I,k,k,k).

    ``s`` is the (B,2) style vector [(Om-0.3)*5, Dz-1]; ``first`` marks layers whose input is
    the raw Dz-scaled field (dx is None), which receive the extra ``Wn/Dz`` term.
    """
    W = _t(lp["weight"], dtype)
    SW = _t(lp["style_weight"], dtype)
    sb = _t(lp["style_bias"], dtype)
    s = _t(s, dtype).reshape(-1, 2)
    s_mod = s @ SW.T + sb                                   # (B, I)
    sm = s_mod[:, None, :, None, None, None]
    w = W[None] * sm
    norm = torch.sqrt((w * w).sum(dim=(2, 3, 4, 5), keepdim=True) + eps)
    wn = w / norm
    if not vel:
        return wn, None
    ds_mod = SW[:, 1][None, None, :, None, None, None]      # d s_mod / d s1
    dws = W[None] * ds_mod
    dnorm = -(w * dws).sum(dim=(2, 3, 4, 5), keepdim=True) / norm ** 3
    dwn = dws / norm + w * dnorm
    if first:
        Dz = (s[:, 1] + 1.0)[:, None, None, None, None, None]
        dwn = dwn + wn / Dz
    return wn, dwn


# ----------------------------------------------------------------------------------------
# convolutions
# ----------------------------------------------------------------------------------------
def _conv(x, w, b, stride):
    return F.conv3d(x, w, b, stride=stride)


def _up_literal(x, w, b):
    """conv_general_dilated(lhs_dilation=2, padding=1, k=2), literally
    (style_layers_vel.py:235-255): zero-insert, pad 1, VALID correlation."""
    B, C, D, H, Wd = x.shape
    z = x.new_zeros((B, C, 2 * D + 1, 2 * H + 1, 2 * Wd + 1))
    z[:, :, 1:-1:2, 1:-1:2, 1:-1:2] = x
    return F.conv3d(z, w, b)


def _up_fast(x, w, b):
    """Same map as _up_literal: out[2i+a] = sum_ci w[o,ci,1-a] x[ci,i]  (SURVEY App. C.8)."""
    wt = w.flip(2, 3, 4).permute(1, 0, 2, 3, 4).contiguous()
    return F.conv_transpose3d(x, wt, b, stride=2)


class Net:
    """Functional forward pass.  ``q_act``/``q_w`` are optional rounding hooks used by the
    numerics study (emulating fp16/bf16 storage of activations / packed weights)."""

    def __init__(self, style=True, vel=True, dtype=torch.float64, mod_dtype=None,
                 eps=EPS, q_act=None, q_w=None, literal_up=False, ops=None):
        self.style, self.vel = style, vel
        self.dtype = dtype
        self.mod_dtype = mod_dtype or dtype
        self.eps = eps
        self.q_act = q_act or (lambda t: t)
        self.q_w = q_w or (lambda t: t)
        self.up = _up_literal if literal_up else _up_fast
        # operand-level rounding hooks for the numerics study: how x / W enter the primal conv
        # ('xp','wp'), how x / W / dW / dx enter the two tangent convs ('xt','wt','dw','dx')
        ident = lambda t: t
        self.ops = {k: (ops or {}).get(k, ident) for k in ("xp", "wp", "xt", "wt", "dw", "dx")}
        # numerics study only: per-block overrides of `ops` ({block name: ops dict}); the block being
        # evaluated is tracked in cur_block by forward()
        self.ops_by_block = {}
        self.cur_block = None
        self.cap = None          # optional dict: name -> (x, dx) of every stored activation

    # ---- one conv layer, primal + tangent  (style_layers_vel.py:108-147) ----
    def layer(self, lp, x, dx, s, k, stride=1, up=False):
        B = x.shape[0]
        first = dx is None
        bias = _t(lp["bias"], self.dtype)
        if self.style:
            wn, dwn = modulate(lp, s, first, self.vel, self.eps, self.mod_dtype)
            wn = self.q_w(wn.to(self.dtype))
            dwn = self.q_w(dwn.to(self.dtype)) if dwn is not None else None
            per_sample = True
        else:
            wn = self.q_w(_t(lp["weight"], self.dtype))[None]
            dwn = self.q_w(_t(lp["dweight"], self.dtype))[None] if self.vel else None
            per_sample = False
        f = (lambda a, w, b: self.up(a, w, b)) if up else (lambda a, w, b: _conv(a, w, b, stride))
        ys, dys = [], []
        for i in range(B):
            j = i if per_sample else 0
            xi = x[i:i + 1]
            o = self.ops_by_block.get(self.cur_block, self.ops)
            ys.append(f(o["xp"](xi), o["wp"](wn[j]), bias))
            if self.vel:
                dy = f(o["xt"](xi), o["dw"](dwn[j]), None)
                if dx is not None:
                    dy = dy + f(o["dx"](dx[i:i + 1]), o["wt"](wn[j]), None)
                dys.append(dy)
        y = torch.cat(ys, 0)
        dy = torch.cat(dys, 0) if self.vel else None
        return y, dy

    def act(self, x, dx):
        y = torch.where(x >= 0, x, NEG_SLOPE * x)           # jax.nn.leaky_relu
        dy = None
        if dx is not None:
            dy = torch.where(x > 0, dx, NEG_SLOPE * dx)     # layers_vel.py:185 (strict >)
        return y, dy

    # ---- blocks  (style_blocks_vel.py:96-166, :40-85) ----
    def res_block(self, bp, seq, x, dx, s, name=None):
        last_act = seq[-1] == "A"
        main = seq[:-1] if last_act else seq
        y, dy = self.layer(bp["skip"], x, dx, s, 1)
        nconv = main.count("C")
        if nconv > 0:
            c = nconv
            y = y[:, :, c:-c, c:-c, c:-c]
            dy = dy[:, :, c:-c, c:-c, c:-c] if dy is not None else None
        ci = 0
        for ch in main:
            if ch == "C":
                x, dx = self.layer(bp[f"conv_{ci}"], x, dx, s, 3)
                ci += 1
            elif ch == "A":
                x, dx = self.act(x, dx)
                x = self.q_act(x)
                dx = self.q_act(dx) if dx is not None else None
                if self.cap is not None and name is not None:
                    self.cap[name + ".conv_0"] = (x, dx)
            else:
                raise ValueError(f'Layer type "{ch}" not supported. Use C (conv) or A (activation).')
        x = x + y
        dx = dx + dy if dx is not None else None
        if last_act:
            x, dx = self.act(x, dx)
            x = self.q_act(x)
            dx = self.q_act(dx) if dx is not None else None
        return x, dx

    def resample_block(self, bp, seq, x, dx, s):
        ci = 0
        for ch in seq:
            if ch == "U":
                x, dx = self.layer(bp[f"conv_{ci}"], x, dx, s, 2, up=True)
                ci += 1
            elif ch == "D":
                x, dx = self.layer(bp[f"conv_{ci}"], x, dx, s, 2, stride=2)
                ci += 1
            elif ch == "A":
                x, dx = self.act(x, dx)
                x = self.q_act(x)
                dx = self.q_act(dx) if dx is not None else None
            else:
                raise ValueError(f'Layer type "{ch}" not supported.')
        return x, dx

    # ---- whole model ----
    def forward(self, params, x, Om=None, Dz=None, vel_fac=None, return_net=False):
        P = params["params"]
        dt = self.dtype
        x = _t(x, dt)
        Dz = _t(np.atleast_1d(np.asarray(Dz, dtype=np.float64)), torch.float64)
        B = x.shape[0]
        s = None
        if self.style:
            Om = _t(np.atleast_1d(np.asarray(Om, dtype=np.float64)), torch.float64)
            s = torch.stack([(Om - 0.3) * 5.0, Dz - 1.0], dim=-1)
            if self.vel:        # style_nbody_emulator_vel_core.py:128 rounds the style vector to float32 whatever
                s = s.to(torch.float32)     # the compute dtype (found by running the reference over oracle/jaxshim)
            s = s.to(self.mod_dtype)
            if s.shape[0] == 1 and B > 1:
                s = s.expand(B, 2)
        Dzb = Dz.to(dt)[:, None, None, None, None]
        x = x * (Dzb / 6.0)
        x0 = x[:, :, 48:-48, 48:-48, 48:-48]
        x = self.q_act(x)
        dx = None
        if self.cap is not None:
            self.cap["in"] = (x, None)
        blk = {name: (kind, seq) for name, kind, seq, _, _ in BLOCKS}

        def run(name, x, dx):
            kind, seq = blk[name]
            self.cur_block = name
            if kind == "res":
                r = self.res_block(P[name], seq, x, dx, s, name)
            else:
                r = self.resample_block(P[name], seq, x, dx, s)
            if self.cap is not None:
                self.cap[name] = r
            return r

        def crop(t, c):
            return None if t is None else t[:, :, c:-c, c:-c, c:-c]

        def cat(a, b):
            return None if a is None else torch.cat([a, b], dim=1)

        x, dx = run("conv_l00", x, dx)
        y0, dy0 = run("conv_l01", x, dx)
        x, dx = run("down_l0", y0, dy0)
        y0, dy0 = crop(y0, 40), crop(dy0, 40)
        y1, dy1 = run("conv_l1", x, dx)
        x, dx = run("down_l1", y1, dy1)
        y1, dy1 = crop(y1, 16), crop(dy1, 16)
        y2, dy2 = run("conv_l2", x, dx)
        x, dx = run("down_l2", y2, dy2)
        y2, dy2 = crop(y2, 4), crop(dy2, 4)
        x, dx = run("conv_c", x, dx)
        x, dx = run("up_r2", x, dx)
        x, dx = run("conv_r2", cat(y2, x), cat(dy2, dx))
        x, dx = run("up_r1", x, dx)
        x, dx = run("conv_r1", cat(y1, x), cat(dy1, dx))
        x, dx = run("up_r0", x, dx)
        x, dx = run("conv_r00", cat(y0, x), cat(dy0, dx))
        x, dx = run("conv_r01", x, dx)
        if return_net:
            return x, dx
        disp = (x + x0) * 6.0
        if not self.vel:
            return disp
        vf = _t(np.atleast_1d(np.asarray(vel_fac, dtype=np.float64)), dt)[:, None, None, None, None]
        vel = dx * (vf * 6.0) + x0 * (vf * 6.0 / Dzb)
        return disp, vel


# ----------------------------------------------------------------------------------------
# premodulation  (nbody_emulator.py:150-187, :221-266)
# ----------------------------------------------------------------------------------------
def modulate_emulator_parameters(params, Dz, Om, vel, eps=EPS, dtype=torch.float32):
    """Premodulated parameter tree for fixed (Dz, Om).  Takes Dz (not z): the growth factor
    is computed by the caller so that this function has no cosmology dependency."""
    s = torch.tensor([[(float(Om) - 0.3) * 5.0, float(Dz) - 1.0]], dtype=dtype)
    out = {"params": {}}
    for bname, bp in params["params"].items():
        out["params"][bname] = {}
        for lname, lp in bp.items():
            if "style_weight" not in lp:
                out["params"][bname][lname] = lp
                continue
            first = vel and bname == "conv_l00" and lname in ("conv_0", "skip")
            # non-first layers: dx=1 in the reference => no Wn/Dz term
            wn, dwn = modulate(lp, s, first, vel, eps, dtype)
            ent = {"weight": wn[0].numpy(), "bias": np.asarray(lp["bias"])}
            if vel:
                ent = {"weight": wn[0].numpy(), "dweight": dwn[0].numpy(), "bias": np.asarray(lp["bias"])}
            out["params"][bname][lname] = ent
    return out


# ----------------------------------------------------------------------------------------
# seeded parameters (PCG64).  The reference init (lecun_normal / ones / zeros via
# the JAX PRNG) is not reproducible without JAX; after demodulation the weight scale is
# irrelevant.  bias and style_bias are perturbed on purpose so that they are exercised.
# ----------------------------------------------------------------------------------------
def init_params(seed=42, in_chan=3, out_chan=3, mid_chan=64, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    P = {}
    for b, l, co, ci, k in layer_table(in_chan, out_chan, mid_chan):
        fan_in = ci * k ** 3
        P.setdefault(b, {})[l] = {
            "weight": (rng.standard_normal((co, ci, k, k, k)) / np.sqrt(fan_in)).astype(dtype),
            "bias": (0.1 * rng.standard_normal(co)).astype(dtype),
            "style_weight": (rng.standard_normal((ci, 2)) / np.sqrt(ci)).astype(dtype),
            "style_bias": (1.0 + 0.1 * rng.standard_normal(ci)).astype(dtype),
        }
    return {"params": P}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
