"""GPU debug aid: run one sample through libnbe_b200 and compare every stored activation with
the CPU oracle's intermediates.  python tools/debug_layers.py [N] [style|premod] [vel|novel] [split|fp16]"""
import sys, ctypes as C, time
import numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
from oracle.net import Net, rel_l2, modulate_emulator_parameters
from oracle import cosmology as oc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 104
style = (sys.argv[2] if len(sys.argv) > 2 else 'style') == 'style'
vel = (sys.argv[3] if len(sys.argv) > 3 else 'vel') == 'vel'
prec = sys.argv[4] if len(sys.argv) > 4 else 'split'
ACTS = ['in', 'conv_l00.conv_0', 'conv_l00', 'conv_l01.conv_0', 'conv_l01', 'down_l0', 'conv_l1.conv_0', 'conv_l1',
        'down_l1', 'conv_l2.conv_0', 'conv_l2', 'down_l2', 'conv_c.conv_0', 'conv_c', 'up_r2', 'conv_r2.conv_0',
        'conv_r2', 'up_r1', 'conv_r1.conv_0', 'conv_r1', 'up_r0', 'conv_r00.conv_0', 'conv_r00', 'conv_r01.conv_0']
params = nb.init_params(42)
x = np.random.default_rng(1234).standard_normal((1, 3, N, N, N), dtype=np.float32)
z, Om = 0.5, 0.3
Dz = float(oc.growth_factor(z, Om)); vf = float(oc.vel_norm(z, Om))
oparams = params if style else modulate_emulator_parameters(params, Dz, Om, vel)
cls = {(True, True): nb.StyleNBodyEmulatorVelCore, (True, False): nb.StyleNBodyEmulatorCore,
       (False, True): nb.NBodyEmulatorVelCore, (False, False): nb.NBodyEmulatorCore}[(style, vel)]
model = cls(); model.precision = prec
args = ([Om] if style else []) + [Dz] + ([vf] if vel else [])
t = time.time()
out = model.apply(oparams, x, *args)
torch.cuda.synchronize()
print('gpu apply %.2fs' % (time.time() - t), flush=True)
net = Net(style, vel, torch.float32); net.cap = {}
torch.set_num_threads(max(1, torch.get_num_threads()))
t = time.time()
ref = net.forward(oparams, x, Om if style else None, Dz, vf if vel else None)
print('oracle %.1fs' % (time.time() - t), flush=True)
eng = Engine.get()
lib = eng.lib
for aid, name in enumerate(ACTS):
    shp = (C.c_int32 * 4)()
    nbytes = lib.nbe_debug_read_act(eng.h, aid, 0, None, 0, shp)
    d, h, w, c = list(shp)
    def rd(which):
        buf = np.empty((d, h, w, c), dtype=np.float16)
        r = lib.nbe_debug_read_act(eng.h, aid, which, C.c_void_p(buf.ctypes.data), buf.nbytes, shp)
        assert r == buf.nbytes, r
        return buf.astype(np.float32)
    rx, rdx = net.cap[name]
    rx = rx[0].permute(1, 2, 3, 0).numpy()
    hi = rd(0)
    if name == 'in':
        got = hi[..., 0:3] + hi[..., 3:6]
        print('%-18s x %.3e' % (name, rel_l2(got, rx)), flush=True)
        continue
    got = hi + (rd(1) if prec == 'split' else 0)
    proj = float(((got - rx) * rx).sum() / (rx * rx).sum())
    msg = '%-18s %s x %.3e proj %+.2e' % (name, (d, h, w, c), rel_l2(got, rx), proj)
    if vel:
        msg += '  dx %.3e' % rel_l2(rd(2), rdx[0].permute(1, 2, 3, 0).numpy())
    print(msg, flush=True)
if vel:
    print('disp %.3e vel %.3e' % (rel_l2(out[0], ref[0].numpy()), rel_l2(out[1], ref[1].numpy())))
else:
    print('disp %.3e' % rel_l2(out, ref.numpy()))
