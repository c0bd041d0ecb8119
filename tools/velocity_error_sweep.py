"""CPU study behind the velocity gate: seed sweep of the velocity rel-L2 error
of several operand arithmetics against the fp64 oracle, no seed selection.

    python tools/velocity_error_sweep.py <geom> <seed0> <nseeds> [threads]     geom: nc | n128

Arithmetics (emulated on the CPU with exact fp32 accumulation; r1 = one fp16 rounding, r2 = fp16
hi+lo pair, id = plain fp32):
  fp32   everything fp32 (what a faithful fp32 implementation gives: the conditioning of the input)
  prod   the shipped product arithmetic: primal r2 x r2, tangent operands r1
  floor  primal fp32, tangent r1                      -> the tangent-rounding floor alone
  flips  primal r2 x r2, tangent fp32                 -> the mask-flip part alone
  deep   prod, but the small deep blocks (conv_l2 .. conv_r2) in fp32
  deep2  prod, but the deep blocks with hi+lo tangent operands (r2) and primal fp32
  tact   prod with hi+lo tangent ACTIVATIONS (xt, dx = r2; dW, W in the tangent = r1)
Results are appended as JSON lines to profiles/r2_velocity_sweep.jsonl; the fp64 outputs go to
tests/golden/sweep/<geom>_<seed>.npz (float32; with the emulated product error when it was computed) --
the GPU seed-sweep test (tests/test_gpu_parity.py::test_velocity_seed_sweep) checks the kernels against
them.  A 5th argument "truth" computes only the fp64 outputs (seeds whose variants are already logged).
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, rel_l2
from oracle import cosmology as oc

geom, seed0, nseeds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
torch.set_num_threads(int(sys.argv[4]) if len(sys.argv) > 4 else 6)
truth_only = len(sys.argv) > 5 and sys.argv[5] == "truth"
os.makedirs('tests/golden/sweep', exist_ok=True)
P = init_params(42)
shape, z, Om = {'nc': ((104, 112, 120), 1.0, 0.25), 'n128': ((128, 128, 128), 0.5, 0.3),
                'n104': ((104, 104, 104), 0.5, 0.3)}[geom]
Dz = float(np.float32(oc.growth_factor(z, Om))); vf = float(np.float32(oc.vel_norm(z, Om)))
h = torch.float16
ident = lambda t: t
def r1(t): return t.to(h).to(t.dtype)
def r2(t):
    a = t.to(h).to(t.dtype); return a + (t - a).to(h).to(t.dtype)
def ops(xp, wp, xt, wt, dw, dx): return dict(xp=xp, wp=wp, xt=xt, wt=wt, dw=dw, dx=dx)
PROD = ops(r2, r2, r1, r1, r1, r1)
F32 = ops(ident, ident, ident, ident, ident, ident)
DEEP = ('conv_l2', 'down_l2', 'conv_c', 'up_r2', 'conv_r2')
variants = {
    'fp32': (F32, {}),
    'prod': (PROD, {}),
    'floor': (ops(ident, ident, r1, r1, r1, r1), {}),
    'flips': (ops(r2, r2, ident, ident, ident, ident), {}),
    'deep': (PROD, {b: F32 for b in DEEP}),
    'deep2': (PROD, {b: ops(ident, ident, r2, r2, r2, r2) for b in DEEP}),
    'tact': (ops(r2, r2, r2, r1, r1, r2), {}),
}
out = 'profiles/r2_velocity_sweep.jsonl'
for seed in range(seed0, seed0 + nseeds):
    x = np.random.default_rng(seed).standard_normal((1, 3) + shape, dtype=np.float32)
    t0 = time.time()
    with torch.no_grad():
        d64, v64 = [t.numpy() for t in Net(True, True, torch.float64).forward(P, x, float(np.float32(Om)), Dz, vf)]
        row = dict(geom=geom, shape=shape, seed=seed, z=z, Om=Om)
        for name, (o, byb) in ({} if truth_only else variants).items():
            net = Net(True, True, torch.float32, ops=o)
            net.ops_by_block = byb
            d, v = [t.numpy() for t in net.forward(P, x, float(np.float32(Om)), Dz, vf)]
            row[name] = [rel_l2(d, d64), rel_l2(v, v64)]
    row['sec'] = round(time.time() - t0)
    np.savez_compressed(f'tests/golden/sweep/{geom}_{seed}.npz', disp=d64.astype(np.float32), vel=v64.astype(np.float32),
                        seed=seed, shape=shape, z=z, Om=Om)
    if not truth_only:
        with open(out, 'a') as f:
            f.write(json.dumps(row) + '\n')
    print(json.dumps(row), flush=True)
