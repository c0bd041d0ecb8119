"""Round-2 golden fixtures (fp64 / fp32 CPU oracle; run in the build container, results committed):

    python tools/make_golden_r2.py heavy scale01 scale10 n224_full:<seed> n224_block:<seed> pk256

heavy      N = 104 Style+vel with HEAVY-TAILED, SCALE-SPREAD weights (Student-t(3) entries, log-normal per
           input-channel scale x3, style_weight x2, bias x3): fp16 hi/lo split, 2^8 weight scaling and
           subnormal handling of the packed operands are exercised away from the N(0,1)/sqrt(fan_in) init
scale01/10 N = 104 with the input field x0.1 / x10 (range check of the fp16 activations)
n224_full  224^3 -> 128^3 (BASELINE config 3 subbox): fp64 truth of the eight 32^3 CORNER blocks of the output
           (1/8 of the volume), each an fp64 forward of the 128^3 input window behind it (VALID convs are
           translation-consistent; fp64 windows larger than 128^3 do not fit this container's 62 GB)
n224_block two 32^3 blocks (a corner and the centre) of another seed
pk256      256^3 box, ndiv 2 (eight 224^3 subboxes), fp32 oracle: P(k) of the CIC density of the displaced
           lattice (oracle/density.py) + a stride-8 subsample of disp / vel
"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, layer_table
from oracle import cosmology as oc, subbox as osb, density as od

OUT = 'tests/golden'
torch.set_num_threads(int(os.environ.get('NT', os.cpu_count())))


def field(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def heavy_params(seed=21):
    rng = np.random.Generator(np.random.PCG64(seed))
    P = {}
    for b, l, co, ci, k in layer_table():
        fan_in = ci * k ** 3
        w = rng.standard_t(3, size=(co, ci, k, k, k)) / np.sqrt(3 * fan_in)
        w *= np.exp(rng.standard_normal(ci))[None, :, None, None, None]           # per-input-channel scale spread
        w *= np.exp(1.5 * rng.standard_normal(co))[:, None, None, None, None]     # per-output-channel (removed by demodulation)
        P.setdefault(b, {})[l] = {
            "weight": w.astype(np.float32),
            "bias": (0.3 * rng.standard_t(3, size=co)).astype(np.float32),
            "style_weight": (2.0 * rng.standard_normal((ci, 2)) / np.sqrt(ci)).astype(np.float32),
            "style_bias": (1.0 + 0.3 * rng.standard_normal(ci)).astype(np.float32),
        }
    return {"params": P}


def cosmo(z, Om):
    return float(np.float32(oc.growth_factor(z, Om))), float(np.float32(oc.vel_norm(z, Om)))


def run(P, x, Om, Dz, vf, dtype):
    with torch.no_grad():
        return [t.numpy() for t in Net(True, True, dtype).forward(P, x, float(np.float32(Om)), Dz, vf)]


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def with_cond(P, x, z, Om, **meta):
    Dz, vf = cosmo(z, Om)
    d, v = run(P, x, Om, Dz, vf, torch.float64)
    d32, v32 = run(P, x, Om, Dz, vf, torch.float32)
    return dict(disp=d.astype(np.float32), vel=v.astype(np.float32), cond_disp=rel(d32, d), cond_vel=rel(v32, v), z=z, Om=Om, **meta)


def g_heavy():
    return with_cond(heavy_params(21), field((1, 3, 104, 104, 104), 1234), 0.5, 0.3, seed=1234, wseed=21, N=104)


def g_scale(s):
    return with_cond(init_params(42), (field((1, 3, 104, 104, 104), 1234) * np.float32(s)), 0.5, 0.3, seed=1234, scale=s, N=104)


def n224_blocks(seed, offsets):
    """fp64 truth of 32^3 output blocks of the 224^3 -> 128^3 subbox: block at output offset (a, b, c) is an fp64
    forward of the 128^3 input window starting there (VALID convolutions are translation-consistent).  Larger fp64
    windows do not fit this container: torch's fp64 conv3d materialises 27 x 64 columns per output voxel (52 GB
    at 160^3)."""
    P = init_params(42)
    x = field((1, 3, 224, 224, 224), seed)
    Dz, vf = cosmo(0.5, 0.3)
    ds, vs = [], []
    for (a, b, c) in offsets:
        t = time.time()
        db, vb = run(P, x[:, :, a:a + 128, b:b + 128, c:c + 128], 0.3, Dz, vf, torch.float64)
        ds.append(db[0].astype(np.float32)); vs.append(vb[0].astype(np.float32))
        print('  block', (a, b, c), '%.0fs' % (time.time() - t), flush=True)
    return dict(disp=np.stack(ds), vel=np.stack(vs), offsets=np.array(offsets, np.int32), seed=seed, N=224, z=0.5, Om=0.3,
                truth='fp64 oracle, 32^3 output blocks from 128^3 input windows')


def g_n224_full(seed):        # the eight corner blocks: 1/8 of the output volume, spread over it
    return n224_blocks(seed, [(a, b, c) for a in (0, 96) for b in (0, 96) for c in (0, 96)])


def g_n224_block(seed):       # two blocks: one corner, one centre
    return n224_blocks(seed, [(0, 0, 0), (48, 48, 48)])


def g_pk256():
    size, ndiv = (256, 256, 256), (2, 2, 2)
    box = field((3,) + size, 256)
    z, Om = 0.5, 0.3
    Dz, vf = cosmo(z, Om)
    net = Net(True, True, torch.float32)
    P = init_params(42)
    def f(x):
        with torch.no_grad():
            return [o.numpy() for o in net.forward(P, x, float(np.float32(Om)), Dz, vf)]
    d, v = osb.process_box(f, box, size, ndiv, dtype=np.float32, output_dtype=np.float32)
    pk = od.power_spectrum(od.delta_from_psi(d.astype(np.float64), 256.0, worder=2), 256.0, MAS="CIC")
    return dict(k=np.asarray(pk[0]), Pk=np.asarray(pk[1]), Nmodes=np.asarray(pk[2]), disp8=d[:, ::8, ::8, ::8].astype(np.float32),
                vel8=v[:, ::8, ::8, ::8].astype(np.float32), seed=256, size=size, ndiv=ndiv, z=z, Om=Om, truth='fp32 oracle')


def g_config2():
    """BASELINE config 2: batch 8 x 128^3, z = linspace(0,2,8), Om = linspace(0.1,0.5,8), x ~ N(0,1) seed 1234
    (SURVEY 8d); fp32 oracle, sample by sample (per-sample weights)."""
    B = 8
    zs, Oms = np.linspace(0, 2, B).astype(np.float32), np.linspace(0.1, 0.5, B).astype(np.float32)
    x = field((B, 3, 128, 128, 128), 1234)
    P = init_params(42)
    ds, vs = [], []
    for b in range(B):
        Dz, vf = cosmo(float(zs[b]), float(Oms[b]))
        d, v = run(P, x[b:b + 1], float(Oms[b]), Dz, vf, torch.float32)
        ds.append(d[0]); vs.append(v[0])
        print('  sample', b, flush=True)
    return dict(disp=np.stack(ds).astype(np.float32), vel=np.stack(vs).astype(np.float32), z=zs, Om=Oms, seed=1234, truth='fp32 oracle')


for name in sys.argv[1:]:
    t = time.time()
    if name == 'heavy': r = g_heavy(); fn = 'heavy104'
    elif name == 'scale01': r = g_scale(0.1); fn = 'scale01'
    elif name == 'scale10': r = g_scale(10.0); fn = 'scale10'
    elif name.startswith('n224_full:'): s = int(name.split(':')[1]); r = g_n224_full(s); fn = f'n224_full_{s}'
    elif name.startswith('n224_block:'): s = int(name.split(':')[1]); r = g_n224_block(s); fn = f'n224_block_{s}'
    elif name == 'pk256': r = g_pk256(); fn = 'pk256'
    elif name == 'config2': r = g_config2(); fn = 'config2_b8'
    else: raise SystemExit(name)
    np.savez_compressed(os.path.join(OUT, fn + '.npz'), **r)
    print(name, 'done in %.0fs' % (time.time() - t), {k: (float(v) if np.ndim(v) == 0 and not isinstance(v, str) else '') for k, v in r.items() if k.startswith('cond')}, flush=True)
