"""Generate tests/golden/*.npz with the fp64 CPU oracle (oracle/).  Run here (CPU container);
the fixtures travel to the GPU box, the oracle run does not have to.

    python tools/make_golden.py [name ...]

Inputs are regenerated from seeds inside the tests; only the oracle outputs are stored.
NOTE: the reference itself (JAX) cannot run in this image, so these are oracle outputs
("parity unpinned" at network level, see oracle/__init__.py)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, modulate_emulator_parameters
from oracle import cosmology as oc, subbox as osb

OUT = 'tests/golden'
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(os.cpu_count())
P = init_params(42)


def field(shape, seed=1234):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def run(x, z, Om, style=True, vel=True, params=P):
    z = np.atleast_1d(np.asarray(z, dtype=np.float64)); Om = np.atleast_1d(np.asarray(Om, dtype=np.float64))
    Dz = oc.growth_factor(z, Om).astype(np.float32).astype(np.float64)   # the product works from fp32 scalars
    vf = oc.vel_norm(z, Om).astype(np.float32).astype(np.float64)
    out = Net(style, vel, torch.float64).forward(params, x, Om.astype(np.float32).astype(np.float64), Dz, vf)
    return [o.numpy() for o in (out if vel else [out])]


def g_n104():
    d, v = run(field((1, 3, 104, 104, 104)), 0.5, 0.3)
    return dict(disp=d, vel=v, z=0.5, Om=0.3, seed=1234, N=104)


def g_batch2():
    z, Om = [0.0, 2.0], [0.1, 0.5]
    d, v = run(field((2, 3, 104, 104, 104), 77), z, Om)
    return dict(disp=d, vel=v, z=z, Om=Om, seed=77, N=104)


def g_noncubic():
    d, v = run(field((1, 3, 104, 112, 120), 5), 1.0, 0.25)
    return dict(disp=d, vel=v, z=1.0, Om=0.25, seed=5, shape=(104, 112, 120))


def g_n128():
    d, v = run(field((1, 3, 128, 128, 128), 9), 0.5, 0.3)
    return dict(disp=d, vel=v, z=0.5, Om=0.3, seed=9, N=128)


def g_box():
    size, ndiv = (8, 8, 16), (1, 1, 2)
    box = field((3,) + size, 31)
    z, Om = 0.5, 0.3
    Dz = np.float64(np.float32(oc.growth_factor(z, Om))); vf = np.float64(np.float32(oc.vel_norm(z, Om)))
    net = Net(True, True, torch.float64)
    f = lambda x: [o.numpy() for o in net.forward(P, x, np.float64(np.float32(Om)), Dz, vf)]
    d, v = osb.process_box(f, box, size, ndiv, dtype=np.float32, output_dtype=np.float64)
    return dict(disp=d, vel=v, z=z, Om=Om, seed=31, size=size, ndiv=ndiv)


ALL = dict(n104=g_n104, batch2=g_batch2, noncubic=g_noncubic, box=g_box, n128=g_n128)
for name in (sys.argv[1:] or list(ALL)):
    t = time.time()
    r = ALL[name]()
    np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **r)
    print(name, 'done in %.0fs' % (time.time() - t), flush=True)
