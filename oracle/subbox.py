"""CPU oracle of the periodic pad/crop decomposition (test infrastructure).

Restates /root/reference/src/jax_nbody_emulator/subbox.py: SubboxConfig.__post_init__
(:45-58), _get_anchor (:60-66), _compute_indices (:68-79), _get_crop_inds (:81-97) and the
serial loop of SubboxProcessor.process_box (:139-219), with plain numpy integer math.
All results are integer tables and must be reproduced bit-exactly by the product.
"""
from __future__ import annotations

import numpy as np

DEFAULT_PAD = ((48, 48), (48, 48), (48, 48))


def crop_size(size, ndiv):
    return tuple(int(s) // int(d) for s, d in zip(size, ndiv))        # floors: remainder dropped


def anchor(idx, size, ndiv):
    c = crop_size(size, ndiv)
    return ((idx // (ndiv[1] * ndiv[2])) * c[0],
            ((idx // ndiv[2]) % ndiv[1]) * c[1],
            (idx % ndiv[2]) * c[2])


def axis_indices(a, c, p0, p1, s):
    """arange(a-p0, a+c+p1) % s  -- may wrap more than once when pad > size."""
    return np.arange(a - p0, a + c + p1) % s


def crop_inds(idx, size, ndiv, pad=DEFAULT_PAD):
    a = anchor(idx, size, ndiv)
    c = crop_size(size, ndiv)
    out = [slice(None)]
    for d in range(3):
        i = axis_indices(a[d], c[d], pad[d][0], pad[d][1], size[d])
        out.append(i.reshape((-1,) + (1,) * (3 - d - 1)))
    return tuple(out)


def add_inds(idx, size, ndiv):
    return crop_inds(idx, size, ndiv, ((0, 0),) * 3)


def n_subboxes(ndiv):
    return int(np.prod(ndiv))


def process_box(apply_fn, input_box, size, ndiv, pad=DEFAULT_PAD, compute_vel=True,
                dtype=np.float32, output_dtype=np.float32, in_chan=3):
    """The reference's serial loop: gather -> cast -> apply -> cast -> paste.

    ``apply_fn(x[None])`` returns disp (1,C,c0,c1,c2) or (disp, vel)."""
    dis = np.zeros((in_chan,) + tuple(size), dtype=output_dtype)
    vel = np.zeros((in_chan,) + tuple(size), dtype=output_dtype) if compute_vel else None
    for idx in range(n_subboxes(ndiv)):
        ci = crop_inds(idx, size, ndiv, pad)
        x = np.asarray(input_box[ci], dtype=dtype)[None]
        r = apply_fn(x)
        ai = add_inds(idx, size, ndiv)
        if compute_vel:
            dis[ai] = np.asarray(r[0][0]).astype(output_dtype)
            vel[ai] = np.asarray(r[1][0]).astype(output_dtype)
        else:
            dis[ai] = np.asarray(r[0]).astype(output_dtype)
    return (dis, vel) if compute_vel else dis
