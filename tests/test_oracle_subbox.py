"""Pin the subbox oracle (oracle/subbox.py) and the product's SubboxConfig against the
expectations of the reference's own tests (/root/reference/tests/test_subbox.py:86-204) and
against each other.  Integer work: everything here is exact."""
import numpy as np
import pytest

from oracle import subbox as osb
import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200.subbox import shard_range

CONFIGS = [((256, 256, 256), (2, 2, 2)), ((128, 128, 128), (2, 2, 2)), ((64, 64, 64), (2, 2, 2)),
           ((512, 512, 512), (4, 4, 4)), ((8, 8, 16), (1, 1, 2)), ((96, 64, 160), (3, 2, 5)),
           ((100, 64, 64), (3, 2, 2)), ((32, 32, 32), (1, 1, 1))]


def test_anchor_known_answers():            # test_subbox.py:86-95
    size, ndiv = (256, 256, 256), (2, 2, 2)
    assert osb.anchor(0, size, ndiv) == (0, 0, 0)
    assert osb.anchor(1, size, ndiv) == (0, 0, 128)
    assert osb.anchor(7, size, ndiv) == (128, 128, 128)
    cfg = nb.SubboxConfig(size=size, ndiv=ndiv)
    assert cfg._get_anchor(0) == (0, 0, 0) and cfg._get_anchor(1) == (0, 0, 128) and cfg._get_anchor(7) == (128, 128, 128)


def test_config_attributes():               # test_subbox.py:40-84
    cfg = nb.SubboxConfig(size=(256, 256, 256), ndiv=(2, 2, 2))
    assert cfg.NDIM == 3 and cfg.n_subboxes == 8 and cfg.crop_size == (128, 128, 128)
    assert isinstance(cfg.crop_size, tuple) and all(isinstance(c, int) for c in cfg.crop_size)
    assert len(cfg.all_crop_inds) == 8 and len(cfg.all_add_inds) == 8
    assert cfg.all_crop_inds[0][0] == slice(None) and len(cfg.all_crop_inds[0]) == 4
    assert cfg.in_chan == 3 and cfg.padding == ((48, 48),) * 3
    assert cfg.dtype == np.float32 and cfg.output_dtype == np.float32


def test_periodic_wrap_present():           # test_subbox.py:121-134
    ci = osb.crop_inds(0, (256, 256, 256), (2, 2, 2))
    z = ci[1].ravel()
    assert np.any(z >= 208) and np.any(z < 128) and z.size == 224
    assert np.array_equal(z[:48], np.arange(208, 256)) and np.array_equal(z[48:], np.arange(0, 176))


@pytest.mark.parametrize("size,ndiv", CONFIGS)
def test_tables_product_equals_oracle(size, ndiv):
    cfg = nb.SubboxConfig(size=size, ndiv=ndiv)
    assert int(cfg.n_subboxes) == osb.n_subboxes(ndiv)
    assert cfg.crop_size == osb.crop_size(size, ndiv)
    for idx in range(int(cfg.n_subboxes)):
        a, b = cfg.all_crop_inds[idx], osb.crop_inds(idx, size, ndiv)
        c, d = cfg.all_add_inds[idx], osb.add_inds(idx, size, ndiv)
        for k in (1, 2, 3):
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k])
            assert c[k].shape == d[k].shape and np.array_equal(c[k], d[k])
            assert a[k].min() >= 0 and a[k].max() < size[k - 1]                 # :140-165
            ai = c[k].ravel()
            assert ai.size == cfg.crop_size[k - 1] and np.all(np.diff(ai) == 1)   # :167-180
    # flat tables handed to the C ABI carry exactly the same integers
    crop, add0, plen = cfg.flat_tables()
    per = sum(plen)
    for idx in range(int(cfg.n_subboxes)):
        ref = np.concatenate([osb.crop_inds(idx, size, ndiv)[k].ravel() for k in (1, 2, 3)])
        assert np.array_equal(crop[idx * per:(idx + 1) * per], ref)
        assert tuple(add0[idx * 3:idx * 3 + 3]) == osb.anchor(idx, size, ndiv)


@pytest.mark.parametrize("size,ndiv", CONFIGS)
def test_every_voxel_covered_once(size, ndiv):     # test_subbox.py:184-204 (+ remainder strips)
    cov = np.zeros(size, dtype=np.int32)
    for idx in range(osb.n_subboxes(ndiv)):
        ai = osb.add_inds(idx, size, ndiv)
        cov[ai[1], ai[2], ai[3]] += 1
    c = osb.crop_size(size, ndiv)
    owned = tuple(slice(0, c[d] * ndiv[d]) for d in range(3))
    assert np.all(cov[owned] == 1)
    rest = cov.copy(); rest[owned] = 0
    assert np.all(rest == 0)            # floor division: the remainder strip is never written


def test_multi_wrap_when_pad_exceeds_box():   # tests use 64^3 boxes with pad 48 > crop 32
    i = osb.axis_indices(0, 8, 48, 48, 8)
    assert i.size == 104 and np.array_equal(i, np.arange(-48, 56) % 8)
    assert np.array_equal(np.bincount(i), np.full(8, 13))


def test_process_box_loop_pastes_identity():
    size, ndiv = (16, 24, 32), (2, 3, 2)
    box = np.random.default_rng(0).standard_normal((3,) + size).astype(np.float32)
    f = lambda x: (x[:, :, 48:-48, 48:-48, 48:-48] * 2, x[:, :, 48:-48, 48:-48, 48:-48] + 1)
    d, v = osb.process_box(f, box, size, ndiv)
    assert np.array_equal(d, box * 2) and np.array_equal(v, box + 1)


@pytest.mark.parametrize("n,world", [(64, 1), (64, 8), (64, 3), (5, 8), (512, 8), (7, 2)])
def test_shard_range_partition(n, world):
    seen = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        assert 0 <= lo <= hi <= n
        seen += list(range(lo, hi))
    assert seen == list(range(n))
    sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
