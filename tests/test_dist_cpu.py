"""World-size-2 gloo test of the multi-rank host logic: contiguous sharding of the subbox
index range and the disjoint-output gather (sum of per-rank boxes).  No GPU involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200.subbox import _gather_outputs, shard_range
from oracle import subbox as osb


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    size, ndiv = (16, 16, 32), (2, 2, 4)
    cfg = nb.SubboxConfig(size=size, ndiv=ndiv)
    lo, hi = shard_range(int(cfg.n_subboxes), rank, world)
    box = np.random.default_rng(11).standard_normal((3,) + size).astype(np.float32)
    dis = np.zeros((3,) + size, np.float32)
    vel = np.zeros((3,) + size, np.float32)
    for idx in range(lo, hi):                       # stand-in for the GPU net: a crop-wise map
        ai = cfg.all_add_inds[idx]
        dis[ai] = box[ai] * 2
        vel[ai] = box[ai] - 1
    keep_d, keep_v = dis.copy(), vel.copy()
    d_all, v_all = _gather_outputs(dist, cfg, dis, vel, world, "all")
    d0, v0 = _gather_outputs(dist, cfg, dis, vel, world, "rank0")
    ok = np.array_equal(d_all, box * 2) and np.array_equal(v_all, box - 1)
    # the reduction never writes into the caller's arrays (they may be the processor's cached buffers)
    ok = ok and np.array_equal(dis, keep_d) and np.array_equal(vel, keep_v)
    if rank == 0:
        ok = ok and np.array_equal(d0, box * 2) and np.array_equal(v0, box - 1)
    own = np.zeros(size, bool)
    for idx in range(lo, hi):
        ai = osb.add_inds(idx, size, ndiv)
        own[ai[1], ai[2], ai[3]] = True
    ok = ok and np.all((dis[0] != 0) <= own)
    q.put((rank, bool(ok), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 8, 8, 16)
