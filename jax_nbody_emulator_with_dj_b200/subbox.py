"""Subbox decomposition of large periodic boxes, drop-in for the reference ``subbox.py``.

``SubboxConfig`` reproduces the reference's integer tables bit-exactly (anchor order,
``arange(a-p0, a+c+p1) % size`` periodic indices, floor division of the box -- remainder
strips stay zero; subbox.py:45-97).  ``SubboxProcessor.process_box`` keeps the reference
signature (subbox.py:139-147) but, instead of a serial numpy-gather / device_put / apply /
np.asarray loop, uploads the box once and runs gather, net and paste on the GPU
(``nbe_process_box``).  With ``torch.distributed`` initialised the subboxes are sharded over
the ranks (one process per GPU, contiguous index ranges, no data-path collective).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _lib
from ._engine import Engine, dtype_code, params_fingerprint, _torch
from .cosmology import growth_factor, vel_norm


@dataclass
class SubboxConfig:
    """Configuration for subbox processing (same fields as the reference, subbox.py:25-43)."""
    size: tuple
    ndiv: tuple
    dtype: Any = np.float32
    output_dtype: Any = np.float32
    in_chan: int = 3
    padding: tuple = ((48, 48), (48, 48), (48, 48))

    def __post_init__(self):
        self.NDIM = 3
        self.n_subboxes = np.prod(self.ndiv)
        self.crop_size = tuple(s // d for s, d in zip(self.size, self.ndiv))
        self.all_crop_inds = []
        self.all_add_inds = []
        for idx in range(self.n_subboxes):
            crop_inds, add_inds = self._compute_indices(idx)
            self.all_crop_inds.append(crop_inds)
            self.all_add_inds.append(add_inds)

    def _get_anchor(self, idx):
        n1, n2 = self.ndiv[1], self.ndiv[2]
        c = self.crop_size
        return ((idx // (n1 * n2)) * c[0], ((idx // n2) % n1) * c[1], (idx % n2) * c[2])

    def _compute_indices(self, idx):
        anchor = self._get_anchor(idx)
        crop_inds = self._get_crop_inds(anchor, self.crop_size, self.padding)
        add_inds = self._get_crop_inds(anchor, self.crop_size, ((0, 0),) * self.NDIM)
        return crop_inds, add_inds

    def _get_crop_inds(self, anchor, crop, pad):
        ind = [slice(None)]
        for d in range(self.NDIM):
            lo = anchor[d] - pad[d][0]
            hi = anchor[d] + crop[d] + pad[d][1]
            i = np.arange(lo, hi) % self.size[d]
            ind.append(i.reshape((-1,) + (1,) * (self.NDIM - d - 1)))
        return tuple(ind)

    # ---- compact int32 tables handed to the C ABI
    def flat_tables(self):
        """(crop_idx int32 [n_sub*(p0+p1+p2)], add_idx0 int32 [n_sub*3], plen)."""
        plen = tuple(int(c + p[0] + p[1]) for c, p in zip(self.crop_size, self.padding))
        n = int(self.n_subboxes)
        crop = np.empty((n, sum(plen)), dtype=np.int32)
        add0 = np.empty((n, 3), dtype=np.int32)
        for idx in range(n):
            ci, ai = self.all_crop_inds[idx], self.all_add_inds[idx]
            crop[idx] = np.concatenate([np.asarray(ci[d + 1]).ravel() for d in range(3)])
            add0[idx] = [int(np.asarray(ai[d + 1]).ravel()[0]) if self.crop_size[d] > 0 else 0 for d in range(3)]
        return np.ascontiguousarray(crop.ravel()), np.ascontiguousarray(add0.ravel()), plen


def shard_range(n, rank, world):
    """Contiguous index range of `rank` (C-order => D-slabs first); the first n % world ranks
    get one extra subbox."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _pinned_zeros(shape, np_dtype):
    torch = _torch()
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float16): torch.float16}[np.dtype(np_dtype)]
    t = torch.zeros(shape, dtype=tdt, pin_memory=True)
    return t, t.numpy()


class _OutputPool:
    """Page-locked output boxes, handed to the caller WITHOUT a copy and never written again while the
    caller (or any view of the array) is alive.

    Every GPU DMAs its finished blocks straight into the array ``process_box`` returns.  Page-locking a
    multi-GB buffer costs about as much as a whole 8-GPU box, so buffers are recycled -- but only once
    the array handed out has been garbage-collected (tracked with a weak reference; views keep their
    base alive).  A caller who keeps every result therefore gets a fresh buffer per call (the
    reference's semantics: outputs are freshly allocated, subbox.py:168-170); a caller who drops or
    overwrites the previous result before the next call reuses the same pinned memory at no cost."""

    def __init__(self):
        self.entries = []      # [key, tensor, weakref-to-array | None, owned range]

    def take(self, key, shape, np_dtype, own_range):
        import weakref
        for e in self.entries:
            if e[0] == key and (e[2] is None or e[2]() is None):
                arr = e[1].numpy()
                if e[3] != own_range:      # voxels outside the owned range must read zero
                    arr.fill(0)
                e[2], e[3] = weakref.ref(arr), own_range
                return arr
        t, arr = _pinned_zeros(shape, np_dtype)
        self.entries.append([key, t, weakref.ref(arr), own_range])
        # drop recyclable buffers of other shapes / dtypes
        self.entries = [e for e in self.entries if e[0] == key or (e[2] is not None and e[2]() is not None)]
        return arr


class SubboxProcessor:
    """Unified subbox processor for all four model variants (subbox.py:99-137)."""

    def __init__(self, model, params, config: SubboxConfig):
        from .models import (NBodyEmulatorCore, NBodyEmulatorVelCore, StyleNBodyEmulatorCore,
                             StyleNBodyEmulatorVelCore)
        self.model = model
        self.params = params          # read at call time; tests re-assign it
        self.config = config
        t = type(model)
        if t in (NBodyEmulatorCore, NBodyEmulatorVelCore):
            self.premodulate = True
        elif t in (StyleNBodyEmulatorCore, StyleNBodyEmulatorVelCore):
            self.premodulate = False
        if t in (NBodyEmulatorVelCore, StyleNBodyEmulatorVelCore):
            self.compute_vel = True
        elif t in (NBodyEmulatorCore, StyleNBodyEmulatorCore):
            self.compute_vel = False
        self._tables = None
        self._pool = _OutputPool()
        self._pinned_in = None     # (key, array, registered-by-us) of the host input currently page-locked
        self.last_gather = None    # {"bytes", "ms", "GBps"} of the most recent NCCL output gather
        self.time_gather = False   # True: align the ranks with a barrier first, so "ms" is the collective alone

    def close(self):
        """Release the page-lock on the cached input buffer (also called on garbage collection: a
        host array must never be freed while it is still registered with CUDA)."""
        pin, self._pinned_in = getattr(self, "_pinned_in", None), None
        if pin is not None and pin[2]:
            try:
                import ctypes as C
                eng = Engine.get()
                eng.lib.nbe_host_unregister(eng.h, C.c_void_p(pin[0][0]))
            except Exception:
                pass

    def __del__(self):
        self.close()

    def _pin_input(self, eng, box):
        """Page-lock the caller's array in place so that the upload is a single asynchronous DMA;
        cached for repeated calls on the same buffer (which is kept alive while registered)."""
        import ctypes as C
        key = (box.ctypes.data, box.nbytes)
        if self._pinned_in is not None and self._pinned_in[0] == key:
            return
        if self._pinned_in is not None:
            if self._pinned_in[2]:
                eng.lib.nbe_host_unregister(eng.h, C.c_void_p(self._pinned_in[0][0]))
            self._pinned_in = None
        rc = eng.lib.nbe_host_register(eng.h, C.c_void_p(box.ctypes.data), box.nbytes)
        if rc >= 0:
            self._pinned_in = (key, box, rc == 1)

    def merged_config(self, merge):
        """Halo-amortising tiling: the same box cut into `merge`-times larger
        subboxes.  VALID convolutions are translation-consistent and every output voxel sees the
        same sequence of MMAs, so the result is bit-identical to the reference decomposition while
        the halo recompute drops from ((c+96)/c)^3 to the larger tiles' ratio."""
        cfg = self.config
        merge = tuple(int(m) for m in merge)
        if any(n % m for n, m in zip(cfg.ndiv, merge)):
            raise ValueError(f"merge {merge} must divide ndiv {tuple(cfg.ndiv)}")
        key = ("merged", merge)
        if getattr(self, "_merged", None) is None or self._merged[0] != key:
            m = SubboxConfig(size=cfg.size, ndiv=tuple(n // k for n, k in zip(cfg.ndiv, merge)), dtype=cfg.dtype,
                             output_dtype=cfg.output_dtype, in_chan=cfg.in_chan, padding=cfg.padding)
            if m.crop_size != tuple(c * k for c, k in zip(cfg.crop_size, merge)):
                raise ValueError("merge is only exact when ndiv divides the box size")
            self._merged = (key, m, m.flat_tables())
        return self._merged[1], self._merged[2]

    def _engines(self, devices, dist):
        torch = _torch()
        if not torch.cuda.is_available():
            Engine.get()                      # raises the "no CPU fallback" error
        if devices is None or devices == "all":
            if dist is not None and dist.get_world_size() > 1:
                devices = [torch.cuda.current_device()]       # one process per GPU: this rank's device
            else:
                devices = list(range(torch.cuda.device_count()))
        elif isinstance(devices, int):
            devices = [devices]
        devices = [int(d) for d in devices]
        if not devices or len(set(devices)) != len(devices):
            raise ValueError(f"devices must be distinct CUDA device indices, got {devices}")
        return [Engine.get(d) for d in devices]

    def process_box(self, input_box, z, Om, desc="Processing subboxes", show_progress=True,
                    shard=None, gather="all", copy=None, merge=None, out=None, devices=None):
        """Process the whole box; returns displacement (C,D,H,W) or (displacement, velocity) as
        numpy arrays of ``config.output_dtype`` (reference signature: subbox.py:139-147).

        devices: CUDA devices this call drives.  None (default) = every visible GPU when the process is
        alone, this rank's current device under ``torch.distributed`` (one process per GPU).  With
        several devices the subboxes are cut into contiguous shares, one host thread per GPU
        (``nbe_process_box_multi``): all GPUs read their (D, H) window from the one page-locked input
        and write their finished blocks into the one output box -- nothing is duplicated on the host.
        Returned arrays: page-locked memory the GPUs wrote directly; they belong to the caller and are
        never touched again while referenced (``_OutputPool``); ``copy=True`` forces an ordinary
        pageable numpy copy instead.
        shard: None = shard over torch.distributed ranks iff initialised with world_size > 1;
        (rank, world) forces a split (each rank returns only its own voxels, rest zero) unless
        ``gather`` is "all" / "rank0": then the blocks are all-gathered on the DEVICES over NCCL
        (NVLink), re-assembled there and copied to the host once (gloo / CPU tensors in the tests: a
        sum of the disjoint boxes).  With gather="rank0" only rank 0 returns arrays (others: None).
        merge: e.g. (2, 2, 1) processes 2x2x1 reference subboxes as one larger tile (bit-identical
        output, fewer halo FLOPs, more activation memory).
        out: (disp, vel) or disp -- caller-provided C-contiguous arrays of shape (C,)+size and
        output_dtype (e.g. np.memmap for boxes that do not fit in RAM; input_box may be a memmap
        too).  Only the voxels owned by this call's subboxes are written; nothing is zeroed.
        """
        cfg = self.config
        tables = None
        if merge is not None and tuple(merge) != (1, 1, 1):
            cfg, tables = self.merged_config(merge)
        torch = _torch()
        if self.params is None:
            raise ValueError("No parameters loaded. Use load_params=True in create_emulator.")
        in_np = np.dtype(cfg.dtype)
        out_np = np.dtype(cfg.output_dtype)
        box = np.asarray(input_box)
        if box.shape != (cfg.in_chan,) + tuple(cfg.size):
            raise ValueError(f"input_box has shape {box.shape}, expected {(cfg.in_chan,) + tuple(cfg.size)}")
        # the reference casts each crop to config.dtype on the host before upload (subbox.py:200-202)
        box = np.ascontiguousarray(box, dtype=in_np)
        Dz = np.float32(growth_factor(z, Om))
        vf = np.float32(vel_norm(z, Om)) if self.compute_vel else np.float32(0)

        n = int(cfg.n_subboxes)
        dist = torch.distributed if torch.distributed.is_available() and torch.distributed.is_initialized() else None
        if shard is None:
            shard = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.get_world_size() > 1 else (0, 1)
        rank, world = shard
        lo, hi = shard_range(n, rank, world)

        engines = self._engines(devices, dist)
        fp = params_fingerprint(self.params)           # once per call, shared by all GPUs' engines
        for eng in engines:
            eng.set_precision(self.model.precision)
            eng.set_params(self.params, self.premodulate, self.compute_vel, self.model.eps, fingerprint=fp)
            eng.modulate(None if self.premodulate else np.float32(Om), Dz)
        eng = engines[0]
        if tables is None:
            if self._tables is None:
                self._tables = cfg.flat_tables()
            tables = self._tables
        crop_idx, add0, plen = tables
        shape = (cfg.in_chan,) + tuple(cfg.size)
        self._pin_input(eng, box)

        nccl_gather = (world > 1 and dist is not None and gather in ("all", "rank0") and out is None
                       and dist.get_backend() == "nccl")
        if nccl_gather:
            if len(engines) != 1:
                raise ValueError("the NCCL output gather runs one process per GPU: pass a single device")
            return self._process_gather_nccl(dist, cfg, eng, box, in_np, out_np, plen, crop_idx, add0, lo, hi, n, rank,
                                             world, Dz, vf, gather, copy)

        if out is not None:
            outs = out if isinstance(out, (tuple, list)) else (out,)
            if len(outs) != (2 if self.compute_vel else 1):
                raise ValueError("out must be (disp, vel) for velocity models and disp otherwise")
            for a in outs:
                if a.shape != shape or a.dtype != out_np or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError(f"out arrays must be C-contiguous {shape} of {out_np}")
            if world > 1 and dist is not None and gather in ("all", "rank0"):
                raise ValueError("out= receives only this rank's voxels: use gather='none' with out=")
            dis, vel = outs[0], (outs[1] if self.compute_vel else None)
            copy = False
        else:
            dis = self._pool.take(("d", shape, out_np.name), shape, out_np, (lo, hi))
            vel = self._pool.take(("v", shape, out_np.name), shape, out_np, (lo, hi)) if self.compute_vel else None
        bar = None
        if show_progress:
            from tqdm import tqdm
            bar = tqdm(total=hi - lo, desc=desc, ncols=80,
                       bar_format='{desc}: {percentage:3.0f}%|{bar:30}| {n_fmt}/{total_fmt} [{elapsed}<{remaining}]')
        args = (box, dtype_code(in_np), cfg.size, cfg.crop_size, plen, crop_idx, add0, lo, hi - lo, Dz, vf, dis, vel,
                dtype_code(out_np))
        if len(engines) > 1:
            Engine.process_box_multi(engines, *args)
        else:
            eng.process_box(*args)
        if bar is not None:
            bar.update(hi - lo)
            bar.close()
        if copy:
            dis = np.array(dis)
            vel = np.array(vel) if vel is not None else None
        if world > 1 and dist is not None and gather in ("all", "rank0"):
            dis, vel = _gather_outputs(dist, cfg, dis, vel, world, gather)
        if self.compute_vel:
            return dis, vel
        return dis

    def process_box_streamed(self, plane_source, slab_sink, z, Om, devices=None, max_slabs=None):
        """Boxes that fit neither host nor device memory (BASELINE config 5: 2048^3 = 103 GB in, 206 GB
        out): the box is walked as ndiv[0] D-slabs of subboxes.  For slab k the caller's
        ``plane_source(planes, out)`` fills ``out[:, j]`` (shape (C, len(planes), S1, S2), config.dtype)
        with input plane ``planes[j]`` -- the slab's crop[0] planes plus the periodic 48-plane halo on
        both sides -- into page-locked staging memory; every GPU pulls its (D, H) window straight from
        that staging buffer (``nbe_process_box_multi``) and DMAs its finished blocks into a page-locked
        output slab; ``slab_sink(k, d0, disp, vel)`` then receives (C, crop[0], S1, S2) views (valid only
        during the call; ``vel`` is None for displacement-only models) of planes [d0, d0 + crop[0]).
        Staging is double-buffered: while the GPUs run slab k, a second host thread fills slab k + 1 and
        the sink drains slab k - 1.  Host memory: 2 x (crop[0]+96 input planes + crop[0] output planes per
        field), whatever the box size.  The subboxes, their tables and their results are exactly
        those of ``process_box`` (same kernels, same order): outputs are bit-identical.
        Returns the number of slabs processed."""
        from concurrent.futures import ThreadPoolExecutor
        cfg = self.config
        torch = _torch()
        if self.params is None:
            raise ValueError("No parameters loaded. Use load_params=True in create_emulator.")
        in_np, out_np = np.dtype(cfg.dtype), np.dtype(cfg.output_dtype)
        S0, S1, S2 = (int(v) for v in cfg.size)
        n0, n1, n2 = (int(v) for v in cfg.ndiv)
        c0 = int(cfg.crop_size[0])
        p_lo, p_hi = (int(v) for v in cfg.padding[0])
        P0 = c0 + p_lo + p_hi
        Dz = np.float32(growth_factor(z, Om))
        vf = np.float32(vel_norm(z, Om)) if self.compute_vel else np.float32(0)
        dist = torch.distributed if torch.distributed.is_available() and torch.distributed.is_initialized() else None
        engines = self._engines(devices, dist)
        for eng in engines:
            eng.set_precision(self.model.precision)
            eng.set_params(self.params, self.premodulate, self.compute_vel, self.model.eps)
            eng.modulate(None if self.premodulate else np.float32(Om), Dz)
        if self._tables is None:
            self._tables = cfg.flat_tables()
        crop_idx, add0, plen = self._tables
        per = sum(plen)
        # slab-local tables: the staging buffer holds the slab's planes in gather order, so the D table
        # is the identity and the D anchor is 0; H and W tables are those of the first slab's subboxes
        nsl = n1 * n2
        tab = crop_idx.reshape(-1, per)[:nsl].copy()
        tab[:, :plen[0]] = np.arange(P0, dtype=np.int32)[None]
        anc = add0.reshape(-1, 3)[:nsl].copy()
        anc[:, 0] = 0
        tab, anc = np.ascontiguousarray(tab.ravel()), np.ascontiguousarray(anc.ravel())
        C_ = int(cfg.in_chan)
        nf = 2 if self.compute_vel else 1
        inb = [_pinned_zeros((C_, P0, S1, S2), in_np) for _ in range(2)]
        outb = [[_pinned_zeros((C_, c0, S1, S2), out_np) for _ in range(nf)] for _ in range(2)]
        n_slabs = n0 if max_slabs is None else min(n0, int(max_slabs))

        def fill(k):
            planes = (k * c0 - p_lo + np.arange(P0)) % S0
            plane_source(planes, inb[k % 2][1])

        def run(k):
            o = outb[k % 2]
            Engine.process_box_multi(engines, inb[k % 2][1], dtype_code(in_np), (P0, S1, S2), cfg.crop_size, plen, tab, anc,
                                     0, nsl, Dz, vf, o[0][1], o[1][1] if self.compute_vel else None, dtype_code(out_np),
                                     out_size0=c0)

        def sink(k):
            o = outb[k % 2]
            slab_sink(k, k * c0, o[0][1], o[1][1] if self.compute_vel else None)

        with ThreadPoolExecutor(max_workers=1) as ex:
            fill(0)
            for k in range(n_slabs):
                fut = ex.submit(run, k)
                if k + 1 < n_slabs:
                    fill(k + 1)
                if k > 0:
                    sink(k - 1)
                fut.result()
            sink(n_slabs - 1)
        return n_slabs

    def _process_gather_nccl(self, dist, cfg, eng, box, in_np, out_np, plen, crop_idx, add0, lo, hi, n, rank, world,
                             Dz, vf, gather, copy):
        """One process per GPU: this rank's subboxes stay on its GPU as (count, 3, c0, c1, c2) records,
        ncclAllGather moves every record once over NVLink, the box is re-assembled on the device by a
        strided copy and goes to the host in one DMA."""
        torch = _torch()
        dev = torch.device("cuda", eng.device)
        tdt = {"float32": torch.float32, "float16": torch.float16}[out_np.name]
        c = tuple(int(v) for v in cfg.crop_size)
        nd = tuple(int(v) for v in cfg.ndiv)
        n_max = -(-n // world)
        rec = (3,) + c
        nf = 2 if self.compute_vel else 1
        # [field][rank-local record]: one send buffer for both fields => a single collective
        send = torch.empty((nf, n_max) + rec, dtype=tdt, device=dev)
        if hi - lo < n_max:
            send[:, hi - lo:].zero_()
        eng.process_box_blocks(box, dtype_code(in_np), cfg.size, cfg.crop_size, plen, crop_idx, add0, lo, hi - lo, Dz, vf,
                               send[0], send[1] if self.compute_vel else None, dtype_code(out_np))
        recv = torch.empty((world, nf, n_max) + rec, dtype=tdt, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.time_gather:       # ranks finish their subboxes at different times: without this the interval
            dist.barrier()         # measures the wait for the slowest rank, not the transfer
            torch.cuda.synchronize(dev)
        e0.record()
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
        e1.record()
        outs = [None, None]
        if gather == "all" or rank == 0:
            shape = (cfg.in_chan,) + tuple(cfg.size)
            for f in range(nf):
                if n % world == 0:
                    blocks = recv[:, f].reshape((n,) + rec)
                else:
                    blocks = torch.cat([recv[r, f, :shard_range(n, r, world)[1] - shard_range(n, r, world)[0]]
                                        for r in range(world)])
                full = blocks.view(nd + rec).permute(3, 0, 4, 1, 5, 2, 6).reshape(3, nd[0] * c[0], nd[1] * c[1], nd[2] * c[2])
                host = self._pool.take(("d" if f == 0 else "v", shape, out_np.name), shape, out_np, (0, n))
                ht = torch.from_numpy(host)
                if tuple(full.shape) == shape:
                    ht.copy_(full, non_blocking=True)
                else:        # remainder strips of a non-divisible box stay zero
                    ht[:, :full.shape[1], :full.shape[2], :full.shape[3]].copy_(full)
                outs[f] = host
            torch.cuda.synchronize(dev)
            if copy:
                outs = [np.array(a) if a is not None else None for a in outs]
        else:
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        nbytes = recv.numel() * recv.element_size()
        link = nbytes * (world - 1) // world       # bytes this rank receives from its peers over NVLink
        self.last_gather = {"bytes": int(nbytes), "ms": float(ms), "GBps": nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None,
                            "nvlink_bytes_in_per_rank": int(link),
                            "nvlink_GBps_in_per_rank": link / (ms * 1e-3) / 1e9 if ms > 0 else None,
                            "ranks_aligned": bool(self.time_gather),
                            "collective": "ncclAllGather of (subbox, 3, c0, c1, c2) records, device to device"}
        if self.compute_vel:
            return outs[0], outs[1]
        return outs[0]


def _gather_outputs(dist, cfg, dis, vel, world, mode):
    """Host / gloo fallback of the output gather (CPU tests; NCCL runs _process_gather_nccl): sum the
    disjoint per-rank outputs -- each voxel is owned by exactly one rank, the others hold zeros.  The
    reduction runs on a private copy, so the caller's (possibly cached) arrays are never clobbered."""
    torch = _torch()
    on_gpu = dist.get_backend() == "nccl"
    outs = []
    for a in (dis, vel):
        if a is None:
            outs.append(None)
            continue
        t = torch.from_numpy(a).clone()
        if on_gpu:
            t = t.cuda()
        if mode == "all":
            dist.all_reduce(t)
        else:
            dist.reduce(t, dst=0)
        outs.append(t.cpu().numpy())
    return outs[0], outs[1]
