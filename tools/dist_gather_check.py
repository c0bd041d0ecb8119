"""N-GPU check of the sharded process_box: every rank computes its contiguous subbox range, the
disjoint outputs are gathered over NCCL (gather="all"), and rank 0 compares with its own
single-rank run of the whole box (bit-exact: same kernels, same tiling tables).
torchrun --nproc-per-node N tools/dist_gather_check.py [box=256] [ndiv=2]"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
box = np.random.default_rng(7).standard_normal((3, S, S, S), dtype=np.float32)
cfg = nb.SubboxConfig(size=(S, S, S), ndiv=(nd, nd, nd))
proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), nb.init_params(42), cfg)
d, v = proc.process_box(box, 0.5, 0.3, show_progress=False, shard=(rank, world), gather="all")
ok = True
if rank == 0:
    d1, v1 = proc.process_box(box, 0.5, 0.3, show_progress=False, shard=(0, 1), gather="none")
    ok = np.array_equal(d, d1) and np.array_equal(v, v1)
    print(f"world {world}: gathered == single-rank: {ok}; |disp| mean {np.abs(d).mean():.4f}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
