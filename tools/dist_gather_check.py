"""torchrun --nproc-per-node N tools/dist_gather_check.py [out.json]
One process per GPU: process a 256^3 box sharded over the ranks with the device-side NCCL gather and
compare, on every rank, with the same box processed by that rank alone.  Driven by
tests/test_gpu_layers_and_host.py::test_nccl_block_gather_two_ranks."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jax_nbody_emulator_with_dj_b200 as nb

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 2
size, ndiv = (S, S, S), (nd, nd, nd)
box = np.random.default_rng(7).standard_normal((3,) + size, dtype=np.float32)
proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), nb.init_params(42), nb.SubboxConfig(size=size, ndiv=ndiv))
d, v = proc.process_box(box, 0.5, 0.3, show_progress=False)                       # sharded + NCCL all-gather (default)
g = dict(proc.last_gather)
proc.time_gather = True
d, v = proc.process_box(box, 0.5, 0.3, show_progress=False)                       # warm: communicator and buffers exist
g = dict(proc.last_gather)
proc.time_gather = False
d1, v1 = proc.process_box(box, 0.5, 0.3, show_progress=False, shard=(0, 1), gather="none")   # this rank alone
ok = torch.tensor([int(np.array_equal(d, d1) and np.array_equal(v, v1))], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
d0, v0 = proc.process_box(box, 0.5, 0.3, show_progress=False, gather="rank0")
ok0 = bool(np.array_equal(d0, d1)) if rank == 0 else (d0 is None)
if rank == 0:
    res = {"world": world, "bit_identical_all_ranks": bool(ok.item()) and ok0, "gather": g}
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
