"""Developer tool (N GPUs, torchrun): the env-sharded solve + the optional NCCL all_gather of the wrenches against the
same batch solved on one GPU.  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/nccl_gather_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import biped_pympc_b200 as bp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
N = 4099                                       # not divisible by the world size
inp, rot = bp.workload.make_batch(N, seed=77)  # every rank generates the same full batch
full = [torch.as_tensor(a, device=dev) for a in inp]
rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
ctx = bp.Context(N, dev.index)
solver = bp.ShardedSolver(N, solve_fn=lambda ins, r: ctx.solve(ins, r, num_calls=4, iters_per_call=5))
w_local = solver.solve(solver.local(full), rot_t[solver.lo:solver.hi])
w_all = solver.gather(w_local)
w_ref = ctx.solve(full, rot_t, num_calls=4, iters_per_call=5)
ok = torch.equal(w_all, w_ref)
print(f"rank {rank}/{world}: shard [{solver.lo}, {solver.hi})  gathered {tuple(w_all.shape)}  bit-identical to the single-GPU solve: {ok}", flush=True)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
