"""Developer tool: where the end-to-end time of the host entry goes (4 096 envs x 5 iterations): wall clock of the device-resident
call + synchronize, of the host entry, and of the host entry with its knobs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import biped_pympc_b200 as bp
B = 4096
inp, rot = bp.workload.make_batch(B, seed=20261018)
hi = [None if i in (1, 2) else torch.as_tensor(a).clone().pin_memory() for i, a in enumerate(inp)]
hr = torch.as_tensor(rot, dtype=torch.float32).clone().pin_memory()
hw = torch.empty((B, 2, 6), dtype=torch.float32).pin_memory()
di = [None if t is None else t.cuda() for t in hi]
dr = hr.cuda()
ctx = bp.Context(B, 0)
def wall(fn, n=200):
    for _ in range(20): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort(); return 1e3 * ts[n // 2]
def dev():
    ctx.solve(di, dr); torch.cuda.synchronize()
print("knobs", {k: v for k, v in os.environ.items() if k.startswith("SRBD_HOST")})
print("device-resident call + synchronize, wall ms: %.4f" % wall(dev))
print("host entry (pinned), wall ms: %.4f" % wall(lambda: ctx.solve_host(hi, hr, hw)))
