"""Developer tool: distribution of the host entry's wall time per call (4 096 envs x 5 iterations, pinned buffers)."""
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
ctx = bp.Context(B, 0)
tick = ctx.bind_host(hi, hr, hw)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); tick(1, 5, 0.0); ts.append(1e3 * (time.perf_counter() - t0))
print("first 8 calls ms:", " ".join("%.3f" % t for t in ts[:8]))
s = sorted(ts[10:])
print("calls 10..299: mean %.4f median %.4f p90 %.4f p99 %.4f max %.4f" % (sum(s) / len(s), s[len(s) // 2], s[int(0.9 * len(s))], s[int(0.99 * len(s))], s[-1]))
