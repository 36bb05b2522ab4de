"""Developer tool (GPU): end-to-end time of srbd_mpc_solve_host from pinned host buffers (the bench's e2e leg) under the
host-entry knobs.  python tests/tools/time_e2e.py   (run with SRBD_HOST_HYBRID / SRBD_HOST_CHUNK set to compare)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import biped_pympc_b200 as bp  # noqa: E402

B = 4096
inp, rot = bp.workload.make_batch(B, seed=20261018)
h_in = [torch.as_tensor(a).pin_memory() for a in inp]
h_in[1] = h_in[2] = None
h_rot = torch.as_tensor(rot, dtype=torch.float32).pin_memory()
h_w = torch.empty((B, 2, 6), dtype=torch.float32).pin_memory()
ctx = bp.Context(B, 0)
for _ in range(20):
    ctx.solve_host(h_in, h_rot, h_w, num_calls=1, iters_per_call=5)
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    ctx.solve_host(h_in, h_rot, h_w, num_calls=1, iters_per_call=5)
    ts.append(time.perf_counter() - t0)
ts.sort()
print("hybrid=%s chunk=%s: median %.1f us  p10 %.1f  p90 %.1f  -> %.2f M solves/s" % (
    os.environ.get("SRBD_HOST_HYBRID", "0"), os.environ.get("SRBD_HOST_CHUNK", "default"), 1e6 * ts[100], 1e6 * ts[20], 1e6 * ts[180], B / ts[100] / 1e6))
