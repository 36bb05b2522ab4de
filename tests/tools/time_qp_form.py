"""Developer tool (GPU): bandwidth of the stand-alone assembly kernel over batch sizes.  python tests/tools/time_qp_form.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import biped_pympc_b200 as bp  # noqa: E402

for B in (4096, 16384, 65536, 262144):
    inp, _ = bp.workload.make_batch(B, seed=1)
    ins = [torch.as_tensor(a, device="cuda:0") for a in inp]
    out = bp.qp_form(ins)
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda:0")
    ts = []
    for i in range(23):
        flush.fill_(0.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); bp.qp_form(ins, out=out); b.record(); b.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    print(f"B={B:7d}  {ms * 1e3:8.1f} us   {B * 21664 / ms / 1e6:7.1f} GB/s  ({B * 21664 / ms / 1e6 / 6540.8 * 100:.0f} % of 6540.8)")
