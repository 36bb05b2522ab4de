"""Developer tool: time the PDIPM kernel alone (4096 envs x 5 iterations) with CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import biped_pympc_b200 as bp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda:0"
inp, _ = bp.workload.make_batch(B, seed=20261018)
ins = [torch.as_tensor(a, device=dev) for a in inp]
Hv, f, Av, b, Gv, d = bp.qp_form(ins)
outs = [torch.empty((B, n), dtype=torch.float64, device=dev) for n in bp._native.IP_NNZ_OUT]
work = torch.empty((B, bp._native.WORK_DOUBLES), dtype=torch.float64, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ts = []
for i in range(23):
    flush.fill_(0.0)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=iters, cold_start=True, stage_invariant=True, out=outs, work=work)
    e.record(); e.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(e))
ts.sort()
print(os.environ.get("SRBD_MPC_LIB", "default"), "pdipm ms: median %.4f min %.4f" % (ts[len(ts) // 2], ts[0]), "checksum %.15e" % float(outs[0].sum()))
