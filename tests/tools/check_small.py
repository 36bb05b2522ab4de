"""Developer tool: run the PDIPM kernel on a few envs on cuda:0 and print the error against the
long-double C oracle (used with compute-sanitizer / different builds).  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import biped_pympc_b200 as bp
from oracle import c_oracle as co

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
inp, _ = bp.workload.make_batch(B, seed=4242)
qp = co.qp_former(inp)
ref = co.pdipm(qp, iters, long_double=True)
dev = "cuda:0"
Hv, f, Av, b, Gv, d = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in qp]
for si in (True, False):
    out = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=iters, cold_start=True, stage_invariant=si)
    torch.cuda.synchronize()
    errs = []
    for i in range(4):
        o = out[i].cpu().numpy()
        errs.append(max(np.linalg.norm(o[e] - ref[i][e]) / np.linalg.norm(ref[i][e]) for e in range(B)))
    print("stage_invariant", si, "iters", iters, "max rel err x,s,z,y:", ["%.2e" % v for v in errs], flush=True)
