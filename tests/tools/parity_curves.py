"""Developer tool: per-iteration parity curves (DESIGN.md "Accuracy").  Max over envs of the norm-wise
relative error of x, s, z, y against the 80-bit C oracle, for the double C oracle (the reference
formulation evaluated in fp64) and for the kernel code run under the CPU warp emulator."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import biped_pympc_b200 as bp
from helpers import emul_pdipm
from oracle import c_oracle as co

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
inp, _ = bp.workload.make_batch(B, seed=4242)
qp = co.qp_former(inp)
err = lambda a, b: max(np.linalg.norm(a[e] - b[e]) / np.linalg.norm(b[e]) for e in range(B))
print("iters | reference formulation in fp64 (x s z y) | kernel (x s z y)")
for it in (1, 2, 3, 5, 10, 15, 20, 25, 30):
    l = co.pdipm(qp, it, long_double=True); d = co.pdipm(qp, it); k, _ = emul_pdipm(qp, it)
    print(it, "|", " ".join("%.1e" % err(d[i], l[i]) for i in range(4)), "|", " ".join("%.1e" % err(k[i], l[i]) for i in range(4)), flush=True)
