"""Generates tests/golden/srbd_golden_v1.npz.

The reference ships no golden vectors for this path (test/ is empty; CasADi cannot be
imported here), so these fixtures are minted from the oracle restatement
(oracle/srbd_oracle.py) -- inputs from the reference's only fixture (the nominal standing
pose of srbd_constraints.py:243-288) plus seeded synthetic walking states.  Outputs:
  qp_*      : qp_former outputs, float64 oracle
  ip5_*     : 5 PDIPM iterations from the reference cold start, evaluated in long double
              (x87 80-bit) with the reference's full-KKT LDL^T semantics, rounded to float64
  ip20_*    : same for 20 iterations (the controller's 4 x 5)
  ip5f_*    : 5 iterations, float64 oracle (shows the float64 round-off of the reference's
              own formulation next to the long-double values)
Run:  python tests/golden/make_golden.py     (about 3 minutes)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import srbd_oracle as so  # noqa: E402

spec = importlib.util.spec_from_file_location("wl", os.path.join(ROOT, "biped-pympc_b200", "workload.py"))
wl = importlib.util.module_from_spec(spec)
spec.loader.exec_module(wl)


def main():
    nominal = wl.nominal_standing_inputs()
    walk, rot = wl.make_batch(5, seed=20261018)
    stand, _ = wl.make_batch(1, seed=7, standing=True)
    inputs = [np.concatenate([a, b, c], 0) for a, b, c in zip(nominal, walk, stand)]
    B = inputs[0].shape[0]
    out = {f"in_{n}": a for n, a in zip(wl.INPUT_NAMES, inputs)}
    names_qp = ("H", "f", "A", "b", "G", "d")
    qp = [[] for _ in range(6)]
    ip = {k: [[] for _ in range(6)] for k in ("ip5", "ip20", "ip5f")}
    for e in range(B):
        o = so.qp_former(*[a[e] for a in inputs])
        for i in range(6):
            qp[i].append(o[i])
        Hv, f, Av, b, Gv, d = o
        x, s, z, y = so.cold_start(Gv, d)
        for key, iters, dt in (("ip5", 5, np.longdouble), ("ip20", 20, np.longdouble), ("ip5f", 5, np.float64)):
            r = so.pdipm_iterations(Hv, Gv, Av, f, d, b, x, s, z, y, num_iters=iters, dtype=dt)
            for i in range(6):
                ip[key][i].append(np.asarray(r[i], np.float64))
        print("env", e, "done", flush=True)
    for n, v in zip(names_qp, qp):
        out[f"qp_{n}"] = np.stack(v)
    for key, lists in ip.items():
        for n, v in zip(("x", "s", "z", "y", "res", "mu"), lists):
            out[f"{key}_{n}"] = np.stack(v)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "srbd_golden_v1.npz"), **out)
    print("saved", B, "envs")


if __name__ == "__main__":
    main()
