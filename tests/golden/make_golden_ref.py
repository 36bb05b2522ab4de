"""Generates tests/golden/srbd_ref_shim_v1.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

The reference ships no golden vectors for this path and CasADi is not installable here.  This
script imports the reference's unmodified files

    /root/reference/biped_pympc/casadi/srbd_centroidal_model.py   (dynamics, RK4)
    /root/reference/biped_pympc/casadi/srbd_constraints.py        (qp_former)
    /root/reference/biped_pympc/casadi/sparse_pdipm_solver.py     (sparse_pdipm_multiple_iterations)

with `casadi` resolved to tests/golden/casadi_shim/casadi.py -- a numeric stand-in for the absent
third-party dependency (see its header for what is restated) -- builds the two functions exactly as
srbd_constraints.py:20-81 and generate_solver_function.py:60-114 do (patterns captured from the
qp_former outputs, num_iters=5, beta=1e-8), and evaluates them on:

    env 0      the reference's nominal standing input set (srbd_constraints.py:243-288)
    env 1      BASELINE config 1: single T1 standing (workload.config1_t1_standing)
    env 2..    seeded synthetic walking states of workload.make_batch

following the controller's call sequence (mpc_controller_cusadi.py:138-169): cold start x=0,
s=max(d,1), z=1, y=1, then four chained 5-iteration calls.  Saved: the CCS patterns the reference code
produces, the six qp_former outputs, and the iterates after every call (ip5_* after the first,
ip10_*, ip15_*, ip20_* after the later ones).

It only runs in the build container (needs /root/reference); the fixture travels, the script is
committed next to it.  Run (about a minute each):
    python tests/golden/make_golden_ref.py 22                               -> srbd_ref_shim_v1.npz
    CASADI_SHIM_ORDER=natural python tests/golden/make_golden_ref.py 22     -> srbd_ref_shim_v1_natural.npz
(the second is the LDL elimination-order sensitivity variant, see casadi_shim/casadi.py)
"""
import importlib.util
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"


def import_reference(use_shim=True):
    """reference modules over the shim (or over a real CasADi, should one be importable: use_shim=False),
    without running biped_pympc/__init__.py (which pulls in the whole controller stack)"""
    if use_shim:
        sys.path.insert(0, os.path.join(HERE, "casadi_shim"))
    import casadi  # noqa: F401
    assert ("casadi_shim" in casadi.__file__) == bool(use_shim)
    for name, sub in (("biped_pympc", "biped_pympc"), ("biped_pympc.casadi", "biped_pympc/casadi")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = pkg
    from biped_pympc.casadi.srbd_constraints import SRBD_Constraints
    from biped_pympc.casadi.sparse_pdipm_solver import sparse_pdipm_multiple_iterations
    return SRBD_Constraints, sparse_pdipm_multiple_iterations


def load_workload():
    spec = importlib.util.spec_from_file_location("wl", os.path.join(ROOT, "biped-pympc_b200", "workload.py"))
    wl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wl)
    return wl


def main(n_walk=6, use_shim=True, save=True):
    SRBD_Constraints, sparse_pdipm_multiple_iterations = import_reference(use_shim)
    wl = load_workload()
    t0 = time.time()
    qp_former = SRBD_Constraints(10).qp_mat_former                      # srbd_constraints.py:235-236
    print("qp_former built: %d instructions, %.1f s" % (qp_former.n_instructions(), time.time() - t0), flush=True)

    nominal = wl.nominal_standing_inputs()
    cfg1 = wl.config1_t1_standing()
    walk, _ = wl.make_batch(n_walk, seed=20261018)
    inputs = [np.concatenate([a, b, c], 0) for a, b, c in zip(nominal, cfg1, walk)]
    B = inputs[0].shape[0]

    # generate_solver_function.py:60-76: patterns from a qp_former evaluation
    out0 = qp_former(*[a[0] for a in inputs])
    H, f, A, b, G, d = out0
    H_col, H_row = H.sparsity().get_ccs()
    A_col, A_row = A.sparsity().get_ccs()
    G_col, G_row = G.sparsity().get_ccs()
    nz, num_eq, num_ineq = H.shape[0], A.shape[0], G.shape[0]
    print("dims", nz, num_eq, num_ineq, "nnz H A G", H.nnz(), A.nnz(), G.nnz(), flush=True)

    t0 = time.time()
    solver = sparse_pdipm_multiple_iterations(                         # generate_solver_function.py:106-114
        nz, num_eq, num_ineq, H_row, H_col, A_row, A_col, G_row, G_col, beta=1e-8, num_iters=5)
    print("solver built: %d instructions, %.1f s" % (solver.n_instructions(), time.time() - t0), flush=True)

    res = {"pattern_H_colptr": np.array(H_col, np.int32), "pattern_H_row": np.array(H_row, np.int32),
           "pattern_A_colptr": np.array(A_col, np.int32), "pattern_A_row": np.array(A_row, np.int32),
           "pattern_G_colptr": np.array(G_col, np.int32), "pattern_G_row": np.array(G_row, np.int32),
           "solver_instructions": np.array(solver.n_instructions()),
           "qp_former_instructions": np.array(qp_former.n_instructions())}
    for n, a in zip(wl.INPUT_NAMES, inputs):
        res["in_" + n] = a
    qp_names = ("H", "f", "A", "b", "G", "d")
    ip_names = ("x", "s", "z", "y", "res", "mu")
    acc = {}
    for e in range(B):
        t0 = time.time()
        o = qp_former(*[a[e] for a in inputs])
        vals = [np.array(m.nonzeros()) for m in o]
        for m, ref_pat in zip((o[0], o[2], o[4]), ((H_col, H_row), (A_col, A_row), (G_col, G_row))):
            assert tuple(map(list, m.sparsity().get_ccs())) == tuple(map(list, ref_pat))
        for n, v in zip(qp_names, vals):
            acc.setdefault("qp_" + n, []).append(v)
        Hv, fv, Av, bv, Gv, dv = vals
        # mpc_controller_cusadi.py:138-141
        x = np.zeros(nz)
        s = np.maximum(dv - np.asarray(o[4]) @ x, 1.0)
        z = np.ones(num_ineq)
        y = np.ones(num_eq)
        for call in range(4):                                           # mpc_controller_cusadi.py:144-169
            r = solver(Hv, Gv, Av, fv, dv, bv, x, s, z, y)
            r = [np.asarray(q).reshape(-1) for q in r]
            x, s, z, y = r[0], r[1], r[2], r[3]
            for n, v in zip(ip_names, r):
                acc.setdefault("ip%d_%s" % (5 * (call + 1), n), []).append(v)
        print("env %d: %.1f s   res(20) = %s" % (e, time.time() - t0, r[4]), flush=True)
    for k, v in acc.items():
        res[k] = np.stack(v)
    if not save:
        return res
    order = os.environ.get("CASADI_SHIM_ORDER", "mindeg")
    name = "srbd_ref_shim_v1.npz" if order == "mindeg" else "srbd_ref_shim_v1_%s.npz" % order
    np.savez_compressed(os.path.join(HERE, name), **res)
    print("saved", name, B, "envs")
    return res


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 6)
