"""The C restatement (oracle/srbd_oracle.c: sparse no-pivot LDL^T of the full KKT) against the
numpy restatement, the long-double build and the committed golden fixtures."""
import numpy as np

import biped_pympc_b200 as bp
from helpers import oracle_ip_batch, oracle_qp_batch, rel_err
from oracle import c_oracle as co
from oracle import srbd_oracle as so


def test_c_patterns_equal_numpy_patterns():
    import ctypes
    lib = co.lib()
    for which, cp, rw in (("A", so.A_COLPTR, so.A_ROW), ("G", so.G_COLPTR, so.G_ROW)):
        a = ctypes.POINTER(ctypes.c_int)(); b = ctypes.POINTER(ctypes.c_int)()
        nnz = lib.srbd_oracle_pattern(ord(which), ctypes.byref(a), ctypes.byref(b))
        assert nnz == len(rw)
        assert np.array_equal(np.ctypeslib.as_array(a, shape=(241,)), cp)
        assert np.array_equal(np.ctypeslib.as_array(b, shape=(nnz,)), rw)
    assert 3000 < co.kkt_fill() < 8000      # fill of the stage-interleaved ordering


def test_c_qp_former_matches_numpy_and_golden(golden):
    inputs = [golden[f"in_{n}"] for n in bp.workload.INPUT_NAMES]
    got = co.qp_former(inputs)
    for n, v in zip("HfAbGd", got):
        ref = golden[f"qp_{n}"]
        assert np.abs(v - ref).max() <= 2e-15 * max(1.0, np.abs(ref).max())
    inp, _ = bp.workload.make_batch(16, seed=123)
    a, b = co.qp_former(inp), oracle_qp_batch(inp)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() <= 2e-15 * max(1.0, np.abs(y).max())


def test_c_pdipm_matches_golden(golden):
    qp = [golden[f"qp_{n}"] for n in "HfAbGd"]
    d5 = co.pdipm(qp, 5)
    l5 = co.pdipm(qp, 5, long_double=True)
    for i, n in enumerate(("x", "s", "z", "y")):
        assert rel_err(l5[i], golden[f"ip5_{n}"]) < 1e-11        # two 80-bit evaluations, different pivot order
        assert rel_err(d5[i], golden[f"ip5_{n}"]) < 1e-9         # float64 round-off of the full-KKT LDL
    l20 = co.pdipm(qp, 20, long_double=True)
    for i, n in enumerate(("x", "s", "z", "y")):
        assert rel_err(l20[i], golden[f"ip20_{n}"]) < 1e-7       # path sensitivity at 20 iterations


def test_c_pdipm_matches_numpy_dense_oracle():
    inp, _ = bp.workload.make_batch(4, seed=77)
    qp = co.qp_former(inp)
    a = co.pdipm(qp, 3)
    b = oracle_ip_batch(qp, 3)
    for i in range(4):
        assert rel_err(a[i], b[i]) < 1e-9
    # warm-started continuation
    c = co.pdipm(qp, 2, x=a[0], s=a[1], z=a[2], y=a[3])
    d = co.pdipm(qp, 5)
    for i in range(6):
        assert np.array_equal(c[i], d[i])


def test_config1_single_t1_standing_oracles_agree():
    """BASELINE configs[0] (single T1 biped, standing, batch 1): numpy restatement == C restatement ==
    80-bit build, and the emulated kernel code reproduces it; the standing solution carries the weight."""
    from helpers import emul_pdipm
    inp = bp.workload.config1_t1_standing()
    qp_c, qp_np = co.qp_former(inp), oracle_qp_batch(inp)
    for a, b in zip(qp_c, qp_np):
        assert np.abs(a - b).max() <= 2e-15 * max(1.0, np.abs(b).max())
    # one iteration: all four evaluations agree.  Later the standing problem drives s to the 1e-8 clamp
    # and the numpy restatement's dense pivoted LU of the full KKT drifts (z: 7e-7 at 5 iterations,
    # 4e-4 at 10) while the no-pivot LDL^T in 64 and 80 bits and the kernel's block elimination stay
    # within 1e-10 of each other up to 20 iterations ("Accuracy")
    l1, n1 = co.pdipm(qp_c, 1, long_double=True), oracle_ip_batch(qp_c, 1)
    for i in range(4):
        assert rel_err(n1[i], l1[i]) < 1e-9
    for it in (5, 20):
        d, l = co.pdipm(qp_c, it), co.pdipm(qp_c, it, long_double=True)
        k, _ = emul_pdipm(qp_c, it)
        for i in range(4):
            assert rel_err(d[i], l[i]) < 1e-9 and rel_err(k[i], l[i]) < 1e-9
    l20 = co.pdipm(qp_c, 20, long_double=True)
    u0 = l20[0][0, 120:132]
    # left / right symmetric; the first-stage vertical force is the weight plus the tracking correction
    assert abs(u0[2] - u0[5]) < 1e-4 * abs(u0[2]) and abs(u0[2] + u0[5] - 40.0 * 9.81) < 0.15 * 40.0 * 9.81


def test_pdipm_converges_to_the_kkt_point_of_the_qp():
    """Independent of any restated solver: a convex QP's KKT conditions are necessary and sufficient.
    On BASELINE config 1 the iterates of the reference algorithm (C oracle, and the kernel code under the
    emulator) reach a point that satisfies them on the qp_former data itself -- stationarity, primal
    feasibility, dual feasibility, complementarity (limited by the 1e-8 clamps and the beta / delta
    regularisation of sparse_pdipm_solver.py:416-439,511-515)."""
    from helpers import emul_pdipm
    inp = bp.workload.config1_t1_standing()
    qp = co.qp_former(inp)
    Hv, f, Av, b, Gv, d = [a[0] for a in qp]
    A = so.dense_from_ccs(Av, so.A_COLPTR, so.A_ROW, (140, 240))
    G = so.dense_from_ccs(Gv, so.G_COLPTR, so.G_ROW, (160, 240))
    for outs in (co.pdipm(qp, 60), emul_pdipm(qp, 60)[0]):
        x, s, z, y = [a[0] for a in outs[:4]]
        assert np.abs(Hv * x + f + G.T @ z + A.T @ y).max() < 1e-5          # stationarity
        assert np.abs(A @ x - b).max() < 1e-9                               # equalities
        assert (G @ x - d).max() < 1e-7 and z.min() >= 0                    # inequalities, dual sign
        assert np.abs(z * (d - G @ x)).max() < 1e-4                         # complementarity
    # and the optimum carries the weight: sum of vertical forces of stage 0 within 15 % of m g
    assert abs(x[122] + x[125] - 40.0 * 9.81) < 0.15 * 40.0 * 9.81


def test_tolerance_mode_of_the_oracle_and_of_the_device_code():
    """generate_solver_function.py:187 as a runtime mode: an env stops after the first iteration whose pre-step residual
    norms and post-step mu are below tol.  (i) stopping at iteration n returns exactly the n-iteration iterate; (ii) the
    kernel's device code (under the warp emulator) stops every env at the oracle's iteration, at different intermediate
    iterations across the batch."""
    import biped_pympc_b200 as bp
    from helpers import emul_pdipm
    inp, _ = bp.workload.make_batch(12, seed=7, standing=True)
    qp = co.qp_former(inp)
    out, it = co.pdipm_tol(qp, 30, 1e-1)
    assert it.min() < 30 and len(set(it.tolist())) >= 3
    for e in (0, 5, 11):
        fixed = co.pdipm([a[e:e + 1] for a in qp], int(it[e]))
        for i in range(5):
            assert np.array_equal(fixed[i][0], out[i][e])
        assert (out[4][e] < 1e-1).all() or it[e] == 30
    capped, itc = co.pdipm_tol(qp, 7, 1e-30)            # never satisfied: the cap
    assert (itc == 7).all() and np.array_equal(capped[0], co.pdipm(qp, 7)[0])
    ref, it_ld = co.pdipm_tol(qp, 30, 1e-1, long_double=True)
    got, itd = emul_pdipm(qp, 30, tol=1e-1, uniform=True)
    assert np.array_equal(itd, it_ld) and np.array_equal(it, it_ld)
    for i in range(4):
        assert rel_err(got[i], ref[i]) < 5e-8
