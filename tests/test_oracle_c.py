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
