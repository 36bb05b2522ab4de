"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/srbd_mpc.h declares; patterns agree with the oracle; no compute call is made."""
import ctypes
import os
import re

import numpy as np

import biped_pympc_b200 as bp
from oracle import srbd_oracle as so

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "srbd_mpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srbd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = bp._native.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in srbd_mpc.h but not exported"
    assert set(declared) == set(bp._native.EXPORTED_SYMBOLS)
    assert lib.srbd_version() == 100


def test_compat_libraries_export_evaluate():
    """lib<fn>.so::evaluate is what CusadiFunction.py:30-46 binds."""
    libdir = os.path.join(ROOT, "biped-pympc_b200", "lib")
    for name in ("libqp_former.so", "libsparse_pdipm_multiple_iterations.so"):
        lib = ctypes.CDLL(os.path.join(libdir, name))
        assert hasattr(lib, "evaluate")


def test_nnz_tables():
    lib = bp._native.load()
    assert [lib.srbd_qp_former_nnz_in(i) for i in range(17)] == list(so.QP_FORMER_NNZ_IN)
    assert [lib.srbd_qp_former_nnz_out(i) for i in range(6)] == list(so.QP_FORMER_NNZ_OUT)
    assert [lib.srbd_pdipm_nnz_in(i) for i in range(10)] == list(so.PDIPM_NNZ_IN)
    assert [lib.srbd_pdipm_nnz_out(i) for i in range(6)] == list(so.PDIPM_NNZ_OUT)
    assert lib.srbd_qp_former_nnz_in(17) < 0 and lib.srbd_pdipm_nnz_out(-1) < 0
    assert lib.srbd_pdipm_work_doubles(3, 0) == 3 * bp._native.WORK_DOUBLES
    assert lib.srbd_pdipm_work_doubles(3, bp._native.STAGE_INVARIANT) == 3 * bp._native.WORK_DOUBLES_FAST == 3 * 2048


def test_patterns_match_oracle_structure():
    """Index arithmetic of csrc/srbd_pattern.h vs the oracle's independent derivation from
    the symbolic dependency masks."""
    for which, (cp, rw, shape) in {"A": (so.A_COLPTR, so.A_ROW, (140, 240)), "G": (so.G_COLPTR, so.G_ROW, (160, 240)),
                                   "H": (so.H_COLPTR, so.H_ROW, (240, 240))}.items():
        colptr, row, nrow, ncol = bp.pattern(which)
        assert (nrow, ncol) == shape
        assert np.array_equal(colptr, cp) and np.array_equal(row, rw)


def test_error_paths_do_not_exit():
    lib = bp._native.load()
    assert lib.srbd_qp_form(None, None, 4, None) == -1
    assert lib.srbd_pdipm(None, None, None, None, 4, 5, 0.0, 0, None) == -1
    assert lib.srbd_error_string(-1).decode() == "invalid argument"
    h = ctypes.c_void_p()
    assert lib.srbd_ctx_create(ctypes.byref(h), 0, -1) == -1
    res = ctypes.c_int(7)
    assert lib.srbd_pdipm_is_stage_invariant(None, 4, None, ctypes.byref(res)) == -1
    assert lib.srbd_marshal_inputs(None, 13, 0, 0, 40.0, 1.0, None, 4, None) == -1
    assert lib.srbd_mpc_solve(None, None, None, 4, 4, 5, 0.0, None, None, None, None, None, None, None, None) == -1
    assert lib.srbd_mpc_solve_warm(None, None, None, 4, 20, 0.0, 2, None, None, None, None, None, None, None, None, None) == -1
    assert lib.srbd_mpc_solve_host(None, None, None, 4, 4, 5, 0.0, None, None, None) == -1
    assert lib.srbd_wrench(None, None, None, 4, None) == -1
    assert lib.srbd_reference_trajectory(None, 0.01, None, None, None, None, None, 4, None) == -1
    assert lib.srbd_gait(None, None, None, None, 0.001, 1, 10, None, None, None, None, 4, None) == -1
    assert lib.srbd_update_state(None, None, 43, 4, None) == -1
    assert lib.srbd_lowlevel(None, None, None, None, 0.001, 0, 1, 4, None) == -1
    # alignment contract of srbd_qp_form (16-byte vectors): checked before anything touches the device
    PP = ctypes.c_void_p * 17
    ins = PP(*[0x1000 + 0x100 * i for i in range(17)]); outs = (ctypes.c_void_p * 6)(*[0x9000 + 0x100 * i for i in range(6)])
    ins[3] = 0x1308                                       # x_ref at an odd double offset
    assert lib.srbd_qp_form(ins, outs, 4, None) == -1
    empty = bp.biped_controller._LoopBuffers()          # all-null buffers are rejected, not dereferenced
    assert lib.srbd_update_state(ctypes.byref(empty), None, 43, 4, None) == -1


def test_loop_buffers_struct_mirrors_the_header():
    """ctypes mirror of srbd_loop_buffers: same size and the same field names in the same order."""
    lib = bp._native.load()
    assert ctypes.sizeof(bp.biped_controller._LoopBuffers) == lib.srbd_loop_buffers_size()
    text = open(os.path.join(ROOT, "include", "srbd_mpc.h")).read()
    body = text[text.index("typedef struct srbd_loop_buffers {"):text.index("} srbd_loop_buffers;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*([A-Za-z_0-9]+)\s*;", body)
    assert names == [n for n, _ in bp.biped_controller._LoopBuffers._fields_]


def test_function_specs_mirror_reference_signatures():
    q, p = bp.QP_FORMER_FUNCTION, bp.PDIPM_FUNCTION
    assert (q.name(), q.n_in(), q.n_out()) == ("qp_former", 17, 6)
    assert (p.name(), p.n_in(), p.n_out()) == ("sparse_pdipm_multiple_iterations", 10, 6)
    assert [q.nnz_out(i) for i in range(6)] == [240, 240, 1196, 140, 280, 160]
    rows, cols = q.triplet_out(2)
    assert rows.shape == (1196,) and rows.max() == 139 and cols.max() == 239
    rows, cols = q.triplet_out(1)
    assert np.array_equal(rows, np.arange(240)) and not cols.any()
