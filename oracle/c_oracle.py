"""ctypes binding of the C oracle (oracle/srbd_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib(long_double=False):
    key = "ld" if long_double else "d"
    if key not in _libs:
        path = os.path.join(HERE, "lib", "liboracle_ld.so" if long_double else "liboracle.so")
        src = os.path.join(HERE, "srbd_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
            build()
        _libs[key] = ctypes.CDLL(path)
    return _libs[key]


_P = ctypes.POINTER(ctypes.c_double)


def _arr(arrays):
    arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
    return arrays, (_P * len(arrays))(*[a.ctypes.data_as(_P) for a in arrays])


def qp_former(inputs17, nthreads=0, long_double=False):
    L = lib(long_double)
    fn = L.srbd_oracle_qp_former_ld if long_double else L.srbd_oracle_qp_former
    B = inputs17[0].shape[0]
    ins, inp = _arr(inputs17)
    outs, outp = _arr([np.zeros((B, n)) for n in (240, 240, 1196, 140, 280, 160)])
    fn(inp, outp, B, nthreads)
    return outs


def pdipm(qp, iters, x=None, s=None, z=None, y=None, nthreads=0, long_double=False):
    """qp = [H, f, A, b, G, d] (qp_former output order).  Cold start when x is None."""
    L = lib(long_double)
    fn = L.srbd_oracle_pdipm_ld if long_double else L.srbd_oracle_pdipm
    Hv, f, Av, b, Gv, d = qp
    B = Hv.shape[0]
    cold = x is None
    z0 = lambda n: np.zeros((B, n))
    ins, inp = _arr([Hv, Gv, Av, f, d, b, z0(240) if cold else x, z0(160) if cold else s,
                     z0(160) if cold else z, z0(140) if cold else y])
    outs, outp = _arr([np.zeros((B, n)) for n in (240, 160, 160, 140, 4, 1)])
    fn(inp, outp, B, int(iters), int(cold), nthreads)
    return outs


def pdipm_tol(qp, max_iters, tol, x=None, s=None, z=None, y=None, nthreads=0, long_double=False):
    """Tolerance mode: env e stops after the first iteration whose (|rx|, |rs|, |re|, mu) are all < tol
    (generate_solver_function.py:187; pre-step residual norms + post-step mu, sparse_pdipm_solver.py:524-530),
    at most max_iters.  Returns (outs6, iters_done)."""
    L = lib(long_double)
    fn = L.srbd_oracle_pdipm_tol_ld if long_double else L.srbd_oracle_pdipm_tol
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                   ctypes.c_int, ctypes.c_int]
    Hv, f, Av, b, Gv, d = qp
    B = Hv.shape[0]
    cold = x is None
    z0 = lambda n: np.zeros((B, n))
    ins, inp = _arr([Hv, Gv, Av, f, d, b, z0(240) if cold else x, z0(160) if cold else s,
                     z0(160) if cold else z, z0(140) if cold else y])
    outs, outp = _arr([np.zeros((B, n)) for n in (240, 160, 160, 140, 4, 1)])
    itd = np.zeros(B, dtype=np.int32)
    fn(inp, outp, itd.ctypes.data_as(ctypes.c_void_p), B, int(max_iters), float(tol), int(cold), nthreads)
    return outs, itd


def kkt_fill():
    return lib().srbd_oracle_kkt_fill()
