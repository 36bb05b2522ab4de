"""Shared helpers of the test-suite (oracle evaluation over a batch, emulator binding)."""
import ctypes
import os
import subprocess

import numpy as np

import biped_pympc_b200 as bp
from oracle import srbd_oracle as so

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
QP_OUT_NAMES = ("H", "f", "A", "b", "G", "d")
IP_OUT_NAMES = ("x", "s", "z", "y", "res", "mu")


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_qp_batch(inputs17):
    B = inputs17[0].shape[0]
    outs = [[] for _ in range(6)]
    for e in range(B):
        o = so.qp_former(*[a[e] for a in inputs17])
        for i in range(6):
            outs[i].append(o[i])
    return [np.ascontiguousarray(np.stack(o)) for o in outs]


def oracle_ip_batch(qp, iters, x=None, s=None, z=None, y=None, dtype=np.float64, structured=False):
    """qp = [H, f, A, b, G, d] batch arrays; cold start when x is None."""
    from oracle import structured_proto as sp
    Hv, f, Av, b, Gv, d = qp
    B = Hv.shape[0]
    outs = [[] for _ in range(6)]
    for e in range(B):
        if x is None:
            x0, s0, z0, y0 = so.cold_start(Gv[e], d[e])
        else:
            x0, s0, z0, y0 = x[e], s[e], z[e], y[e]
        if structured:
            r = sp.pdipm_iterations_structured(Hv[e], Gv[e], Av[e], f[e], d[e], b[e], x0, s0, z0, y0, num_iters=iters)
        else:
            r = so.pdipm_iterations(Hv[e], Gv[e], Av[e], f[e], d[e], b[e], x0, s0, z0, y0, num_iters=iters, dtype=dtype)
        for i in range(6):
            outs[i].append(np.asarray(r[i], np.float64).reshape(-1))
    return [np.stack(o) for o in outs]


# ---- host emulation of the device code ------------------------------------------------------
_EMUL = None


def emul_lib():
    global _EMUL
    if _EMUL is not None:
        return _EMUL
    src = os.path.join(ROOT, "tests", "emul", "pdipm_emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "libpdipm_emul.so")
    csrc = os.path.join(ROOT, "biped-pympc_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", csrc, src, "-o", out],
                       check=True)
    _EMUL = ctypes.CDLL(out)
    return _EMUL


def emul_pdipm(qp, iters, x=None, s=None, z=None, y=None, tol=0.0, uniform=True, lane_order=1, precision=0):
    """Runs the PDIPM *device code* under the CPU warp emulator.  Returns (outs6, iters_done).
    precision: PdipmArgs::precision of the fast kernel (0 fp64; 1 / 2 / 3 = fp32 pivot blocks + 0 / 1 / 2 refinement steps).
    uniform: False = general kernel (srbd_pdipm_core.h, arbitrary values on the pattern),
             True / 2 = stage-invariant fast kernel (srbd_pdipm_fast.h),
             1 = stage-invariant instantiation of the general kernel."""
    if uniform is True:
        uniform = 2
    lib = emul_lib()
    Hv, f, Av, b, Gv, d = qp
    B = Hv.shape[0]
    cold = x is None
    zeros = lambda n: np.zeros((B, n))
    ins = [Hv, Gv, Av, f, d, b, zeros(240) if cold else x, zeros(160) if cold else s,
           zeros(160) if cold else z, zeros(140) if cold else y]
    ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
    outs = [np.zeros((B, n)) for n in (240, 160, 160, 140, 4, 1)]
    work = np.zeros((B, bp._native.WORK_DOUBLES))   # = PDIPM_WORK_DOUBLES (srbd_args.h)
    itd = np.zeros(B, np.int32)
    P = ctypes.POINTER(ctypes.c_double)
    inarr = (P * 10)(*[a.ctypes.data_as(P) for a in ins])
    outarr = (P * 6)(*[a.ctypes.data_as(P) for a in outs])
    lib.srbd_emul_set_order(int(lane_order))
    lib.srbd_emul_set_f32_factor(int(precision))
    lib.srbd_emul_pdipm(inarr, outarr, work.ctypes.data_as(P), itd.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                        B, int(iters), ctypes.c_double(tol), int(cold), int(uniform), 0)
    lib.srbd_emul_set_f32_factor(0)
    return outs, itd


def emul_tick(inputs17, rot, iters, tol=0.0, lane_order=1, pair=False, warm_from=None, reset=None, work=None):
    """The fused tick (assembly from the 17 qp_former inputs + PDIPM + wrench) of the fast kernel under
    the CPU warp emulator.  inputs17[1] / [2] may be None.  Returns (outs6, wrench (B,2,6) fp32).
    warm_from = (x, y): warm mode 1; (x, y, z): warm mode 2 (stage-shifted; `work` must then be the array a previous
    call used, which holds that tick's contact schedule)."""
    lib = emul_lib()
    B = inputs17[0].shape[0]
    ins = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in inputs17]
    rot32 = np.ascontiguousarray(rot, dtype=np.float32)
    outs = [np.zeros((B, n)) for n in (240, 160, 160, 140, 4, 1)]
    mode = 0
    if warm_from is not None:                 # the previous tick's solution, found in the output buffers
        outs[0][:] = warm_from[0]; outs[3][:] = warm_from[1]
        mode = 1
        if len(warm_from) == 3:
            outs[2][:] = warm_from[2]
            mode = 2
    rst = None if reset is None else np.ascontiguousarray(reset, dtype=np.int32)
    if work is None:
        work = np.zeros((B, bp._native.WORK_DOUBLES))
    assert work.shape == (B, bp._native.WORK_DOUBLES) and work.flags.c_contiguous
    wrench = np.zeros((B, 2, 6), np.float32)
    itd = np.zeros(B, np.int32)
    P = ctypes.POINTER(ctypes.c_double)
    inarr = (P * 17)(*[a.ctypes.data_as(P) if a is not None else None for a in ins])
    outarr = (P * 6)(*[a.ctypes.data_as(P) for a in outs])
    lib.srbd_emul_set_order(int(lane_order))
    lib.srbd_emul_tick(inarr, rot32.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), outarr, work.ctypes.data_as(P),
                       wrench.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), itd.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                       B, int(iters), ctypes.c_double(tol), int(pair), int(mode),
                       rst.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if rst is not None else None)
    lib.srbd_emul_set_order(1)
    return outs, wrench


_STACK_EMUL = None


def stack_emul_lib():
    """Host build of the small fp32 kernels around the MPC tick (csrc/srbd_stack_impl.h -> tests/emul/stack_emul.cpp)."""
    global _STACK_EMUL
    if _STACK_EMUL is not None:
        return _STACK_EMUL
    src = os.path.join(ROOT, "tests", "emul", "stack_emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "libstack_emul.so")
    csrc = os.path.join(ROOT, "biped-pympc_b200", "csrc")
    deps = [src, os.path.join(csrc, "srbd_stack_impl.h"), os.path.join(csrc, "srbd_t1_impl.h"), os.path.join(ROOT, "include", "srbd_mpc.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", csrc, src,
                        "-o", out], check=True)
    _STACK_EMUL = ctypes.CDLL(out)
    return _STACK_EMUL
