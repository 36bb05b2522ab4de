"""ctypes binding of libsrbd_mpc.so (the C-ABI declared in include/srbd_mpc.h).

There is no CPU fallback: if the shared object is missing the import of any compute entry
point raises.  torch is used only for device memory and streams."""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRBD_MPC_LIB") or os.path.join(HERE, "lib", "libsrbd_mpc.so")   # env override: developer A/B builds

COLD_START = 1
STAGE_INVARIANT = 2
F32_FACTOR, F32_REFINE, F32_REFINE2 = 4, 8, 16      # optional reduced-precision mode (include/srbd_mpc.h)
PRECISIONS = {"f64": 0, "f32": 1, "f32_refined": 2, "f32_refined2": 3}   # srbd_ctx_set_precision / pdipm(precision=)
_PRECISION_FLAGS = {0: 0, 1: F32_FACTOR, 2: F32_FACTOR | F32_REFINE, 3: F32_FACTOR | F32_REFINE2}

QP_NNZ_IN = (12, 120, 120, 120, 1, 1, 1, 9, 9, 3, 3, 3, 20, 12, 12, 3, 3)
QP_NNZ_OUT = (240, 240, 1196, 140, 280, 160)
IP_NNZ_IN = (240, 280, 1196, 240, 160, 140, 240, 160, 160, 140)
IP_NNZ_OUT = (240, 160, 160, 140, 4, 1)
WORK_DOUBLES = 6224          # general kernel, per env (srbd_args.h PDIPM_WORK_DOUBLES)
WORK_DOUBLES_FAST = 2048     # fast kernel (STAGE_INVARIANT), per env (the fp64 path touches the first 1 664)

_lib = None
_P = ctypes.POINTER(ctypes.c_double)
_PP = ctypes.POINTER(_P)

EXPORTED_SYMBOLS = (
    "srbd_version", "srbd_error_string", "srbd_last_cuda_error", "srbd_init",
    "srbd_qp_former_nnz_in", "srbd_qp_former_nnz_out", "srbd_pdipm_nnz_in", "srbd_pdipm_nnz_out",
    "srbd_pattern", "srbd_loop_buffers_size", "srbd_update_state", "srbd_lowlevel", "srbd_gait", "srbd_reference_trajectory", "srbd_marshal_inputs", "srbd_marshal_inputs_iw", "srbd_qp_form", "srbd_pdipm_work_doubles", "srbd_pdipm", "srbd_pdipm_is_stage_invariant", "srbd_wrench",
    "srbd_ctx_create", "srbd_ctx_destroy", "srbd_ctx_set_precision", "srbd_mpc_solve", "srbd_mpc_solve_warm", "srbd_mpc_solve_host",
    "srbd_qp_former_evaluate", "srbd_sparse_pdipm_multiple_iterations_evaluate", "srbd_dfma_peak",
    "srbd_t1_kinematics", "srbd_t1_inverse_kinematics",
)


class SrbdError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc exists).  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        from . import _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise SrbdError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint
    lib.srbd_version.restype = ci
    lib.srbd_error_string.restype = ctypes.c_char_p
    lib.srbd_error_string.argtypes = [ci]
    lib.srbd_last_cuda_error.restype = ctypes.c_char_p
    lib.srbd_init.argtypes = [ci]
    lib.srbd_pattern.argtypes = [ctypes.c_char, ctypes.POINTER(ctypes.POINTER(ci)), ctypes.POINTER(ctypes.POINTER(ci)),
                                 ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.srbd_qp_form.argtypes = [vp, vp, ci, vp]
    lib.srbd_pdipm_work_doubles.restype = ctypes.c_size_t
    lib.srbd_pdipm_work_doubles.argtypes = [ci, cu]
    lib.srbd_pdipm.argtypes = [vp, vp, vp, vp, ci, ci, cd, cu, vp]
    lib.srbd_pdipm_is_stage_invariant.argtypes = [vp, ci, vp, ctypes.POINTER(ci)]
    lib.srbd_marshal_inputs.argtypes = [vp, ci, ci, ci, cd, cd, vp, ci, vp]
    lib.srbd_marshal_inputs_iw.argtypes = [vp, vp, vp, ci, ci, ci, cd, cd, vp, ci, vp]
    lib.srbd_wrench.argtypes = [vp, vp, vp, ci, vp]
    lib.srbd_loop_buffers_size.restype = ctypes.c_size_t
    lib.srbd_update_state.argtypes = [vp, vp, ci, ci, vp]
    lib.srbd_lowlevel.argtypes = [vp, vp, vp, vp, cd, ci, ci, ci, vp]
    lib.srbd_gait.argtypes = [vp, vp, vp, vp, cd, ci, ci, vp, vp, vp, vp, ci, vp]
    lib.srbd_reference_trajectory.argtypes = [vp, cd, vp, vp, vp, vp, vp, ci, vp]
    lib.srbd_ctx_create.argtypes = [ctypes.POINTER(vp), ci, ci]
    lib.srbd_ctx_destroy.argtypes = [vp]
    lib.srbd_ctx_set_precision.argtypes = [vp, ci]
    lib.srbd_mpc_solve.argtypes = [vp, vp, vp, ci, ci, ci, cd, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.srbd_mpc_solve_warm.argtypes = [vp, vp, vp, ci, ci, cd, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.srbd_mpc_solve_host.argtypes = [vp, vp, vp, ci, ci, ci, cd, vp, vp, vp]
    lib.srbd_qp_former_evaluate.restype = ctypes.c_float
    lib.srbd_qp_former_evaluate.argtypes = [vp, vp, vp, ci]
    lib.srbd_sparse_pdipm_multiple_iterations_evaluate.restype = ctypes.c_float
    lib.srbd_sparse_pdipm_multiple_iterations_evaluate.argtypes = [vp, vp, vp, ci]
    lib.srbd_dfma_peak.argtypes = [ci, ci, ci, ctypes.POINTER(cd)]
    lib.srbd_t1_kinematics.argtypes = [vp, vp, vp, ci, vp]
    lib.srbd_t1_inverse_kinematics.argtypes = [vp, vp, ci, vp]
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        lib = load()
        msg = lib.srbd_error_string(rc).decode()
        cuda = lib.srbd_last_cuda_error().decode()
        raise SrbdError(f"srbd_mpc: {msg} ({rc})" + (f": {cuda}" if cuda else ""))


def pattern(which: str):
    """(colptr, row, nrow, ncol) numpy arrays of the frozen CCS pattern 'H' | 'A' | 'G'."""
    import numpy as np
    lib = load()
    cp = ctypes.POINTER(ctypes.c_int)()
    rw = ctypes.POINTER(ctypes.c_int)()
    nrow, ncol, nnz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib.srbd_pattern(which.encode(), ctypes.byref(cp), ctypes.byref(rw), ctypes.byref(nrow),
                           ctypes.byref(ncol), ctypes.byref(nnz)))
    colptr = np.ctypeslib.as_array(cp, shape=(ncol.value + 1,)).copy()
    row = np.ctypeslib.as_array(rw, shape=(nnz.value,)).copy()
    return colptr, row, nrow.value, ncol.value


# ---------------------------------------------------------------------------------------------
# torch-tensor helpers (device memory + stream plumbing only)
# ---------------------------------------------------------------------------------------------
def _ptr_array(tensors: Sequence, allow_none=False):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if (t is None and allow_none) else t.data_ptr()
    return arr


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_inputs(tensors, nnz, B=None, device_type="cuda"):
    import torch
    if len(tensors) != len(nnz):
        raise ValueError(f"expected {len(nnz)} tensors, got {len(tensors)}")
    out = []
    for i, (t, n) in enumerate(zip(tensors, nnz)):
        if t.dtype != torch.float64:
            raise TypeError(f"argument {i}: expected float64, got {t.dtype}")
        if t.device.type != device_type:
            raise ValueError(f"argument {i}: expected a {device_type} tensor")
        if B is None:
            B = t.shape[0]
        if t.numel() != B * n:
            raise ValueError(f"argument {i}: expected {B}x{n} elements, got shape {tuple(t.shape)}")
        out.append(t if t.is_contiguous() else t.contiguous())
    return out, B


def _check_buffers(what, tensors, nnz, B, dev, dtype=None):
    """Caller-provided output / scratch tensors: dtype, device, contiguity and room for B envs.  The kernels write
    B * nnz elements whatever the tensor's own shape says, so a short buffer would be a silent out-of-bounds write."""
    import torch
    dtype = dtype or torch.float64
    if len(tensors) != len(nnz):
        raise ValueError(f"{what}: expected {len(nnz)} tensors, got {len(tensors)}")
    for i, (t, n) in enumerate(zip(tensors, nnz)):
        if t is None:
            continue
        if t.dtype != dtype:
            raise TypeError(f"{what}[{i}]: expected {dtype}, got {t.dtype}")
        if t.device != dev:
            raise ValueError(f"{what}[{i}]: expected a tensor on {dev}, got {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{what}[{i}]: must be contiguous")
        if t.numel() < B * n:
            raise ValueError(f"{what}[{i}]: needs {B} x {n} elements, has {t.numel()}")


def qp_form(inputs: Sequence, out: Optional[List] = None) -> List:
    """qp_former: 17 (B, nnz) fp64 CUDA tensors -> [H, f, A, b, G, d] CCS value tensors."""
    import torch
    lib = load()
    ins, B = _check_inputs(inputs, QP_NNZ_IN)
    dev = ins[0].device
    if out is None:
        out = [torch.empty((B, n), dtype=torch.float64, device=dev) for n in QP_NNZ_OUT]
    else:
        _check_buffers("qp_form out", out, QP_NNZ_OUT, B, dev)
    with torch.cuda.device(dev):
        check(lib.srbd_qp_form(_ptr_array(ins), _ptr_array(out), B, _stream()))
    return out


def marshal(f32_tensors: Sequence, mass: float, mu: float, fix_layout=False, q_stride_bug=False, out=None,
            with_linearisation_point=True, i_body=None, rotation_body=None) -> List:
    """mpc_controller_cusadi.py:54-95 in one launch: 12 fp32 device tensors (x0, x_ref, R_body, I_world,
    root_position, foot_position, contact_table, dt_mpc, Q, R, residual_lin_accel, residual_ang_accel)
    -> the 17 fp64 (B, nnz) qp_former inputs (entries 1, 2 are None unless with_linearisation_point).
    With i_body (3,3) and rotation_body (B,3,3) the kernel computes I_world = R I_body R' itself
    (srbd_marshal_inputs_iw) and f32_tensors[3] may be None."""
    import torch
    lib = load()
    ts = []
    on_device_iw = i_body is not None
    if on_device_iw:
        if rotation_body is None:
            raise ValueError("marshal: i_body needs rotation_body")
        for name, t, n in (("i_body", i_body, 9), ("rotation_body", rotation_body, None)):
            if t.dtype != torch.float32 or t.device.type != "cuda" or not t.is_contiguous() or (n and t.numel() != n):
                raise TypeError(f"marshal: {name} must be a contiguous float32 CUDA tensor" + (f" of {n} elements" if n else ""))
        f32_tensors = list(f32_tensors)
        f32_tensors[3] = rotation_body          # placeholder with the right shape for the checks below; not read
    for i, t in enumerate(f32_tensors):
        if t.dtype != torch.float32 or t.device.type != "cuda":
            raise TypeError(f"marshal: argument {i} must be a float32 CUDA tensor")
        ts.append(t.contiguous())
    if len(ts) != 12:
        raise ValueError("marshal: expected 12 tensors")
    B, dev = ts[0].shape[0], ts[0].device
    # per-env element counts of the 12 inputs (Q and R are shared by all envs): B comes from x0 alone, so check the rest
    per_env = (12, 120, 9, 9, 3, 6, 20, 1, None, None, 3, 3)
    for i, (t, n) in enumerate(zip(ts, per_env)):
        if t.device != dev:
            raise ValueError(f"marshal: argument {i} is on {t.device}, expected {dev}")
        if n is not None and t.numel() < B * n:
            raise ValueError(f"marshal: argument {i} needs {B} x {n} elements, has {t.numel()}")
    if ts[8].numel() < 12 or ts[9].numel() < 12:
        raise ValueError("marshal: Q and R need at least 12 entries")
    if out is None:
        out = [torch.empty((B, n), dtype=torch.float64, device=dev) if (with_linearisation_point or i not in (1, 2)) else None
               for i, n in enumerate(QP_NNZ_IN)]
    else:
        _check_buffers("marshal out", out, QP_NNZ_IN, B, dev)
        if any(out[i] is None for i in range(len(QP_NNZ_IN)) if i not in (1, 2)):
            raise ValueError("marshal out: only entries 1 and 2 (x, u) may be None")
    with torch.cuda.device(dev):
        if on_device_iw:
            if rotation_body.numel() < B * 9 or rotation_body.device != dev or i_body.device != dev:
                raise ValueError("marshal: rotation_body must hold B x 9 elements on the inputs' device")
            check(lib.srbd_marshal_inputs_iw(_ptr_array(ts), ctypes.c_void_p(i_body.data_ptr()), ctypes.c_void_p(rotation_body.data_ptr()),
                                             int(ts[8].numel()), int(bool(fix_layout)), int(bool(q_stride_bug)),
                                             float(mass), float(mu), _ptr_array(out, allow_none=True), B, _stream()))
        else:
            check(lib.srbd_marshal_inputs(_ptr_array(ts), int(ts[8].numel()), int(bool(fix_layout)), int(bool(q_stride_bug)),
                                          float(mass), float(mu), _ptr_array(out, allow_none=True), B, _stream()))
    return out


def reference_trajectory(f32_tensors: Sequence, step: float, world_position_desired, yaw_desired, first_run, x0=None, x_ref=None):
    """base_controller.py:164-257 in one launch: 9 fp32 device tensors (root_euler, root_position,
    root_angular_velocity_w, root_velocity_w, rotation_body, desired_velocity_b, desired_angular_velocity_b,
    desired_height, dt_mpc) -> (x0 (B,12), x_ref (B,10,12)) fp32; the knot-point state
    (world_position_desired, yaw_desired, first_run [torch.bool]) is updated in place."""
    import torch
    lib = load()
    if len(f32_tensors) != 9:
        raise ValueError("reference_trajectory: expected 9 tensors")
    ts = []
    for i, t in enumerate(f32_tensors):
        if t.dtype != torch.float32 or t.device.type != "cuda":
            raise TypeError(f"reference_trajectory: argument {i} must be a float32 CUDA tensor")
        ts.append(t.contiguous())
    B, dev = ts[0].shape[0], ts[0].device
    for name, t, dt, shape in (("world_position_desired", world_position_desired, torch.float32, (B, 3)),
                               ("yaw_desired", yaw_desired, torch.float32, (B,)), ("first_run", first_run, torch.bool, (B,))):
        if t.dtype != dt or t.device != dev or tuple(t.shape) != shape or not t.is_contiguous():
            raise TypeError(f"reference_trajectory: {name} must be a contiguous {dt} tensor of shape {shape} on {dev}")
    if x0 is None:
        x0 = torch.empty((B, 12), dtype=torch.float32, device=dev)
    if x_ref is None:
        x_ref = torch.empty((B, 10, 12), dtype=torch.float32, device=dev)
    for name, t, shape in (("x0", x0, (B, 12)), ("x_ref", x_ref, (B, 10, 12))):
        if t.dtype != torch.float32 or t.device != dev or tuple(t.shape) != shape or not t.is_contiguous():
            raise TypeError(f"reference_trajectory: {name} must be a contiguous float32 tensor of shape {shape} on {dev}")
    with torch.cuda.device(dev):
        check(lib.srbd_reference_trajectory(_ptr_array(ts), float(step), world_position_desired.data_ptr(), yaw_desired.data_ptr(),
                                            first_run.data_ptr(), x0.data_ptr(), x_ref.data_ptr(), B, _stream()))
    return x0, x_ref


def is_stage_invariant(inputs: Sequence) -> bool:
    """Synchronous probe: True when [Q_val, G_val, A_val, ...] is qp_former-shaped for every env
    (identical stage blocks, unit coefficients on x_{k+1} / Mx rows), i.e. the fast kernel applies."""
    import torch
    lib = load()
    chk, B = _check_inputs(list(inputs[:3]), IP_NNZ_IN[:3])
    res = ctypes.c_int(0)
    with torch.cuda.device(chk[0].device):
        check(lib.srbd_pdipm_is_stage_invariant(_ptr_array(chk + [None] * 7, allow_none=True), B, _stream(), ctypes.byref(res)))
    return bool(res.value)


def pdipm(inputs: Sequence, num_iters: int = 5, tol: float = 0.0, cold_start: bool = False,
          stage_invariant: bool = False, out: Optional[List] = None, work=None, iters_done=None,
          precision: str = "f64") -> List:
    """sparse_pdipm_multiple_iterations with a runtime iteration count.
    inputs: [Q_val, G_val, A_val, f, h, b, x, s, z, y] (x,s,z,y may be None with cold_start).
    precision: "f64" (parity path) | "f32" | "f32_refined" | "f32_refined2" (optional reduced-precision modes of the fast
    kernel, include/srbd_mpc.h; need stage_invariant=True).
    Returns [x, s, z, y, residuals(B,4), mu(B,1)]."""
    import torch
    lib = load()
    inputs = list(inputs)
    if cold_start:
        inputs = inputs[:6] + [None] * 4
    chk, B = _check_inputs(inputs[:6], IP_NNZ_IN[:6])
    rest = inputs[6:]
    if not cold_start:
        rest, _ = _check_inputs(rest, IP_NNZ_IN[6:], B)
    ins = chk + list(rest)
    dev = chk[0].device
    if out is None:
        out = [torch.empty((B, n), dtype=torch.float64, device=dev) for n in IP_NNZ_OUT]
    else:
        _check_buffers("pdipm out", out, IP_NNZ_OUT, B, dev)
        if any(o is None for o in out):
            raise ValueError("pdipm out: all six output tensors are required")
    flags = (COLD_START if cold_start else 0) | (STAGE_INVARIANT if stage_invariant else 0) | _PRECISION_FLAGS[PRECISIONS[precision]]
    per_env = WORK_DOUBLES_FAST if stage_invariant else WORK_DOUBLES     # = srbd_pdipm_work_doubles(1, flags)
    if work is None:
        work = torch.empty((B, per_env), dtype=torch.float64, device=dev)
    else:
        _check_buffers("pdipm work", [work], [per_env], B, dev)
    if iters_done is not None:
        _check_buffers("pdipm iters_done", [iters_done], [1], B, dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib.srbd_pdipm(_ptr_array(ins, allow_none=True), _ptr_array(out), ctypes.c_void_p(work.data_ptr()),
                             ctypes.c_void_p(iters_done.data_ptr()) if iters_done is not None else None,
                             B, int(num_iters), float(tol), flags, _stream()))
    return out


def wrench(x_sol, rotation_body):
    import torch
    lib = load()
    B = x_sol.shape[0]
    rot = rotation_body.to(torch.float32).contiguous()
    out = torch.empty((B, 2, 6), dtype=torch.float32, device=x_sol.device)
    with torch.cuda.device(x_sol.device):
        check(lib.srbd_wrench(ctypes.c_void_p(x_sol.data_ptr()), ctypes.c_void_p(rot.data_ptr()),
                              ctypes.c_void_p(out.data_ptr()), B, _stream()))
    return out


def _check_qp_inputs(inputs17):
    """_check_inputs for the 17 qp_former inputs; entries 1, 2 (x, u) may be None = all ones."""
    keep = [i for i in range(len(QP_NNZ_IN)) if i not in (1, 2) or inputs17[i] is not None]
    chk, B = _check_inputs([inputs17[i] for i in keep], [QP_NNZ_IN[i] for i in keep])
    ins = [None] * len(QP_NNZ_IN)
    for i, t in zip(keep, chk):
        ins[i] = t
    return ins, B, chk[0].device


class Context:
    """Library-owned workspace for the fused tick (srbd_ctx_*)."""

    def __init__(self, max_B: int, device: int = -1):
        lib = load()
        self._h = ctypes.c_void_p()
        check(lib.srbd_ctx_create(ctypes.byref(self._h), int(max_B), int(device)))
        self.max_B = max_B

    def set_precision(self, precision: str = "f64"):
        """Precision of the ticks this context launches: "f64" (default, parity path), "f32", "f32_refined",
        "f32_refined2" (include/srbd_mpc.h: optional reduced-precision mode)."""
        check(load().srbd_ctx_set_precision(self._h, PRECISIONS[precision]))

    def close(self):
        if getattr(self, "_h", None):
            load().srbd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, inputs17, rotation_body, num_calls=1, iters_per_call=5, tol=0.0, want_iterates=False,
              iters_done=None):
        """Device tensors in -> (wrench (B,2,6) fp32[, x, s, z, y, res])."""
        import torch
        lib = load()
        ins, B, dev = _check_qp_inputs(inputs17)
        if B > self.max_B:
            raise ValueError(f"Context.solve: batch of {B} envs exceeds the context's max_B = {self.max_B}")
        rot = rotation_body.to(torch.float32).contiguous()
        if rot.numel() < B * 9 or rot.device != dev:
            raise ValueError("Context.solve: rotation_body must hold B x 9 elements on the inputs' device")
        if iters_done is not None:
            _check_buffers("Context.solve iters_done", [iters_done], [1], B, dev, dtype=torch.int32)
        w = torch.empty((B, 2, 6), dtype=torch.float32, device=dev)
        extra = [None] * 5
        if want_iterates:
            extra = [torch.empty((B, n), dtype=torch.float64, device=dev) for n in (240, 160, 160, 140, 4)]
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        with torch.cuda.device(dev):
            check(lib.srbd_mpc_solve(self._h, _ptr_array(ins, allow_none=True), ptr(rot), B, int(num_calls), int(iters_per_call),
                                     float(tol), ptr(w), *[ptr(t) for t in extra], ptr(iters_done), _stream()))
        return (w, *extra) if want_iterates else w

    def solve_warm(self, inputs17, rotation_body, iters=20, tol=0.0, reset=None, want_iterates=False, mode="xy"):
        """EXTENSION (the reference cold-starts every tick): one tick warm-started from what this context found at
        its previous solve_warm call.  mode "xy": (x, y) kept, s = max(h - G x, 1) and z = 1 re-initialised (20
        iterations per tick).  mode "shift": (x, y, z) kept, shifted by one stage when the contact schedule rolled on,
        s, z re-centred where the schedule changed (5-10 iterations per tick); see include/srbd_mpc.h.
        The first call is a cold start.  reset: optional int32 (B,) device tensor, nonzero = cold-start
        that env (episode reset).  inputs17[1] / [2] may be None.  Returns like solve()."""
        mode_i = {"xy": 1, "shift": 2}[mode]
        import torch
        lib = load()
        ins, B, dev = _check_qp_inputs(inputs17)
        rot = rotation_body.to(torch.float32).contiguous()
        w = torch.empty((B, 2, 6), dtype=torch.float32, device=dev)
        if getattr(self, "_warm", None) is None:
            self._warm = [torch.zeros((self.max_B, n), dtype=torch.float64, device=dev) for n in (240, 160, 160, 140, 4)]
            reset = torch.ones(B, dtype=torch.int32, device=dev)        # nothing to start from yet
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        x, s_, z, y, res = self._warm
        with torch.cuda.device(dev):
            check(lib.srbd_mpc_solve_warm(self._h, _ptr_array(ins, allow_none=True), ptr(rot), B, int(iters), float(tol), mode_i,
                                          ptr(reset), ptr(w), ptr(x), ptr(s_), ptr(z), ptr(y), ptr(res), None, _stream()))
        return (w, x[:B], s_[:B], z[:B], y[:B], res[:B]) if want_iterates else w

    def solve_host(self, inputs17_host, rotation_body_host, wrench_host, num_calls=1, iters_per_call=5, tol=0.0,
                   x_out_host=None, res_out_host=None):
        """HOST (ideally pinned) tensors in/out; H2D + kernels + D2H; synchronous.
        inputs17_host[1] / [2] (x, u: the linearisation point) may be None = all ones, the constant
        the reference call site uses (mpc_controller_cusadi.py:55-56); they are then not copied.
        A control loop that refills the same buffers every tick should bind them once: bind_host()."""
        return self.bind_host(inputs17_host, rotation_body_host, wrench_host, x_out_host, res_out_host)(num_calls, iters_per_call, tol)

    def bind_host(self, inputs17_host, rotation_body_host, wrench_host, x_out_host=None, res_out_host=None):
        """Validates the host buffers ONCE and returns a callable `run(num_calls=1, iters_per_call=5, tol=0.0)` that executes
        srbd_mpc_solve_host on them and returns wrench_host.  The checks and the pointer marshalling of 19 tensors cost
        ~45 us of Python per call -- 6 % of a 4 096-env tick -- and a control loop overwrites the same (pinned) buffers in
        place every tick.  The callable keeps the tensors alive; it must not be used after the context is closed."""
        import torch
        lib = load()
        B = inputs17_host[0].shape[0]
        if B > self.max_B:
            raise ValueError(f"Context.solve_host: batch of {B} envs exceeds the context's max_B = {self.max_B}")
        for i, (t, n) in enumerate(zip(inputs17_host, QP_NNZ_IN)):
            if t is None:
                if i not in (1, 2):
                    raise ValueError(f"solve_host: input {i} is required")
                continue
            if t.dtype != torch.float64 or t.device.type != "cpu" or not t.is_contiguous() or t.numel() != B * n:
                raise ValueError(f"solve_host: input {i} must be a contiguous float64 host tensor of {B} x {n} elements")
        for name, t, n, dt in (("rotation_body_host", rotation_body_host, 9, torch.float32), ("wrench_host", wrench_host, 12, torch.float32),
                               ("x_out_host", x_out_host, 240, torch.float64), ("res_out_host", res_out_host, 4, torch.float64)):
            if t is None and name in ("x_out_host", "res_out_host"):
                continue
            if t.dtype != dt or t.device.type != "cpu" or not t.is_contiguous() or t.numel() < B * n:
                raise ValueError(f"solve_host: {name} must be a contiguous {dt} host tensor with room for {B} x {n} elements")
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        keep = (list(inputs17_host), rotation_body_host, wrench_host, x_out_host, res_out_host)
        arr, p_rot, p_w, p_x, p_res = _ptr_array(inputs17_host, allow_none=True), ptr(rotation_body_host), ptr(wrench_host), ptr(x_out_host), ptr(res_out_host)
        fn, ctx = lib.srbd_mpc_solve_host, self

        def run(num_calls=1, iters_per_call=5, tol=0.0, _keep=keep):
            rc = fn(ctx._h, arr, p_rot, B, int(num_calls), int(iters_per_call), float(tol), p_w, p_x, p_res)
            if rc:
                check(rc)
            return wrench_host
        return run


def t1_kinematics(joint_angle, want_jacobian=True):
    """Batched T1 forward kinematics (+ contact Jacobian): joint_angle (B,2,6) fp32 CUDA -> (foot_position (B,2,3),
    jacobian (B,2,6,6) or None).  Replaces core/robot/t1.py:96-138 (CasADi on the CPU, env 0 only)."""
    import torch
    lib = load()
    if joint_angle.dtype != torch.float32 or joint_angle.device.type != "cuda" or joint_angle.shape[-2:] != (2, 6):
        raise TypeError("t1_kinematics: joint_angle must be a float32 CUDA tensor of shape (B, 2, 6)")
    q = joint_angle.contiguous()
    B, dev = q.shape[0], q.device
    p = torch.empty((B, 2, 3), dtype=torch.float32, device=dev)
    J = torch.empty((B, 2, 6, 6), dtype=torch.float32, device=dev) if want_jacobian else None
    with torch.cuda.device(dev):
        check(lib.srbd_t1_kinematics(ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(p.data_ptr()),
                                     ctypes.c_void_p(J.data_ptr()) if J is not None else None, B, _stream()))
    return p, J


def t1_inverse_kinematics(foot_position):
    """T1.analytical_IK (t1.py:139-213) for both legs of every env: (B,2,3) fp32 CUDA -> (B,2,6)."""
    import torch
    lib = load()
    if foot_position.dtype != torch.float32 or foot_position.device.type != "cuda" or foot_position.shape[-2:] != (2, 3):
        raise TypeError("t1_inverse_kinematics: foot_position must be a float32 CUDA tensor of shape (B, 2, 3)")
    p = foot_position.contiguous()
    B, dev = p.shape[0], p.device
    q = torch.empty((B, 2, 6), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.srbd_t1_inverse_kinematics(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(q.data_ptr()), B, _stream()))
    return q


def dfma_peak(iters=20000, blocks=148 * 8, threads=256) -> float:
    lib = load()
    out = ctypes.c_double()
    check(lib.srbd_dfma_peak(iters, blocks, threads, ctypes.byref(out)))
    return out.value
