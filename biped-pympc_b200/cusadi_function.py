"""Host-side mirror of the reference's operator object ``CusadiFunction``
(biped_pympc/cusadi/src/CusadiFunction.py:6-117) for the two hot-path functions.

Same public surface -- ``evaluate(inputs)``, ``outputs_sparse``, ``inputs_sparse``,
``getDenseOutput(i)``, ``eval_time``, ``fn_name``, ``num_instances`` -- but:
  * metadata (nnz, sparsity) comes from the library's frozen patterns instead of a loaded
    ``casadi.Function`` (CasADi is not needed at run time);
  * no ``work`` tensor of ``B * sz_w`` doubles and no zero-fill of outputs before each call
    (CusadiFunction.py:104-109): the kernels write every output element;
  * the call is asynchronous on the current torch stream (the reference's ``evaluate``
    blocks on ``cudaEventSynchronize``, generateCUDACode.py:166-177);
  * ``getDenseOutput`` scatters through precomputed index tensors instead of building a
    ``torch.sparse_coo_tensor`` with ``torch.tensor(range(B))`` on every call.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import _native


class FunctionSpec:
    """The slice of the casadi.Function interface CusadiFunction relies on."""

    def __init__(self, name, names_in, nnz_in, names_out, shapes_out, patterns_out):
        self._name, self._names_in, self._nnz_in = name, tuple(names_in), tuple(nnz_in)
        self._names_out, self._shapes_out, self._patterns_out = tuple(names_out), tuple(shapes_out), tuple(patterns_out)

    def name(self): return self._name
    def n_in(self): return len(self._nnz_in)
    def n_out(self): return len(self._shapes_out)
    def nnz_in(self, i): return self._nnz_in[i]
    def name_in(self, i): return self._names_in[i]
    def name_out(self, i): return self._names_out[i]
    def size1_out(self, i): return self._shapes_out[i][0]
    def size2_out(self, i): return self._shapes_out[i][1]
    def sz_w(self): return 0

    def triplet_out(self, i):
        """(rows, cols) of output i's nonzeros in nnz (column-major) order."""
        pat = self._patterns_out[i]
        n1, n2 = self._shapes_out[i]
        if pat is None:                      # dense column vector / dense matrix
            k = np.arange(n1 * n2)
            return (k % n1).astype(np.int64), (k // n1).astype(np.int64)
        colptr, row, _, ncol = _native.pattern(pat)
        col = np.repeat(np.arange(ncol), np.diff(colptr))
        return row.astype(np.int64), col.astype(np.int64)

    def nnz_out(self, i):
        pat = self._patterns_out[i]
        if pat is None:
            return self._shapes_out[i][0] * self._shapes_out[i][1]
        return {"H": 240, "A": 1196, "G": 280}[pat]


# srbd_constraints.py:75-79
QP_FORMER_FUNCTION = FunctionSpec(
    "qp_former",
    ("x0", "x", "u", "x_ref", "dt", "m", "mu", "R_body", "I_world", "body_pos", "left_foot_pos",
     "right_foot_pos", "contact_table", "Q", "R", "residual_lin_accel", "residual_ang_accel"),
    _native.QP_NNZ_IN, ("H", "f", "A", "b", "G", "d"),
    ((240, 240), (240, 1), (140, 240), (140, 1), (160, 240), (160, 1)),
    ("H", None, "A", None, "G", None))
# sparse_pdipm_solver.py:532-534
PDIPM_FUNCTION = FunctionSpec(
    "sparse_pdipm_multiple_iterations",
    ("Q_val", "G_val", "A_val", "f", "h", "b", "x", "s", "z", "y"), _native.IP_NNZ_IN,
    ("x_new", "s_new", "z_new", "y_new", "residuals", "mu_new"),
    ((240, 1), (160, 1), (160, 1), (140, 1), (4, 1), (1, 1)), (None,) * 6)


class B200Function:
    """``CusadiFunction``-compatible operator backed by the hand-written sm_100a kernels."""

    def __init__(self, fn, num_instances: int, device="cuda", num_iters: int = 5, tol: float = 0.0,
                 stage_invariant=None):
        import torch
        # `fn` may be the function's name or any object with a .name() -- in particular the casadi.Function the reference
        # passes to CusadiFunction (mpc_controller_cusadi.py:28-34).  Either way the metadata comes from the library's own
        # frozen patterns (FunctionSpec): a casadi.Function has no triplet_out(), and the kernels are tied to these
        # patterns anyway.  The caller's object is kept for reference only.
        name = fn if isinstance(fn, str) else fn.name()
        specs = {"qp_former": QP_FORMER_FUNCTION, "sparse_pdipm_multiple_iterations": PDIPM_FUNCTION}
        if name not in specs:
            raise ValueError(f"no B200 kernel for function {name!r}")
        self.fn_casadi_given = None if isinstance(fn, str) else fn
        fn = specs[name]
        _native.load()                      # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _native.SrbdError("B200Function needs a CUDA device (there is no CPU fallback)")
        self.fn_casadi = fn
        self.fn_name = fn.name()
        self.num_instances = int(num_instances)
        # stage_invariant: True / False force the fast / general kernel; None (default) probes every
        # batch (the reference's evaluate is synchronous too, CusadiFunction.py:43) and picks the fast
        # kernel for qp_former-shaped data
        self.num_iters, self.tol, self.stage_invariant = int(num_iters), float(tol), stage_invariant
        self._device = torch.device(device)
        self.eval_time = 0.0
        B = self.num_instances
        self._input_tensors: List = [None] * fn.n_in()
        self._output_tensors = [torch.empty((B, fn.nnz_out(i)), dtype=torch.float64, device=self._device)
                                for i in range(fn.n_out())]
        # scratch for either kernel (the general one needs the larger buffer); stage_invariant=True callers only need
        # WORK_DOUBLES_FAST per env
        per_env = _native.WORK_DOUBLES_FAST if stage_invariant is True else _native.WORK_DOUBLES
        self._work = (torch.empty((B, per_env), dtype=torch.float64, device=self._device)
                      if self.fn_name != "qp_former" else None)
        self.iters_done = (torch.zeros(B, dtype=torch.int32, device=self._device)
                           if self.fn_name != "qp_former" else None)
        self.inputs_sparse = self._input_tensors
        self.outputs_sparse = self._output_tensors
        self._scatter_idx = {}

    def evaluate(self, inputs: Sequence):
        if len(inputs) != self.fn_casadi.n_in():
            raise ValueError(f"{self.fn_name}: expected {self.fn_casadi.n_in()} inputs, got {len(inputs)}")
        # the reference's operator has a fixed batch: its buffers are allocated for num_instances envs and the generated
        # kernel is launched with num_instances threads (CusadiFunction.py:43-46); a different batch is an error here
        for i, t in enumerate(inputs):
            if t is not None and t.numel() != self.num_instances * self.fn_casadi.nnz_in(i):
                raise ValueError(f"{self.fn_name}: input {i} has {t.numel()} elements, expected "
                                 f"{self.num_instances} x {self.fn_casadi.nnz_in(i)} (num_instances is fixed at construction)")
        self._input_tensors = list(inputs)
        self.inputs_sparse = self._input_tensors
        if self.fn_name == "qp_former":
            _native.qp_form(inputs, out=self._output_tensors)
        else:
            inv = self.stage_invariant
            if inv is None:
                inv = _native.is_stage_invariant(inputs)
            _native.pdipm(inputs, num_iters=self.num_iters, tol=self.tol, stage_invariant=bool(inv),
                          out=self._output_tensors, work=self._work, iters_done=self.iters_done)

    def getDenseOutput(self, out_idx: int):
        import torch
        fn = self.fn_casadi
        n1, n2 = fn.size1_out(out_idx), fn.size2_out(out_idx)
        if out_idx not in self._scatter_idx:
            rows, cols = fn.triplet_out(out_idx)
            self._scatter_idx[out_idx] = torch.as_tensor(rows * n2 + cols, device=self._device)
        flat = torch.zeros((self.num_instances, n1 * n2), dtype=torch.float64, device=self._device)
        flat[:, self._scatter_idx[out_idx]] = self._output_tensors[out_idx]
        return flat.view(self.num_instances, n1, n2)
