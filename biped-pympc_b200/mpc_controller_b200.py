"""Host glue of the GPU MPC back-end: the B200 counterpart of
``MPCControllerCusadi`` (biped_pympc/convex_mpc/mpc_controller_cusadi.py:15-205).

``MPCControllerB200.run()`` takes and returns exactly what the reference's ``run()`` does
(``foot_wrench (B,2,6) fp32, cost (B,)``) and consumes the same controller attributes
(``contact_table, dt_mpc, mass, mu, Q, R, residual_*_accel, cfg.{decimation, dt}``,
``state_estimate_data`` and ``desired_state_data`` as ``BaseMPCController`` holds them,
base_controller.py:61-257; ``x0 / x_ref / R_body`` are built on the device by ``build_state`` or, with
``fused_state_builder = False``, by the host class's own methods), so ``BipedController`` only needs a new
key in its solver switch (biped_controller.py:36-45).  It is written as a mixin-style class
with no dependency on the reference package (which cannot be imported without CasADi).

Marshalling follows mpc_controller_cusadi.py:54-95 including its three layout quirks
(SURVEY.md 8b): ``R_body`` and the ``(B,10,2)`` contact table are flattened row-major into
buffers the kernel reads in CasADi column-major order (Q1, Q2), and the 13-long ``MPCConf.Q``
is strided by 12 so that env e reads ``Qflat[12e : 12e+12]`` (Q3).  ONE switch governs all
three: ``reference_quirks = True`` (the default) reproduces the reference's ``run()`` bit for
bit, ``reference_quirks = False`` passes the mathematically intended layouts and ``Q[:12]``
for every env.  ``fix_layout`` / ``reproduce_q_stride_bug`` override the individual quirks
(``None`` = follow ``reference_quirks``); an incoherent combination with a ``Q`` that is not
12 long warns once.
"""
from __future__ import annotations

from typing import Tuple

from . import _native


def marshal_inputs(ctrl, fix_layout: bool = False, reproduce_q_stride_bug: bool = False, with_linearisation_point=True,
                   torch_i_world: bool = False):
    """The 17 fp64 (B, nnz) tensors of mpc_controller_cusadi.py:77-95 from controller state, built by ONE
    kernel (srbd_marshal_inputs_iw) instead of ~25 small torch ops; `marshal_inputs_torch` is the literal
    torch restatement the tests compare it with.  I_world = R I_body R' (:60-63) is computed inside the same launch
    (fp32, within 2 ulp of torch's two batched matmuls); torch_i_world=True computes it with torch like the reference and
    passes it in, which reproduces the reference's own rounding on the same back-end."""
    import torch
    sed = ctrl.state_estimate_data
    rot = sed.rotation_body
    f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=ctrl.device)
    iw_args = {}
    if torch_i_world:
        I_w = rot @ ctrl.biped.I_body.unsqueeze(0) @ rot.transpose(1, 2)
    else:
        I_w = None
        iw_args = dict(i_body=f32(ctrl.biped.I_body).contiguous(), rotation_body=f32(rot).contiguous())
    return _native.marshal(
        [f32(ctrl.x0), f32(ctrl.x_ref), f32(ctrl.R_body[:, 0, :, :]), I_w, f32(sed.root_position), f32(sed.foot_position),
         f32(ctrl.contact_table), f32(ctrl.dt_mpc), f32(ctrl.Q), f32(ctrl.R), f32(ctrl.residual_lin_accel),
         f32(ctrl.residual_ang_accel)],
        float(ctrl.mass), float(ctrl.mu), fix_layout, reproduce_q_stride_bug and ctrl.Q.numel() != 12,
        with_linearisation_point=with_linearisation_point, **iw_args)


def marshal_inputs_torch(ctrl, fix_layout: bool = False, reproduce_q_stride_bug: bool = False):
    """Literal torch restatement of mpc_controller_cusadi.py:54-95 (parity reference of marshal_inputs)."""
    import torch
    B, dev, Hn = ctrl.num_envs, ctrl.device, ctrl.horizon_length
    sed = ctrl.state_estimate_data
    f64 = torch.float64
    x0 = ctrl.x0.to(f64).contiguous()
    ones = torch.ones(B, 12 * Hn, device=dev, dtype=f64)
    x_ref = ctrl.x_ref.contiguous().view(B, -1).to(f64)
    Rb = ctrl.R_body[:, 0, :, :]
    rot = sed.rotation_body
    I_w = rot @ ctrl.biped.I_body.unsqueeze(0) @ rot.transpose(1, 2)
    contact = ctrl.contact_table
    if fix_layout:
        Rb = Rb.transpose(1, 2)           # column-major nnz order of the true R_body
        I_w = I_w.transpose(1, 2)
        contact = contact.transpose(1, 2)  # (B,2,10): flat[k] = contact[k % 10, k // 10]
    Rb = Rb.contiguous().view(B, -1).to(f64)
    I_w = I_w.contiguous().view(B, -1).to(f64)
    contact = contact.contiguous().view(B, -1).to(f64)
    Q = ctrl.Q.to(f64)
    if reproduce_q_stride_bug and Q.numel() != 12:
        Qb = Q.unsqueeze(0).repeat(B, 1).reshape(-1)[:B * 12].view(B, 12)   # env e reads Qflat[12e:12e+12]
    else:
        Qb = Q[:12].unsqueeze(0).repeat(B, 1)
    Rw = ctrl.R.to(f64).unsqueeze(0).repeat(B, 1)
    scalar = lambda v: torch.as_tensor(v, device=dev, dtype=f64) * torch.ones(B, 1, device=dev, dtype=f64)
    return [
        x0, ones, ones.clone(), x_ref, ctrl.dt_mpc.to(f64).reshape(B, 1), scalar(ctrl.mass), scalar(ctrl.mu),
        Rb, I_w, sed.root_position.to(f64).contiguous(), sed.foot_position[:, 0, :].to(f64).contiguous(),
        sed.foot_position[:, 1, :].to(f64).contiguous(), contact, Qb.contiguous(), Rw.contiguous(),
        ctrl.residual_lin_accel.to(f64).contiguous(), ctrl.residual_ang_accel.to(f64).contiguous(),
    ]


def solve_tick(inputs17, rotation_body, num_calls: int = 4, iters_per_call: int = 5, tol: float = 0.0,
               ctx: "_native.Context" = None, want_iterates: bool = False):
    """form -> cold start -> num_calls x iters_per_call Newton iterations -> body wrench,
    entirely on the device (mpc_controller_cusadi.py:97-203)."""
    own = ctx is None
    if own:
        ctx = _native.Context(inputs17[0].shape[0], inputs17[0].device.index or 0)
    try:
        return ctx.solve(inputs17, rotation_body, num_calls, iters_per_call, tol, want_iterates)
    finally:
        if own:
            import torch
            torch.cuda.current_stream().synchronize()
            ctx.close()


class MPCControllerB200:
    """Mixin with the ``init_solver`` / ``run`` pair of the reference's MPC back-ends.

    Use as ``class MPCControllerB200(MPCControllerB200Mixin, BaseMPCController)`` inside the
    reference package or stand-alone by providing the attributes listed in
    the module docstring."""

    MAX_ITER = 4          # calls per tick, mpc_controller_cusadi.py:144
    ITERS_PER_CALL = 5    # generate_solver_function.py:106
    reference_quirks = True          # Q1-Q3 as the reference has them: run() == MPCControllerCusadi.run()
    fix_layout = None                # override of Q1 / Q2 (None = not reference_quirks)
    reproduce_q_stride_bug = None    # override of Q3      (None = reference_quirks)
    tol = 0.0
    _warned_quirks = False

    def _quirk_flags(self):
        fix = (not self.reference_quirks) if self.fix_layout is None else bool(self.fix_layout)
        qbug = self.reference_quirks if self.reproduce_q_stride_bug is None else bool(self.reproduce_q_stride_bug)
        if not fix and not qbug and self.Q.numel() != 12 and not MPCControllerB200._warned_quirks:
            import warnings
            warnings.warn("MPCControllerB200: the reference's R_body / contact-table flattening is reproduced but its "
                          f"stride-12 read of the {self.Q.numel()}-long Q is not: the wrench differs from "
                          "MPCControllerCusadi.run() for every env after env 0 (set reference_quirks or both overrides)")
            MPCControllerB200._warned_quirks = True
        return fix, qbug

    def init_solver(self):
        _native.load()
        dev_index = getattr(self.device, "index", None)
        if dev_index is None:
            import torch
            dev_index = torch.device(self.device).index or 0
        self._ctx = _native.Context(self.num_envs, dev_index)
        return self._ctx

    # True: knot points, initial state and reference trajectory come from ONE kernel (build_state) instead of
    # the host class's ~40 torch ops (base_controller.py:164-257); False: the inherited methods run as they are.
    fused_state_builder = True

    def build_state(self):
        """compute_knot_points + compute_horizon_state + set_initial_state + compute_reference_trajectory
        (base_controller.py:164-257) as one launch of srbd_reference_trajectory.  Needs
        state_estimate_data, desired_state_data, dt_mpc and cfg.{decimation, dt}; creates the persistent
        buffers (x0, x_ref, world_position_desired, yaw_desired, first_run) when the host class has not."""
        import torch
        sed, des = self.state_estimate_data, self.desired_state_data
        B, dev = self.num_envs, sed.rotation_body.device
        def mk(name, shape, dtype=torch.float32, fill=0):
            t = getattr(self, name, None)
            if not isinstance(t, torch.Tensor):
                t = torch.full(shape, fill, dtype=dtype, device=dev)
                setattr(self, name, t)
            return t

        wpd = mk("world_position_desired", (B, 3))
        yaw = mk("yaw_desired", (B,))
        first = mk("first_run", (B,), torch.bool, True)
        x0 = mk("x0", (B, 12))
        x_ref = mk("x_ref", (B, self.horizon_length, 12))
        f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev)  # noqa: E731
        _native.reference_trajectory(
            [f32(sed.root_euler), f32(sed.root_position), f32(sed.root_angular_velocity_w), f32(sed.root_velocity_w),
             f32(sed.rotation_body), f32(des.desired_velocity_b), f32(des.desired_angular_velocity_b),
             f32(des.desired_height), f32(self.dt_mpc)],
            float(self.cfg.decimation * self.cfg.dt), wpd, yaw, first, x0, x_ref)
        # compute_horizon_state (:178-196): the current rotation for the whole horizon; the inverse world
        # inertia and the foot skews it also fills are not read by this back-end
        self.R_body = sed.rotation_body.unsqueeze(1).expand(-1, self.horizon_length, -1, -1)
        return x0, x_ref

    def reset(self, env_ids):
        """BaseMPCController.reset (base_controller.py:259-266): the knot point of these envs re-anchors at the
        current pose on the next tick."""
        first = getattr(self, "first_run", None)
        if first is not None:
            first[env_ids] = True

    def run(self) -> Tuple["torch.Tensor", "torch.Tensor"]:
        import torch
        if self.fused_state_builder and getattr(self, "desired_state_data", None) is not None:
            self.build_state()
        else:
            for hook in ("compute_knot_points", "compute_horizon_state", "set_initial_state", "compute_reference_trajectory"):
                fn = getattr(self, hook, None)
                if fn is not None:
                    fn()
        fix, qbug = self._quirk_flags()
        inputs = marshal_inputs(self, fix, qbug, with_linearisation_point=False)
        foot_wrench = self._ctx.solve(inputs, self.state_estimate_data.rotation_body,
                                      self.MAX_ITER, self.ITERS_PER_CALL, self.tol)
        cost = torch.zeros(self.num_envs, device=self.device)   # reference hard-wires 0 (:185)
        return foot_wrench, cost
