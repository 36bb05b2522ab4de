"""Device-side counterpart of the reference's control stack for the HECTOR biped:
``MPCController`` (biped_pympc/mpc_wrapper.py:3-204) over ``BipedController``
(biped_pympc/biped_controller.py:13-178), i.e. gait generator, state estimator, leg controller, swing-leg
controller and the MPC back-end, with the same step sequence and the same public surface

    set_command -> update_state -> [run_mpc every `decimation` steps] -> run_lowlevel -> get_action

One control step is two kernel launches (``srbd_update_state``, ``srbd_lowlevel``) instead of ~450 small
torch kernels; an MPC step adds the gait table, the state builder, the input builder and the fused solve
(``srbd_gait``, ``srbd_reference_trajectory``, ``srbd_marshal_inputs``, ``srbd_mpc_solve``).  All buffers
are allocated once; nothing on the step path synchronises the host.  The tensors keep the reference's
attribute names under ``.state_estimator.data``, ``.leg_controller.data``, ``.leg_controller.command``,
``.swing_leg_controller`` and ``.gait_generator`` so that RL-side code reading them keeps working.
There is no CPU path: every method raises without the CUDA library and a CUDA device.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace

from . import _native
from .mpc_controller_b200 import MPCControllerB200

_FP = ctypes.c_void_p


class _LoopBuffers(ctypes.Structure):          # = srbd_loop_buffers (include/srbd_mpc.h), same field order
    _fields_ = [(n, _FP) for n in (
        "gait_phase", "dt_mpc", "ssp_durations", "dsp_durations", "swing_time_remaining", "first_swing", "p0",
        "foot_height", "cp1", "cp2", "desired_velocity_b",
        "root_position", "root_euler", "rotation_body", "root_velocity_b", "root_angular_velocity_b", "root_velocity_w",
        "root_angular_velocity_w", "foot_position",
        "q", "qd", "p", "v", "J", "contact_phase", "swing_phase", "contact_bool", "swing_bool",
        "feedforward_force",
        "foot_placement", "foot_placement_b", "pDes", "vDes", "tau", "qDes", "qdDes", "kpjoint", "kdjoint", "action")]


class HectorModel:
    """Constants of core/robot/hector.py:7-41 (the kinematics themselves live in csrc/srbd_lowlevel.cu)."""
    mass = 13.856
    mu = 1.0
    num_dof = 5
    kp = (40.0, 40.0, 70.0, 70.0, 40.0)
    kd = (1.0, 1.0, 0.7, 0.7, 0.7)
    torque_limit = (33.5, 33.5, 33.5, 67.0, 33.5, 33.5, 33.5, 33.5, 67.0, 33.5)

    def __init__(self, device):
        import torch
        self.I_body = torch.tensor([[0.5413, 0.0, 0.0], [0.0, 0.5200, 0.0], [0.0, 0.0, 0.0691]], device=device)
        self.pd_conf = SimpleNamespace(kp=list(self.kp), kd=list(self.kd), torque_limit=list(self.torque_limit))


class BipedControllerB200:
    def __init__(self, num_envs: int, device="cuda", dt: float = 0.001, dt_mpc: float = 0.025, horizon_length: int = 10,
                 decimation: int = 10, Q=None, R=None, ssp_durations: int = 5, dsp_durations: int = 0, swing_height: float = 0.1,
                 swing_reference_frame: str = "base", gait_id: int = 2, mpc_calls: int = 4, mpc_iters_per_call: int = 5):
        import torch
        self.num_envs, self.num_legs, self.device = num_envs, 2, torch.device(device)
        if self.device.type != "cuda":
            raise _native.SrbdError("BipedControllerB200 needs a CUDA device (no CPU fallback)")
        self._lib = _native.load()
        B, dev = num_envs, self.device
        f = lambda *s, fill=0.0: torch.full(s, fill, dtype=torch.float32, device=dev)  # noqa: E731
        self.dt, self.horizon_length, self.mpc_decimation = float(dt), horizon_length, int(decimation)
        self.dt_mpc = f(B, fill=dt_mpc)
        self.world_frame = {"base": 0, "world": 1}[swing_reference_frame]
        if gait_id == 1:       # standing gait, biped_controller.py:26-28
            dsp, ssp = (5, 5), (0, 0)
        elif gait_id == 2:     # walking gait, :29-32
            dsp, ssp = (dsp_durations, dsp_durations), (ssp_durations, ssp_durations)
        else:
            raise ValueError(f"Invalid gait_id: {gait_id}. Must be 1 (standing) or 2 (walking).")
        i2 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev).unsqueeze(0).repeat(B, 1).contiguous()  # noqa: E731
        self.biped = HectorModel(dev)
        # ---- the reference's objects, as namespaces over persistent tensors -------------------------
        self.gait_generator = SimpleNamespace(
            gait_phase=f(B), ssp_durations=i2(ssp), dsp_durations=i2(dsp), dt_mpc=self.dt_mpc,
            mpc_table=torch.ones(B, horizon_length, 2, dtype=torch.int32, device=dev), contact_table=f(B, horizon_length, 2, fill=1.0))
        self.desired_state = SimpleNamespace(desired_velocity_b=f(B, 3), desired_angular_velocity_b=f(B, 3),
                                             desired_height=f(B, fill=0.55))
        self.state_estimator = SimpleNamespace(data=SimpleNamespace(
            root_position=f(B, 3), root_quat=f(B, 4), root_euler=f(B, 3), rotation_body=torch.eye(3, device=dev).repeat(B, 1, 1),
            root_velocity_b=f(B, 3), root_angular_velocity_b=f(B, 3), root_velocity_w=f(B, 3), root_angular_velocity_w=f(B, 3),
            foot_position=f(B, 2, 3)))
        self.leg_controller = SimpleNamespace(
            biped=self.biped,
            data=SimpleNamespace(q=f(B, 2, 5), qd=f(B, 2, 5), p=f(B, 2, 3), v=f(B, 2, 3), J=f(B, 2, 6, 5),
                                 contact_phase=f(B, 2), swing_phase=f(B, 2), contact_bool=f(B, 2, fill=1.0), swing_bool=f(B, 2)),
            command=SimpleNamespace(tau=f(B, 2, 5), qDes=f(B, 2, 5), qdDes=f(B, 2, 5), pDes=f(B, 2, 3), vDes=f(B, 2, 3),
                                    feedfowardforce=f(B, 2, 6), kpjoint=f(B, 2, 5), kdjoint=f(B, 2, 5)))
        swing_dur = (self.gait_generator.ssp_durations * self.dt_mpc[:, None]).float()
        self.swing_leg_controller = SimpleNamespace(
            swing_time_remaining=swing_dur.clone(), first_swing=torch.ones(B, 2, dtype=torch.bool, device=dev), p0=f(B, 2, 3),
            foot_height=f(B, fill=swing_height), cp1_coef=f(B, fill=1 / 3), cp2_coef=f(B, fill=2 / 3),
            foot_placement=f(B, 2, 3), foot_placement_b=f(B, 2, 3))
        self.action = f(B, 10)
        self.torque_limit = torch.tensor(self.biped.torque_limit, device=dev)
        self.mpc_cost = f(B)
        self._k = {n: (ctypes.c_float * len(v))(*v) for n, v in (("kp", self.biped.kp), ("kd", self.biped.kd),
                                                                   ("lim", self.biped.torque_limit))}
        # ---- MPC back-end ---------------------------------------------------------------------------
        m = self.mpc_controller = MPCControllerB200()
        m.num_envs, m.device, m.horizon_length = B, dev, horizon_length
        m.cfg = SimpleNamespace(dt=self.dt, decimation=self.mpc_decimation)
        m.biped, m.mass, m.mu = self.biped, self.biped.mass, self.biped.mu
        m.Q = torch.as_tensor([150, 150, 250, 100, 100, 250, 1, 1, 5, 10, 10, 1, 1] if Q is None else Q, dtype=torch.float32, device=dev)
        m.R = torch.as_tensor([1e-5] * 6 + [1e-4] * 6 if R is None else R, dtype=torch.float32, device=dev)
        m.residual_lin_accel, m.residual_ang_accel = f(B, 3), f(B, 3)
        m.dt_mpc = self.dt_mpc
        m.state_estimate_data, m.desired_state_data = self.state_estimator.data, self.desired_state
        m.contact_table = self.gait_generator.contact_table
        m.MAX_ITER, m.ITERS_PER_CALL = mpc_calls, mpc_iters_per_call
        m.init_solver()
        self._bind()

    # ---- plumbing ---------------------------------------------------------------------------------
    def _bind(self):
        """(re)build the srbd_loop_buffers struct; call again after replacing any of the tensors"""
        g, d, s = self.gait_generator, self.desired_state, self.state_estimator.data
        ld, c, sw = self.leg_controller.data, self.leg_controller.command, self.swing_leg_controller
        t = dict(gait_phase=g.gait_phase, dt_mpc=self.dt_mpc, ssp_durations=g.ssp_durations, dsp_durations=g.dsp_durations,
                 swing_time_remaining=sw.swing_time_remaining, first_swing=sw.first_swing, p0=sw.p0, foot_height=sw.foot_height,
                 cp1=sw.cp1_coef, cp2=sw.cp2_coef, desired_velocity_b=d.desired_velocity_b,
                 root_position=s.root_position, root_euler=s.root_euler, rotation_body=s.rotation_body,
                 root_velocity_b=s.root_velocity_b, root_angular_velocity_b=s.root_angular_velocity_b,
                 root_velocity_w=s.root_velocity_w, root_angular_velocity_w=s.root_angular_velocity_w, foot_position=s.foot_position,
                 q=ld.q, qd=ld.qd, p=ld.p, v=ld.v, J=ld.J, contact_phase=ld.contact_phase, swing_phase=ld.swing_phase,
                 contact_bool=ld.contact_bool, swing_bool=ld.swing_bool, feedforward_force=c.feedfowardforce,
                 foot_placement=sw.foot_placement, foot_placement_b=sw.foot_placement_b, pDes=c.pDes, vDes=c.vDes, tau=c.tau,
                 qDes=c.qDes, qdDes=c.qdDes, kpjoint=c.kpjoint, kdjoint=c.kdjoint, action=self.action)
        self._buf = _LoopBuffers()
        self._keep = t
        for n, v in t.items():
            assert v.is_contiguous() and v.device == self.device, n
            setattr(self._buf, n, v.data_ptr())

    def _stream(self):
        return _native._stream()

    # ---- MPCController (mpc_wrapper.py) surface -------------------------------------------------------
    def set_command(self, twist, height):
        d = self.desired_state
        d.desired_velocity_b[:, :2] = twist[:, :2]
        d.desired_angular_velocity_b[:, 2] = twist[:, 2]
        d.desired_height[:] = height

    set_desired_state = set_command          # BipedController's name for it (biped_controller.py:91-94)

    def update_state(self, state):
        """state (B, >= 33) fp32: [position 3, quaternion wxyz 4, lin vel b 3, ang vel b 3, q 10, qd 10, (tau 10)]."""
        import torch
        if state.dtype != torch.float32 or state.device != self.device or state.dim() != 2 or state.shape[0] != self.num_envs \
                or state.shape[1] < 33 or state.stride(1) != 1:
            raise TypeError("update_state: expected a float32 (B, >=33) tensor on " + str(self.device))
        self.state_estimator.data.root_quat = state[:, 3:7]
        with torch.cuda.device(self.device):
            _native.check(self._lib.srbd_update_state(ctypes.byref(self._buf), state.data_ptr(), int(state.stride(0)),
                                                      self.num_envs, self._stream()))

    def run_mpc(self):
        """biped_controller.py:110-118, 129-142: gait table -> MPC -> feed-forward wrench."""
        import torch
        g = self.gait_generator
        with torch.cuda.device(self.device):
            _native.check(self._lib.srbd_gait(g.gait_phase.data_ptr(), self.dt_mpc.data_ptr(), g.ssp_durations.data_ptr(),
                                              g.dsp_durations.data_ptr(), self.dt, 0, self.horizon_length, None, None,
                                              g.mpc_table.data_ptr(), g.contact_table.data_ptr(), self.num_envs, self._stream()))
        wrench, cost = self.mpc_controller.run()
        self.leg_controller.command.feedfowardforce.copy_(wrench.reshape(-1, 2, 6))
        self.mpc_cost = cost

    def run_lowlevel(self):
        import torch
        with torch.cuda.device(self.device):
            _native.check(self._lib.srbd_lowlevel(ctypes.byref(self._buf), self._k["kp"], self._k["kd"], self._k["lim"], self.dt,
                                                  self.world_frame, 1, self.num_envs, self._stream()))

    # ---- CUDA graphs: the launch-bound control step as one graph replay -----------------------------------
    def capture_graphs(self, state_width: int = 43):
        """Captures the two step shapes -- update_state + run_lowlevel, and update_state + run_mpc + run_lowlevel --
        into CUDA graphs over a static (B, state_width) input buffer; `step(state, mpc)` then costs one copy and one
        graph replay instead of 2 (resp. ~12) launches issued from Python.  The persistent controller state is
        saved around the warm-up runs capture needs, so capturing does not advance the gait or the knot point."""
        import torch
        self._static_state = torch.zeros(self.num_envs, state_width, dtype=torch.float32, device=self.device)
        m, sw, g = self.mpc_controller, self.swing_leg_controller, self.gait_generator

        knot = ("world_position_desired", "yaw_desired", "first_run")
        had_knot = all(isinstance(getattr(m, n, None), torch.Tensor) for n in knot)
        persistent = [g.gait_phase, sw.swing_time_remaining, sw.first_swing, sw.p0] + ([getattr(m, n) for n in knot] if had_knot else [])
        saved = [t.clone() for t in persistent]
        self._static_state[:, 3] = 1.0                    # a valid quaternion and leg posture for the warm-up runs
        self._static_state[:, 2] = 0.55
        self._static_state[:, 13:23] = torch.tensor([0.0, 0.0, 0.6, -1.2, 0.6] * 2, device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.update_state(self._static_state); self.run_mpc(); self.run_lowlevel()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph_low, self._graph_mpc = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_low):
            self.update_state(self._static_state); self.run_lowlevel()
        with torch.cuda.graph(self._graph_mpc):
            self.update_state(self._static_state); self.run_mpc(); self.run_lowlevel()
        for t, v in zip(persistent, saved):
            t.copy_(v)
        if not had_knot:
            m.first_run.fill_(True)                       # the knot point re-anchors at the first real MPC tick
        return self

    def step(self, state, mpc: bool):
        """One control step through the captured graphs: update_state -> [run_mpc] -> run_lowlevel; the joint
        torque is in `self.action` afterwards.  set_command / update_mpc_sampling_time / set_swing_parameters write
        into the buffers the graphs read, so they keep working between steps."""
        self._static_state[:, :state.shape[1]].copy_(state)
        (self._graph_mpc if mpc else self._graph_low).replay()
        return self.action

    def get_action(self):
        return self.action.clone()

    @property
    def command_joint_torque(self):
        return self.action

    def reset(self, env_ids):
        """BipedController.reset (biped_controller.py:83-88) as it actually behaves: the gait phase returns to 0, the MPC
        knot point re-anchors (first_run) and the swing controller starts a fresh swing (first_swing).  The reference
        also calls `.zero()` on the estimator / leg data / leg command tensors, but through `tensor[env_id, ...].zero_()`
        (core/data/robot_data.py:27-38, 108-143), which zeroes a copy made by the advanced index: those buffers -- the
        feed-forward wrench in particular -- keep their values until the next update, here as there."""
        import torch
        self.gait_generator.gait_phase[env_ids] = 0.0
        self.mpc_controller.reset(env_ids)
        self.state_estimator.data.rotation_body[env_ids] = torch.eye(3, device=self.device)
        self.swing_leg_controller.first_swing[env_ids] = True

    # ---- DRL interface (mpc_wrapper.py:44-67) ----------------------------------------------------------
    def update_mpc_sampling_time(self, dt_mpc):
        self.dt_mpc.copy_(dt_mpc)

    def set_swing_parameters(self, foot_height, cp1, cp2):
        sw = self.swing_leg_controller
        sw.foot_height.copy_(foot_height); sw.cp1_coef.copy_(cp1); sw.cp2_coef.copy_(cp2)

    def set_srbd_accel(self, residual_lin_accel, residual_ang_accel):
        self.mpc_controller.residual_lin_accel.copy_(residual_lin_accel)
        self.mpc_controller.residual_ang_accel.copy_(residual_ang_accel)

    # ---- properties (mpc_wrapper.py:72-150) --------------------------------------------------------------
    @property
    def ground_reaction_wrench(self):
        return self.leg_controller.command.feedfowardforce.clone()

    @property
    def contact_state(self):
        return (self.leg_controller.data.contact_phase != -1).float()

    @property
    def contact_phase(self):
        cp = self.leg_controller.data.contact_phase
        return (cp != -1).float() * cp

    @property
    def swing_state(self):
        return (self.leg_controller.data.swing_phase != -1).float()

    @property
    def swing_phase(self):
        sp = self.leg_controller.data.swing_phase
        return (sp != -1).float() * sp

    @property
    def foot_placement(self):
        return self.swing_leg_controller.foot_placement.clone()

    @property
    def foot_placement_b(self):
        return self.swing_leg_controller.foot_placement_b.clone()

    @property
    def ref_foot_pos_b(self):
        return self.leg_controller.command.pDes.clone()

    @property
    def ref_foot_vel_b(self):
        return self.leg_controller.command.vDes.clone()

    @property
    def foot_pos_b(self):
        return self.leg_controller.data.p.clone()

    @property
    def foot_vel_b(self):
        return self.leg_controller.data.v.clone()

    def close(self):
        self.mpc_controller._ctx.close()
