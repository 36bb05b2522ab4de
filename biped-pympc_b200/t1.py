"""Host-side mirror of the reference's T1 robot model (biped_pympc/core/robot/t1.py:50-213) with BATCHED kinematics.

The reference's ``T1.forward_kinematics`` / ``contact_jacobian`` evaluate serialized CasADi functions on the CPU for env
0 only (t1.py:96-138: "TODO: cpu specific, later change to batch process").  ``T1B200`` keeps the method names and
argument meaning -- ``forward_kinematics(joint_angle, leg)``, ``get_p0e(leg)``, ``contact_jacobian(leg)``,
``analytical_IK(p, leg)``, ``hip_horizontal_location(leg)``, ``mass / I_body / mu / num_dof / pd_conf`` -- and returns the
result for EVERY env, computed by one launch of ``srbd_t1_kinematics`` (both legs at once; the second leg's call is
served from the cache when the caller passed the (B,2,6) joint tensor via ``update``)."""
from __future__ import annotations

from types import SimpleNamespace

from . import _native


class T1B200:
    mass = 40.0
    mu = 1.0
    num_dof = 6

    def __init__(self, num_envs: int, device="cuda"):
        import torch
        _native.load()
        self.num_envs, self.device = num_envs, torch.device(device)
        if self.device.type != "cuda":
            raise _native.SrbdError("T1B200 needs a CUDA device (no CPU fallback)")
        self.I_body = torch.tensor([[0.5413, 0.0, 0.0], [0.0, 0.5200, 0.0], [0.0, 0.0, 0.0691]], device=self.device)
        self.hip_yaw = torch.tensor([-0.00, 0.047, -0.126], device=self.device)
        self.hip_roll = torch.tensor([0.0465, 0.015, -0.0705], device=self.device)
        self.pd_conf = SimpleNamespace(kp=[20.0, 20.0, 20.0, 20.0, 15.0, 15.0], kd=[1.0, 1.0, 0.7, 0.7, 0.7, 0.7],
                                       torque_limit=[33.5, 33.5, 33.5, 67.0, 33.5, 33.5, 33.5, 33.5, 33.5, 67.0, 33.5, 33.5])
        self._q = torch.zeros(num_envs, 2, 6, device=self.device)
        self._p = [None, None]
        self._J = [None, None]

    def hip_horizontal_location(self, leg: int):
        import torch
        side = 1 * (1 - leg) - 1 * leg
        cog = torch.zeros(self.num_envs, 3, device=self.device)
        cog[:, 0] = 0.0625 - 0.014
        cog[:, 1] = side * 0.106
        return cog

    def update(self, joint_angle_both):
        """(B,2,6): one launch for both legs; forward_kinematics / get_p0e / contact_jacobian then read the cache."""
        self._q = joint_angle_both.to(self.device).float().contiguous()
        p, J = _native.t1_kinematics(self._q)
        self._p, self._J = [p[:, 0], p[:, 1]], [J[:, 0], J[:, 1]]

    def forward_kinematics(self, joint_angle, leg: int) -> None:
        assert joint_angle.shape[-1] == self.num_dof
        self._q[:, leg] = joint_angle
        p, J = _native.t1_kinematics(self._q)
        self._p[leg], self._J[leg] = p[:, leg], J[:, leg]

    def get_p0e(self, leg: int):
        if self._p[leg] is None:
            raise RuntimeError("Call forward_kinematics first")
        return self._p[leg]

    def contact_jacobian(self, leg: int):
        if self._J[leg] is None:
            raise RuntimeError("Call forward_kinematics first")
        return self._J[leg]

    def analytical_IK(self, p_foot_des_b, leg: int):
        import torch
        both = torch.zeros(p_foot_des_b.shape[0], 2, 3, device=self.device)
        both[:, leg] = p_foot_des_b
        return _native.t1_inverse_kinematics(both)[:, leg]
