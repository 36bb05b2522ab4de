"""BASELINE.json configs[4] / SURVEY 8(d) config 5: closed-loop rollout of the WHOLE control stack
(BipedControllerB200: gait generator + state estimator + HECTOR leg controller + swing-leg controller + MPC)
for a batch of envs on one GPU.  TEST / MEASUREMENT HARNESS, not product code.

The plant is a stand-in (there is no simulator in this image): the single rigid body the MPC models
(srbd_centroidal_model.py:123-166), integrated at 1 kHz with quaternion kinematics and driven by the ground
reaction wrench the controller commands for the feet in contact, plus KINEMATIC legs -- a stance foot stays
where it touched down (its joint angles follow from the body pose by the analytical IK of hector.py:234-291),
a swing leg tracks the controller's joint targets.  It is written with batched torch ops on the device; only
the controller calls are timed (CUDA events around update_state / run_mpc / run_lowlevel of every step).

What it shows: per-control-step and per-MPC-tick latency of the stack at 4096 envs, and that the loop it closes
is physically sane (bodies stay up and follow the command).
"""
from __future__ import annotations

import math

import numpy as np
import torch

G = 9.81


def quat_mul(a, b):
    aw, ax, ay, az = a.unbind(1)
    bw, bx, by, bz = b.unbind(1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], 1)


def quat_rot(q):
    w, x, y, z = (q / q.norm(dim=1, keepdim=True)).unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)


def hector_ik(p, leg):
    """analytical IK of hector.py:234-291 (restated for the plant's stance legs)"""
    side = 1.0 if leg else -1.0
    f = p - torch.tensor([0.0465 - 0.06, -side * (0.047 + 0.015), -0.126 - 0.0705], device=p.device)
    fz = f[:, 2] + 0.042
    d2 = torch.sqrt(f[:, 1] ** 2 + fz ** 2)
    q1 = torch.asin((f[:, 1] / d2).clamp(-1, 1)) + torch.asin(((0.018 + 0.01805) * side / d2).clamp(-1, 1))
    s, c = torch.sin(q1), torch.cos(q1)
    gx, gy, gz = f[:, 0], c * f[:, 1] - s * fz + 0.018 * side, s * f[:, 1] + c * fz
    r2 = gx * gx + gy * gy + gz * gz
    c2 = ((r2 - 2 * 0.22 ** 2) / (2 * 0.22 * 0.22)).clamp(-1, 1)
    s2 = (-torch.sqrt((1 - c2 * c2).clamp(min=1e-6))).clamp(-1, 1)
    q3 = torch.atan2(s2, c2)
    q2 = torch.atan2(-gx, -gz) - torch.atan2(0.22 * s2, 0.22 + 0.22 * c2)
    return torch.stack([torch.zeros_like(q1), q1, q2, q3, -q2 - q3], 1)


class StackRollout:
    def __init__(self, ctrl, seed=0, v_max=0.3, graphs=False):
        self.c, dev, B = ctrl, ctrl.device, ctrl.num_envs
        self.graphs = graphs
        g = torch.Generator(device="cpu").manual_seed(seed)
        r = lambda *s: torch.rand(*s, generator=g).to(dev)  # noqa: E731
        yaw = (2 * r(B) - 1) * math.pi
        self.quat = torch.stack([torch.cos(yaw / 2), torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw / 2)], 1)
        self.pos = torch.stack([4 * r(B) - 2, 4 * r(B) - 2, torch.full((B,), 0.55, device=dev)], 1)
        self.v = torch.zeros(B, 3, device=dev)
        self.w = torch.zeros(B, 3, device=dev)
        self.twist = torch.stack([(2 * r(B) - 1) * v_max, (2 * r(B) - 1) * 0.1, (2 * r(B) - 1) * 0.3], 1)
        self.height = torch.full((B,), 0.55, device=dev)
        ctrl.gait_generator.gait_phase[:] = r(B)
        Rm = quat_rot(self.quat)
        hip = torch.tensor([[0.0465 - 0.06, 0.098, -0.55], [0.0465 - 0.06, -0.098, -0.55]], device=dev)
        self.foot = self.pos[:, None, :] + torch.einsum("bij,fj->bfi", Rm, hip)
        self.foot[:, :, 2] = 0.0
        self.q = torch.zeros(B, 10, device=dev)
        self.qd = torch.zeros(B, 10, device=dev)
        self.prev_contact = torch.ones(B, 2, device=dev)
        self.I_body = torch.diag(torch.tensor([0.5413, 0.5200, 0.0691], device=dev))
        self.mass = ctrl.biped.mass
        self.state = torch.zeros(B, 43, device=dev)
        self.lat = {"step": [], "mpc": []}

    def _stance_legs(self, Rm):
        pb = torch.einsum("bji,bfj->bfi", Rm, self.foot - self.pos[:, None, :])      # R' (foot - pos)
        return torch.cat([hector_ik(pb[:, 0], 0), hector_ik(pb[:, 1], 1)], 1)

    def step(self, k, timed=True):
        c, dt = self.c, self.c.dt
        Rm = quat_rot(self.quat)
        contact = (c.leg_controller.data.contact_phase != -1).float() if k else torch.ones_like(self.prev_contact)
        # kinematic legs: stance from the body pose, swing from the controller's targets
        q_st = self._stance_legs(Rm)
        q_sw = c.leg_controller.command.qDes.reshape(-1, 10)
        m = contact.repeat_interleave(5, dim=1)
        q_new = m * q_st + (1 - m) * q_sw
        self.qd = (q_new - self.q) / dt if k else torch.zeros_like(self.q)
        self.q = q_new
        st = self.state
        st[:, 0:3] = self.pos; st[:, 3:7] = self.quat
        st[:, 7:10] = torch.einsum("bji,bj->bi", Rm, self.v); st[:, 10:13] = torch.einsum("bji,bj->bi", Rm, self.w)
        st[:, 13:23] = self.q; st[:, 23:33] = self.qd
        # ---- the controller: the timed region -------------------------------------------------------
        ev = [None, None, None]
        if timed:
            ev[0], ev[2] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mpc = (k % c.mpc_decimation == 0)
        c.set_command(self.twist, self.height)
        if timed:
            ev[0].record()
        if self.graphs:
            c.step(st, mpc)
        else:
            c.update_state(st)
            if mpc:
                c.run_mpc()
            c.run_lowlevel()
        if timed:
            ev[2].record()
            self.lat["step"].append((ev[0], ev[2], mpc))
        # ---- plant ------------------------------------------------------------------------------------
        contact = (c.leg_controller.data.contact_phase != -1).float()
        touchdown = (contact == 1) & (self.prev_contact == 0)
        fk_w = c.state_estimator.data.foot_position.clone()
        fk_w[:, :, 2] = 0.0
        self.foot = torch.where(touchdown[:, :, None], fk_w, self.foot)
        swing = contact == 0
        self.foot = torch.where(swing[:, :, None], fk_w, self.foot)         # a swinging foot is where its leg puts it
        self.prev_contact = contact
        wr = c.leg_controller.command.feedfowardforce                        # body frame, reaction on the robot = -wrench
        f_w = -torch.einsum("bij,bfj->bfi", Rm, wr[:, :, :3]) * contact[:, :, None]
        m_w = -torch.einsum("bij,bfj->bfi", Rm, wr[:, :, 3:]) * contact[:, :, None]
        tau = torch.cross(self.foot - self.pos[:, None, :], f_w, dim=2).sum(1) + m_w.sum(1)
        I_w = Rm @ self.I_body @ Rm.transpose(1, 2)
        self.w = self.w + dt * torch.linalg.solve(I_w, tau.unsqueeze(-1)).squeeze(-1)
        self.v = self.v + dt * (f_w.sum(1) / self.mass + torch.tensor([0.0, 0.0, -G], device=self.pos.device))
        self.pos = self.pos + dt * self.v
        dq = torch.cat([torch.zeros_like(self.w[:, :1]), 0.5 * dt * self.w], 1)
        self.quat = self.quat + quat_mul(dq, self.quat)
        self.quat = self.quat / self.quat.norm(dim=1, keepdim=True)

    def run(self, steps, warmup=0):
        for k in range(steps):
            self.step(k, timed=k >= warmup)
        torch.cuda.synchronize()
        lat = np.array([[a.elapsed_time(b), float(m)] for a, b, m in self.lat["step"]])
        return lat

    def summary(self):
        Rm = quat_rot(self.quat)
        vb = torch.einsum("bji,bj->bi", Rm, self.v)
        up = Rm[:, 2, 2]
        ok = torch.isfinite(self.pos).all(1) & (up > 0.9) & ((self.pos[:, 2] - 0.55).abs() < 0.12)
        return {"upright_fraction": float(ok.float().mean()), "height_err_max_of_upright": float((self.pos[ok, 2] - 0.55).abs().max()) if ok.any() else float("nan"),
                "vx_err_mean_of_upright": float((vb[ok, 0] - self.twist[ok, 0]).abs().mean()) if ok.any() else float("nan"),
                "tilt_cos_min_of_upright": float(up[ok].min()) if ok.any() else float("nan")}


def main(B=4096, steps=1000, decimation=10, mpc_calls=4, iters_per_call=5, warmup=20, dev="cuda:0", seed=1, graphs=0):
    import biped_pympc_b200 as bp
    # a 12-long Q: the first 12 entries of MPCConf.Q (configuration.py:43-47) for every env, so that the reference's stride-12
    # read of its 13-long Q plays no role in the rollout
    c = bp.BipedControllerB200(B, dev, decimation=decimation, mpc_calls=mpc_calls, mpc_iters_per_call=iters_per_call,
                               Q=[150, 150, 250, 100, 100, 250, 1, 1, 5, 10, 10, 1])
    if graphs:
        c.capture_graphs()
    ro = StackRollout(c, seed=seed, graphs=bool(graphs))
    lat = ro.run(steps, warmup)
    mpc, low = lat[lat[:, 1] == 1, 0], lat[lat[:, 1] == 0, 0]
    pct = lambda a, p: float(np.percentile(a, p)) if len(a) else None  # noqa: E731
    n_mpc = int((lat[:, 1] == 1).sum())
    res = {"envs": B, "cuda_graphs": bool(graphs), "control_steps": steps, "decimation": decimation, "mpc_iterations_per_tick": mpc_calls * iters_per_call,
           "mpc_ticks_timed": n_mpc, "controller_ms_total": float(lat[:, 0].sum()),
           "step_ms_without_mpc": {"p50": pct(low, 50), "p99": pct(low, 99)}, "step_ms_with_mpc": {"p50": pct(mpc, 50), "p99": pct(mpc, 99)},
           "env_control_steps_per_s": float(B * len(lat) / (lat[:, 0].sum() * 1e-3)),
           "mpc_env_ticks_per_s_in_loop": float(B * n_mpc / (mpc.sum() * 1e-3)) if n_mpc else None, **ro.summary()}
    c.close()
    return res


if __name__ == "__main__":
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    kw = dict(a.split("=") for a in sys.argv[1:])
    print(json.dumps(main(**{k: int(v) for k, v in kw.items()})))
