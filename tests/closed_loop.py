"""Closed-loop rollout of the MPC solve on a single-rigid-body plant (BASELINE.json configs[4] in
miniature): gait schedule -> 17 qp_former inputs -> form + PDIPM -> u0 -> plant, for a batch of
envs.  TEST / MEASUREMENT HARNESS, not product code: the reference's gait generator, swing-leg and
leg controllers are outside the hot path; here they are replaced by the simplest
stand-ins that close the loop:

  * plant: the same single rigid body the MPC models (srbd_centroidal_model.py:123-166) integrated at
    1 kHz with exact ZYX Euler kinematics;
  * gait: the contact table semantics of GaitGenerator.mpc_gait (gait_generator.py:215-252), 5 + 5
    knots of single support, phase advanced in real time;
  * feet: a swinging foot moves linearly to a Raibert touchdown point under its hip.

What it is for: next to the fixtures recorded from the reference's own code (tests/golden/), a physically sane
closed loop -- the body stays up, tracks the commanded velocity, respects the friction cone -- is
independent evidence that the restated formulation (signs, frames, layouts Q1/Q2 of SURVEY 8b) is
right.  `solve(inputs17) -> x (B, 240)` is supplied by the caller: the CUDA path in the GPU test,
the C oracle in the CPU test.
"""
from __future__ import annotations

import math

import numpy as np

G = 9.81
HZ = 0.55
DT_SIM = 1e-3
DECIMATION = 10          # MPC tick every 10 plant steps (configuration.py:41)
DT_MPC = 0.025
HORIZON = 10
SSP = 5                  # knots of single support per foot (configuration.py:17-18)
HIP = 0.1                # lateral hip offset
MASS, MU = 40.0, 1.0
I_BODY = np.diag([0.5413, 0.5200, 0.0691])
Q = np.array([150, 150, 250, 100, 100, 250, 1, 1, 5, 10, 10, 1], float)
R = np.array([1e-5] * 6 + [1e-4] * 6)


def rot_zyx(e):
    r, p, y = e[:, 0], e[:, 1], e[:, 2]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rm = np.empty((e.shape[0], 3, 3))
    Rm[:, 0, 0] = cy * cp; Rm[:, 0, 1] = cy * sp * sr - sy * cr; Rm[:, 0, 2] = cy * sp * cr + sy * sr
    Rm[:, 1, 0] = sy * cp; Rm[:, 1, 1] = sy * sp * sr + cy * cr; Rm[:, 1, 2] = sy * sp * cr - cy * sr
    Rm[:, 2, 0] = -sp; Rm[:, 2, 1] = cp * sr; Rm[:, 2, 2] = cp * cr
    return Rm


def contact_table(phase):
    """(B, 10, 2) contact flags over the horizon, left stance first (mpc_gait semantics)."""
    cycle = 2 * SSP
    step0 = (phase * cycle).astype(np.int64)
    steps = (step0[:, None] + np.arange(HORIZON)[None, :]) % cycle
    tab = np.zeros((phase.shape[0], HORIZON, 2))
    tab[:, :, 0] = steps < SSP          # left stance / right swing
    tab[:, :, 1] = steps >= SSP         # right stance / left swing
    return tab


class Rollout:
    def __init__(self, B, seed=0, walking=True):
        rng = np.random.default_rng(seed)
        self.B, self.walking = B, walking
        self.euler = np.zeros((B, 3)); self.euler[:, 2] = rng.uniform(-math.pi, math.pi, B)
        self.euler[:, :2] = rng.normal(0, 0.02, (B, 2))
        self.p = np.zeros((B, 3)); self.p[:, :2] = rng.normal(0, 1, (B, 2)); self.p[:, 2] = HZ + rng.normal(0, 0.01, B)
        # standing has no stepping to recover with (line feet: 7 cm + 4 cm of support): start it at rest
        sig = 0.05 if walking else 0.0
        self.w = rng.normal(0, 1, (B, 3)) * sig; self.v = rng.normal(0, 1, (B, 3)) * sig
        self.v_cmd_b = np.zeros((B, 3))
        if walking:
            self.v_cmd_b[:, 0] = rng.uniform(-0.3, 0.4, B); self.v_cmd_b[:, 1] = rng.uniform(-0.1, 0.1, B)
        self.yawrate_cmd = rng.uniform(-0.3, 0.3, B) if walking else np.zeros(B)
        self.phase = rng.uniform(0, 1, B) if walking else np.zeros(B)
        self.yaw_des = self.euler[:, 2].copy()
        self.p_des = self.p[:, :2].copy()        # anchored position reference (world_position_desired, base_controller.py:223-226)
        Rz = rot_zyx(np.stack([np.zeros(B), np.zeros(B), self.euler[:, 2]], 1))
        self.foot = np.zeros((B, 2, 3))
        for i, sgn in enumerate((+1.0, -1.0)):
            self.foot[:, i, :2] = self.p[:, :2] + (Rz @ np.array([0.0, sgn * HIP, 0.0]))[:, :2]
        self.swing_from = self.foot.copy()
        self.prev_contact = np.ones((B, 2))
        self.t = 0.0
        self.max_cone_violation = 0.0

    # ---- the 17 qp_former inputs (kernel nnz order: column-major contact table) ----
    def inputs(self):
        B = self.B
        Rm = rot_zyx(self.euler)
        v_cmd_w = (Rm @ self.v_cmd_b[:, :, None])[:, :, 0]
        tb = DT_MPC * np.arange(HORIZON)[None, :]
        x_ref = np.zeros((B, HORIZON, 12))
        x_ref[:, :, 2] = self.yaw_des[:, None] + self.yawrate_cmd[:, None] * tb
        x_ref[:, :, 3] = self.p_des[:, 0:1] + v_cmd_w[:, 0:1] * tb
        x_ref[:, :, 4] = self.p_des[:, 1:2] + v_cmd_w[:, 1:2] * tb
        x_ref[:, :, 5] = HZ
        x_ref[:, :, 8] = self.yawrate_cmd[:, None]
        x_ref[:, :, 9] = v_cmd_w[:, 0:1]; x_ref[:, :, 10] = v_cmd_w[:, 1:2]
        tab = contact_table(self.phase) if self.walking else np.ones((B, HORIZON, 2))
        I_w = Rm @ I_BODY[None] @ Rm.transpose(0, 2, 1)
        x0 = np.concatenate([self.euler, self.p, self.w, self.v], 1)
        ones = np.ones((B, 120))
        rep = lambda a: np.repeat(np.asarray(a, float)[None], B, 0)
        return [x0, ones, ones.copy(), x_ref.reshape(B, 120), np.full((B, 1), DT_MPC), np.full((B, 1), MASS), np.full((B, 1), MU),
                Rm.reshape(B, 9),                         # row-major flatten, as mpc_controller_cusadi.py:58 (Q1)
                I_w.transpose(0, 2, 1).reshape(B, 9), self.p.copy(), self.foot[:, 0].copy(), self.foot[:, 1].copy(),
                tab.transpose(0, 2, 1).reshape(B, 20),    # column-major: flat[i] = left_i, flat[10 + i] = right_i
                rep(Q), rep(R), np.zeros((B, 3)), np.zeros((B, 3))], tab[:, 0, :]

    def plant_steps(self, u0, contact):
        B = self.B
        fL, fR, mL, mR = u0[:, 0:3] * contact[:, 0:1], u0[:, 3:6] * contact[:, 1:2], u0[:, 6:9].copy(), u0[:, 9:12].copy()
        mL[:, 0] = 0; mR[:, 0] = 0
        mL *= contact[:, 0:1]; mR *= contact[:, 1:2]
        for f in (fL, fR):   # friction pyramid / unilateral contact actually delivered by the solve
            viol = np.maximum(np.abs(f[:, 0]) - MU * f[:, 2], np.abs(f[:, 1]) - MU * f[:, 2])
            self.max_cone_violation = max(self.max_cone_violation, float(np.max(np.maximum(viol, -f[:, 2]))))
        for _ in range(DECIMATION):
            Rm = rot_zyx(self.euler)
            I_w = Rm @ I_BODY[None] @ Rm.transpose(0, 2, 1)
            tau = np.cross(self.foot[:, 0] - self.p, fL) + np.cross(self.foot[:, 1] - self.p, fR) + mL + mR
            wd = np.linalg.solve(I_w, tau[:, :, None])[:, :, 0]
            vd = (fL + fR) / MASS + np.array([0, 0, -G])
            self.w += DT_SIM * wd; self.v += DT_SIM * vd
            self.p += DT_SIM * self.v
            r, pch, y = self.euler[:, 0], self.euler[:, 1], self.euler[:, 2]
            rd = (np.cos(y) * self.w[:, 0] + np.sin(y) * self.w[:, 1]) / np.cos(pch)
            pd = -np.sin(y) * self.w[:, 0] + np.cos(y) * self.w[:, 1]
            yd = self.w[:, 2] + np.sin(pch) * rd
            self.euler += DT_SIM * np.stack([rd, pd, yd], 1)
        self.t += DT_SIM * DECIMATION
        Rm = rot_zyx(self.euler)
        self.p_des += DT_SIM * DECIMATION * (Rm @ self.v_cmd_b[:, :, None])[:, :2, 0]

    def advance_gait_and_feet(self, contact):
        if not self.walking:
            return
        B = self.B
        tick = DT_SIM * DECIMATION
        self.phase = (self.phase + tick / (2 * SSP * DT_MPC)) % 1.0
        self.yaw_des += tick * self.yawrate_cmd
        nxt = contact_table(self.phase)[:, 0, :]
        Rz = rot_zyx(np.stack([np.zeros(B), np.zeros(B), self.euler[:, 2]], 1))
        t_st = SSP * DT_MPC
        for i, sgn in enumerate((+1.0, -1.0)):
            hip = self.p + (Rz @ np.array([0.0, sgn * HIP, 0.0]))
            target = hip[:, :2] + 0.5 * t_st * self.v[:, :2] + 0.05 * (self.v[:, :2] - (Rz @ self.v_cmd_b[:, :, None])[:, :2, 0])
            lift = (self.prev_contact[:, i] == 1) & (nxt[:, i] == 0)
            self.swing_from[lift, i] = self.foot[lift, i]
            swing = nxt[:, i] == 0
            # fraction of the swing phase elapsed: phase within this foot's swing half of the cycle
            ph = self.phase if i == 1 else (self.phase - 0.5) % 1.0       # right swings first (phase < 0.5)
            frac = np.clip(ph / 0.5, 0, 1)[:, None]
            self.foot[swing, i, :2] = ((1 - frac) * self.swing_from[:, i, :2] + frac * target)[swing]
        self.prev_contact = nxt

    def run(self, solve, ticks):
        for _ in range(ticks):
            ins, contact = self.inputs()
            x = solve(ins)
            u0 = np.asarray(x)[:, 120:132]
            self.plant_steps(u0, contact)
            self.advance_gait_and_feet(contact)
        return self

    def summary(self):
        Rm = rot_zyx(self.euler)
        v_b = (Rm.transpose(0, 2, 1) @ self.v[:, :, None])[:, :, 0]
        return {
            "finite": bool(np.isfinite(self.p).all() and np.isfinite(self.euler).all()),
            "height_err_max": float(np.abs(self.p[:, 2] - HZ).max()),
            "tilt_max": float(np.abs(self.euler[:, :2]).max()),
            "vx_err_mean": float(np.abs(v_b[:, 0] - self.v_cmd_b[:, 0]).mean()),
            "vx_err_max": float(np.abs(v_b[:, 0] - self.v_cmd_b[:, 0]).max()),
            "cone_violation_max": self.max_cone_violation,
        }


if __name__ == "__main__":   # CPU smoke run on the C oracle
    import os, sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    from oracle import c_oracle as co
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    walking = (sys.argv[3] != "stand") if len(sys.argv) > 3 else True
    ro = Rollout(B, seed=3, walking=walking)
    ro.run(lambda ins: co.pdipm(co.qp_former(ins), 20)[0], ticks)
    print(ro.summary())
