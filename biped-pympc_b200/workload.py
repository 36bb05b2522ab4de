"""Synthetic randomized biped states + gait schedules (host side, numpy).

This is the benchmark/test workload of SURVEY.md 8(d) "Config 2": every quantity is
produced the way the reference's host code would produce it
(``biped_pympc/convex_mpc/base_controller.py:178-257`` for x0 / x_ref,
``biped_pympc/core/gait/gait_generator.py:215-252`` for the contact table,
``biped_pympc/convex_mpc/mpc_controller_cusadi.py:54-95`` for the marshalling into the
17 flat fp64 ``(B, nnz)`` tensors of ``qp_former``), including the reference's
flattening quirks (row-major flattening of ``R_body`` and of the
``(10, 2)`` contact table into buffers the kernel reads in CasADi column-major order).
"""
from __future__ import annotations

import numpy as np

HORIZON = 10
# configuration.py:43-52 (first 12 of the 13-long Q); fp32 like the reference
Q_DEFAULT = np.array([150, 150, 250, 100, 100, 250, 1, 1, 5, 10, 10, 1], np.float32).astype(np.float64)
R_DEFAULT = np.array([1e-5] * 6 + [1e-4] * 6, np.float32).astype(np.float64)
I_BODY = np.diag([0.5413, 0.5200, 0.0691])          # hector.py:33-38 / t1.py:70-75
MASS_T1 = 40.0
MU_T1 = 1.0

INPUT_NAMES = ("x0", "x", "u", "x_ref", "dt", "m", "mu", "R_body", "I_world", "body_pos",
               "left_foot_pos", "right_foot_pos", "contact_table", "Q", "R",
               "residual_lin_accel", "residual_ang_accel")
INPUT_NNZ = (12, 120, 120, 120, 1, 1, 1, 9, 9, 3, 3, 3, 20, 12, 12, 3, 3)


def _rot_zyx(roll, pitch, yaw):
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    R = np.empty(roll.shape + (3, 3))
    R[..., 0, 0] = cy * cp; R[..., 0, 1] = cy * sp * sr - sy * cr; R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp; R[..., 1, 1] = sy * sp * sr + cy * cr; R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp;     R[..., 2, 1] = cp * sr;                R[..., 2, 2] = cp * cr
    return R


def mpc_gait_table(gait_phase, ssp=(5, 5), dsp=(0, 0), horizon=HORIZON):
    """Contact table (B, horizon, 2) with the semantics of GaitGenerator.mpc_gait
    (gait_generator.py:215-252): int steps, ssp/dsp = (left, right)."""
    B = gait_phase.shape[0]
    cycle = ssp[0] + ssp[1] + dsp[0] + dsp[1]
    step0 = (gait_phase * cycle).astype(np.int64)
    steps = (step0[:, None] + np.arange(horizon)[None, :]) % cycle
    p1 = steps < ssp[1]
    p2 = (steps >= ssp[1]) & (steps < ssp[1] + dsp[0])
    p3 = (steps >= ssp[1] + dsp[0]) & (steps < ssp[1] + dsp[0] + ssp[0])
    rest = ~(p1 | p2 | p3)
    table = np.zeros((B, horizon, 2))
    table[:, :, 0][p1 | p2 | rest] = 1
    table[:, :, 1][p2 | p3 | rest] = 1
    return table


def make_batch(B, seed=20261018, rank=0, standing=False, fixed_dt=None, zero_residual=False,
               fp32_inputs=True, mass=MASS_T1, mu=MU_T1):
    """Returns (inputs17: list of (B, nnz) float64 C-contiguous arrays in the reference's
    argument order, rotation_body: (B,3,3) float64 row-major)."""
    rng = np.random.default_rng(seed + rank)
    yaw = rng.uniform(-np.pi, np.pi, B)
    roll = rng.normal(0.0, 0.1, B)
    pitch = rng.normal(0.0, 0.1, B)
    if standing:
        yaw[:] = 0; roll[:] = 0; pitch[:] = 0
    Rb = _rot_zyx(roll, pitch, yaw)
    I_world = Rb @ I_BODY[None] @ np.swapaxes(Rb, 1, 2)
    body_pos = np.stack([rng.normal(0, 1, B), rng.normal(0, 1, B), 0.55 + rng.normal(0, 0.02, B)], 1)
    hip = np.zeros((B, 2, 3)); hip[:, 0, 1] = 0.106; hip[:, 1, 1] = -0.106
    feet = body_pos[:, None, :] + np.einsum("bij,bfj->bfi", Rb, hip)
    feet[:, :, :2] += rng.normal(0, 0.03, (B, 2, 2))
    feet[:, :, 2] = 0.0
    omega = rng.normal(0, 0.2, (B, 3)); vel = rng.normal(0, 0.3, (B, 3))
    if standing:
        omega[:] = 0; vel[:] = 0
    x0 = np.concatenate([np.stack([roll, pitch, yaw], 1), body_pos, omega, vel], 1)
    cmd_v = np.zeros((B, 3)); cmd_v[:, :2] = rng.uniform(-0.5, 0.5, (B, 2))
    cmd_w = rng.uniform(-0.5, 0.5, B)
    if standing:
        cmd_v[:] = 0; cmd_w[:] = 0
    dt = rng.uniform(0.02, 0.05, B) if fixed_dt is None else np.full(B, float(fixed_dt))
    # reference trajectory, base_controller.py:213-257 (first tick: desired = current)
    tb = dt[:, None] * np.arange(HORIZON)[None, :]
    v_w = np.einsum("bij,bj->bi", Rb, cmd_v)
    x_ref = np.zeros((B, HORIZON, 12))
    x_ref[:, :, 2] = yaw[:, None] + cmd_w[:, None] * tb
    x_ref[:, :, 3] = body_pos[:, 0:1] + v_w[:, 0:1] * tb
    x_ref[:, :, 4] = body_pos[:, 1:2] + v_w[:, 1:2] * tb
    x_ref[:, :, 5] = 0.55
    x_ref[:, :, 8] = cmd_w[:, None]
    x_ref[:, :, 9] = v_w[:, 0:1]
    x_ref[:, :, 10] = v_w[:, 1:2]
    res_lin = rng.normal(0, 0.5, (B, 3)); res_ang = rng.normal(0, 0.5, (B, 3))
    if zero_residual or standing:
        res_lin[:] = 0; res_ang[:] = 0
    phase = rng.uniform(0, 1, B)
    contact = np.ones((B, HORIZON, 2)) if standing else mpc_gait_table(phase)

    def q(a):   # everything upstream of the kernel is fp32 in the reference
        a = np.asarray(a, np.float64)
        return a.astype(np.float32).astype(np.float64) if fp32_inputs else a

    ones = np.ones((B, 120))
    inputs = [
        q(x0), ones.copy(), ones.copy(), q(x_ref.reshape(B, 120)), q(dt.reshape(B, 1)),
        np.full((B, 1), float(mass)), np.full((B, 1), float(mu)),
        q(Rb.reshape(B, 9)),                    # row-major flattening (quirk Q1)
        q(I_world.reshape(B, 9)),
        q(body_pos), q(feet[:, 0, :]), q(feet[:, 1, :]),
        q(contact.reshape(B, 20)),              # row-major flattening (quirk Q2)
        np.tile(Q_DEFAULT, (B, 1)), np.tile(R_DEFAULT, (B, 1)),
        q(res_lin), q(res_ang),
    ]
    inputs = [np.ascontiguousarray(a, dtype=np.float64) for a in inputs]
    for a, n in zip(inputs, INPUT_NNZ):
        assert a.shape == (B, n), (a.shape, n)
    return inputs, np.ascontiguousarray(q(Rb))


def nominal_standing_inputs():
    """The reference's only fixture for this path: the nominal standing pose of
    srbd_constraints.py:243-288 / generate_solver_function.py:14-58 (inputs only)."""
    m, mu, dt = 13.5, 0.5, 0.04
    body = np.array([0.0, 0.0, 0.5]); lf = np.array([0.1, 0.05, 0.0]); rf = np.array([0.1, -0.05, 0.0])
    z = np.zeros(240)
    for i in range(HORIZON):
        z[i * 12 + 5] = 0.5
        z[120 + i * 12 + 2] = m * 9.81 / 2
        z[120 + i * 12 + 5] = m * 9.81 / 2
        z[120 + i * 12 + 6:120 + i * 12 + 9] = np.cross(lf - body, z[120 + i * 12:120 + i * 12 + 3])
        z[120 + i * 12 + 9:120 + i * 12 + 12] = np.cross(rf - body, z[120 + i * 12 + 3:120 + i * 12 + 6])
    x0 = np.zeros(12); x0[5] = 0.55
    x_ref = np.zeros(120); x_ref[5::12] = 0.55
    Q = np.array([200, 500, 500, 500, 500, 500, 1, 1, 5, 1, 1, 5], float)
    R = np.array([1e-5] * 6 + [1e-2] * 6)
    I_w = np.diag([0.5413, 0.5200, 0.0691])
    vals = [x0, z[:120], z[120:], x_ref, [dt], [m], [mu], np.eye(3).reshape(9, order="F"),
            I_w.reshape(9, order="F"), body, lf, rf, np.ones(20), Q, R, np.zeros(3), np.zeros(3)]
    return [np.asarray(v, np.float64).reshape(1, -1) for v in vals]


def config1_t1_standing():
    """BASELINE.json configs[0] / SURVEY 8(d) config 1: a single T1 biped (m = 40, mu = 1,
    t1.py:70-75), dt = 0.025, horizon 10 (configuration.py:38-40), Q / R of configuration.py:45,50,
    in the standing pose of generate_solver_function.py:26-58 (body z = 0.5, feet (0.1, +-0.05, 0),
    x0[5] = x_ref[5] = 0.55, identity rotation, zero residuals, all contact), linearisation point
    x = u = 1 as the controller passes it (mpc_controller_cusadi.py:55-56)."""
    body = np.array([0.0, 0.0, 0.5]); lf = np.array([0.1, 0.05, 0.0]); rf = np.array([0.1, -0.05, 0.0])
    x0 = np.zeros(12); x0[5] = 0.55
    x_ref = np.zeros(120); x_ref[5::12] = 0.55
    vals = [x0, np.ones(120), np.ones(120), x_ref, [0.025], [MASS_T1], [MU_T1], np.eye(3).reshape(9, order="F"),
            I_BODY.reshape(9, order="F"), body, lf, rf, np.ones(20), Q_DEFAULT, R_DEFAULT, np.zeros(3), np.zeros(3)]
    return [np.asarray(v, np.float64).reshape(1, -1) for v in vals]
