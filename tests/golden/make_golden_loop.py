"""Generates tests/golden/srbd_ref_loop_v1.npz by running the reference's WHOLE control stack.

The reference's top-level wrapper (biped_pympc/mpc_wrapper.py::MPCController -> BipedController ->
GaitGenerator, StateEstimator, LegController(HECTOR), SwingLegController, MPCControllerCusadi) is imported
unmodified from /root/reference and stepped on the CPU the way its users step it, once per control step:

    set_command -> update_state -> [run_mpc every `decimation` steps] -> run_lowlevel -> get_action

with the walking gait (gait_id=2); two short variants cover the `world` swing reference frame and the
standing gait (gait_id=1).  As in make_golden_controller.py the only stand-ins are the `casadi`
module and `CusadiFunction` (CPU evaluation of the two rebuilt functions).  The sensor input is an open-loop,
seeded random walk around a standing posture (this is a recording of the controller's arithmetic, not a
simulation).  Every intermediate quantity the device-side mirror (biped-pympc_b200/biped_controller.py)
produces is recorded per step.
Run:  python tests/golden/make_golden_loop.py      (about a minute)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_ref as mgr  # noqa: E402
from make_golden_controller import CpuCusadiFunction  # noqa: E402

B, STEPS = 6, 160


def build_reference_wrapper(num_envs, decimation=10, reference_frame="base", gait_id=2):
    SRBD_Constraints, sparse_pdipm_multiple_iterations = mgr.import_reference()
    import casadi
    wl = mgr.load_workload()
    qp_former = SRBD_Constraints(10).qp_mat_former
    H, f, A, b, G, d = qp_former(*[a[0] for a in wl.nominal_standing_inputs()])
    (Hc, Hr), (Ac, Ar), (Gc, Gr) = H.sparsity().get_ccs(), A.sparsity().get_ccs(), G.sparsity().get_ccs()
    solver = sparse_pdipm_multiple_iterations(240, 140, 160, Hr, Hc, Ar, Ac, Gr, Gc, beta=1e-8, num_iters=5)
    table = {"srbd_qp_mat.casadi": qp_former, "mpc_multiple_iter_5_solver_240v_140eq_160ineq.casadi": solver}
    casadi.Function.load = staticmethod(lambda path: table[os.path.basename(path)])
    fake = types.ModuleType("biped_pympc.cusadi")
    fake.CASADI_FUNCTION_DIR = "/nonexistent/casadi_functions"
    fake.CusadiFunction = CpuCusadiFunction
    sys.modules["biped_pympc.cusadi"] = fake
    from biped_pympc.configuration.configuration import ControllerConf, MPCConf
    from biped_pympc.mpc_wrapper import MPCController
    mpc_cfg = MPCConf(print_solve_time=False, decimation=decimation)
    mpc_cfg.Q, mpc_cfg.R = mpc_cfg.Q.cpu(), mpc_cfg.R.cpu()
    cfg = ControllerConf(swing_reference_frame=reference_frame)
    return MPCController(cfg, mpc_cfg, num_envs, torch.device("cpu"), gait_id=gait_id), cfg, mpc_cfg


def quat_from_euler(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy], 1)


def main(B=B, STEPS=STEPS, reference_frame="base", gait_id=2, name="srbd_ref_loop_v1.npz", seed=20261018, events=False):
    ctrl, cfg, mpc_cfg = build_reference_wrapper(B, reference_frame=reference_frame, gait_id=gait_id)
    bc = ctrl.biped_controller
    rng = np.random.default_rng(seed)
    bc.gait_generator.gait_phase[:] = torch.as_tensor(rng.uniform(0, 1, B), dtype=torch.float32)
    out = {"phase0": bc.gait_generator.gait_phase.numpy().copy(), "decimation": np.array(mpc_cfg.decimation),
           "dt": np.array(mpc_cfg.dt), "dt_mpc": np.array(mpc_cfg.dt_mpc), "ssp": np.array(cfg.ssp_durations),
           "dsp": np.array(cfg.dsp_durations), "swing_height": np.array(cfg.swing_height),
           "mass": np.array(bc.mpc_controller.mass), "mu": np.array(bc.mpc_controller.mu),
           "Q": mpc_cfg.Q.numpy(), "R": mpc_cfg.R.numpy(), "world_frame": np.array(int(reference_frame == "world")),
           "gait_id": np.array(gait_id)}
    # open-loop sensor stream: a smooth random walk around a crouched standing posture
    q_nom = np.array([0.0, 0.0, 0.6, -1.2, 0.6] * 2)
    q = q_nom + rng.normal(0, 0.05, (B, 10))
    pos = np.stack([rng.normal(0, 1, B), rng.normal(0, 1, B), 0.55 + rng.normal(0, 0.01, B)], 1)
    eul = np.stack([rng.normal(0, 0.05, B), rng.normal(0, 0.05, B), rng.uniform(-3, 3, B)], 1)
    twist = np.stack([rng.uniform(-0.5, 0.5, B), rng.uniform(-0.3, 0.3, B), rng.uniform(-0.5, 0.5, B)], 1)
    twist[0, 0] = 0.004                    # stationary branch of the reference trajectory
    height = 0.55 + rng.normal(0, 0.01, B)
    rec = {}
    put = lambda k, v: rec.setdefault(k, []).append(np.array(v.detach().cpu().numpy() if torch.is_tensor(v) else v))  # noqa: E731
    for t in range(STEPS):
        if events and t == 35:            # episode reset of two envs (BipedController.reset, biped_controller.py:83-88)
            ctrl.reset(torch.tensor([1, 3]))
            out["reset_step"], out["reset_envs"] = np.array(t), np.array([1, 3])
        if events and t == 50:            # the RL-side knobs (mpc_wrapper.py:44-67)
            new_dt = torch.as_tensor(rng.uniform(0.02, 0.05, B), dtype=torch.float32)
            fh = torch.as_tensor(rng.uniform(0.05, 0.15, B), dtype=torch.float32)
            c1 = torch.as_tensor(rng.uniform(0.2, 0.45, B), dtype=torch.float32)
            c2 = torch.as_tensor(rng.uniform(0.55, 0.8, B), dtype=torch.float32)
            ra = torch.as_tensor(rng.normal(0, 0.5, (2, B, 3)), dtype=torch.float32)
            ctrl.update_mpc_sampling_time(new_dt.clone()); ctrl.set_swing_parameters(fh.clone(), c1.clone(), c2.clone())
            ctrl.set_srbd_accel(ra[0].clone(), ra[1].clone())
            out.update(knob_step=np.array(t), knob_dt_mpc=new_dt.numpy(), knob_foot_height=fh.numpy(), knob_cp1=c1.numpy(),
                       knob_cp2=c2.numpy(), knob_res_lin=ra[0].numpy(), knob_res_ang=ra[1].numpy())
        qd = rng.normal(0, 0.5, (B, 10))
        q = q + 0.001 * qd + rng.normal(0, 0.002, (B, 10))
        lin_b = rng.normal(0, 0.3, (B, 3))
        ang_b = rng.normal(0, 0.2, (B, 3))
        pos = pos + 0.001 * rng.normal(0, 0.3, (B, 3))
        eul = eul + 0.001 * rng.normal(0, 0.2, (B, 3))
        quat = quat_from_euler(eul[:, 0], eul[:, 1], eul[:, 2]) * rng.uniform(0.98, 1.02, (B, 1))   # not quite unit
        tau_meas = rng.normal(0, 1.0, (B, 10))
        state = torch.as_tensor(np.concatenate([pos, quat, lin_b, ang_b, q, qd, tau_meas], 1), dtype=torch.float32)
        put("state", state)
        put("twist", twist.astype(np.float32)); put("height", height.astype(np.float32))
        ctrl.set_command(torch.as_tensor(twist, dtype=torch.float32), torch.as_tensor(height, dtype=torch.float32))
        ctrl.update_state(state)
        sed, ld = bc.state_estimator.data, bc.leg_controller.data
        for n in ("root_euler", "rotation_body", "root_velocity_w", "root_angular_velocity_w", "foot_position"):
            put(n, getattr(sed, n))
        for n in ("p", "v", "J", "contact_bool", "swing_bool"):
            put("leg_" + n, getattr(ld, n))
        put("contact_phase", bc.contact_phase); put("swing_phase", bc.swing_phase)
        mpc_step = (t % mpc_cfg.decimation == 0)
        if mpc_step:
            ctrl.run_mpc()
            put("mpc_contact_table", bc.mpc_controller.contact_table)
            put("mpc_x_ref", bc.mpc_controller.x_ref)
            put("mpc_x0", bc.mpc_controller.x0)
        put("wrench", bc.leg_controller.command.feedfowardforce)
        ctrl.run_lowlevel()
        slc, cmd = bc.swing_leg_controller, bc.leg_controller.command
        put("swing_time_remaining", slc.swing_time_remaining); put("first_swing", slc.first_swing)
        put("foot_placement", slc.foot_placement); put("foot_placement_b", slc.foot_placement_b); put("p0", slc.p0)
        for n in ("pDes", "vDes", "tau", "qDes", "qdDes", "kpjoint", "kdjoint"):
            put("cmd_" + n, getattr(cmd, n))
        put("action", ctrl.get_action())
        put("gait_phase_after", bc.gait_generator.gait_phase)
        if t % 40 == 0:
            print("step", t, "action[0]", ctrl.get_action()[0].numpy().round(2).tolist(), flush=True)
    for k, v in rec.items():
        out[k] = np.stack(v)
    out["torque_limit"] = bc.torque_limit.numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("saved", name, {k: v.shape for k, v in out.items() if k in ("state", "action", "leg_J", "wrench")})


if __name__ == "__main__":
    main()                                                                                  # walking, base frame (defaults)
    main(B=4, STEPS=70, reference_frame="world", name="srbd_ref_loop_v1_world.npz", seed=7)  # swing targets in the world frame
    main(B=3, STEPS=30, gait_id=1, name="srbd_ref_loop_v1_standing.npz", seed=9)             # standing gait (swing time 0)
    main(B=4, STEPS=80, name="srbd_ref_loop_v1_events.npz", seed=11, events=True)            # resets and RL-side knobs mid-run
