"""Developer tool (GPU): per-quantity worst absolute error of BipedControllerB200 against the recording of the
reference stack (tests/golden/srbd_ref_loop_v1.npz).  Usage: python tests/tools/loop_errors.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import biped_pympc_b200 as bp  # noqa: E402


def run(g, inject=True, dev="cuda:0"):
    B, T, dec = g["state"].shape[1], g["state"].shape[0], int(g["decimation"])
    c = bp.BipedControllerB200(B, dev, dt=float(g["dt"]), dt_mpc=float(g["dt_mpc"]), decimation=dec, ssp_durations=int(g["ssp"]),
                               dsp_durations=int(g["dsp"]), swing_height=float(g["swing_height"]), Q=g["Q"], R=g["R"],
                               swing_reference_frame="world" if int(g.get("world_frame", 0)) else "base", gait_id=int(g.get("gait_id", 2)))
    c.mpc_controller.reproduce_q_stride_bug = True
    c.gait_generator.gait_phase[:] = torch.as_tensor(g["phase0"], device=dev)
    err, k = {}, 0
    D = lambda t: torch.as_tensor(t, device=dev)  # noqa: E731

    def cmp(name, got, want):
        got = got.detach().float().cpu().numpy().reshape(want.shape)
        fin = np.isfinite(want)
        assert np.array_equal(fin, np.isfinite(got)), name
        err[name] = max(err.get(name, 0.0), float(np.abs(got[fin] - want[fin]).max()) if fin.any() else 0.0)

    sed, ld, cmd, sw = c.state_estimator.data, c.leg_controller.data, c.leg_controller.command, c.swing_leg_controller
    for t in range(T):
        if "reset_step" in g and t == int(g["reset_step"]):
            c.reset(torch.as_tensor(g["reset_envs"], device=dev))
        if "knob_step" in g and t == int(g["knob_step"]):
            c.update_mpc_sampling_time(D(g["knob_dt_mpc"])); c.set_swing_parameters(D(g["knob_foot_height"]), D(g["knob_cp1"]), D(g["knob_cp2"]))
            c.set_srbd_accel(D(g["knob_res_lin"]), D(g["knob_res_ang"]))
        c.set_command(D(g["twist"][t]), D(g["height"][t]))
        c.update_state(D(g["state"][t]))
        for n in ("root_euler", "rotation_body", "root_velocity_w", "root_angular_velocity_w", "foot_position"):
            cmp(n, getattr(sed, n), g[n][t])
        for n in ("p", "v", "J", "contact_bool", "swing_bool"):
            cmp("leg_" + n, getattr(ld, n), g["leg_" + n][t])
        cmp("contact_phase", ld.contact_phase, g["contact_phase"][t]); cmp("swing_phase", ld.swing_phase, g["swing_phase"][t])
        if t % dec == 0:
            c.run_mpc()
            cmp("mpc_contact_table", c.gait_generator.contact_table, g["mpc_contact_table"][k])
            cmp("mpc_x_ref", c.mpc_controller.x_ref, g["mpc_x_ref"][k]); cmp("mpc_x0", c.mpc_controller.x0, g["mpc_x0"][k])
            k += 1
        cmp("wrench", cmd.feedfowardforce, g["wrench"][t])
        if inject:
            cmd.feedfowardforce.copy_(D(g["wrench"][t]))
        c.run_lowlevel()
        for n in ("swing_time_remaining", "first_swing", "foot_placement", "foot_placement_b", "p0"):
            cmp(n, getattr(sw, n), g[n][t])
        for n in ("pDes", "vDes", "tau", "qDes", "qdDes", "kpjoint", "kdjoint"):
            cmp("cmd_" + n, getattr(cmd, n), g["cmd_" + n][t])
        cmp("action", c.get_action(), g["action"][t])
        cmp("gait_phase_after", c.gait_generator.gait_phase, g["gait_phase_after"][t])
    c.close()
    return err


if __name__ == "__main__":
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "srbd_ref_loop_v1.npz")))
    for inject in (True, False):
        print("inject recorded wrench:", inject)
        for n, v in run(g, inject).items():
            print("  %-26s %.3e   (scale %.3g)" % (n, v, float(np.nanmax(np.abs(g[n][np.isfinite(g[n])])))))
