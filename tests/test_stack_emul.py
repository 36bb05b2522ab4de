"""CPU tests of the control-stack kernels: their per-env bodies (csrc/srbd_stack_impl.h), compiled for the host
(tests/emul/stack_emul.cpp, -ffp-contract=off), run against the recordings of the reference's own code -- the same
comparisons tests/test_gpu_parity.py makes on the GPU, so that a regression in this arithmetic is caught without one."""
import ctypes
import os

import numpy as np
import pytest

import biped_pympc_b200 as bp
from helpers import stack_emul_lib

HERE = os.path.dirname(os.path.abspath(__file__))
F, I32 = np.float32, np.int32
fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", name)))


def test_gait_body_matches_the_reference_gait_generator():
    g = _load("srbd_ref_gait_v1.npz")
    lib = stack_emul_lib()
    B, H, T = g["ssp"].shape[0], int(g["horizon"]), g["table"].shape[0]
    phase = g["phase0"].astype(F).copy()
    ssp, dsp = np.ascontiguousarray(g["ssp"], I32), np.ascontiguousarray(g["dsp"], I32)
    dt_mpc = g["dt_mpc0"].astype(F).copy()
    c, s, tab = np.zeros((B, 2), F), np.zeros((B, 2), F), np.zeros((B, H, 2), I32)
    for t in range(T):
        if t == int(g["switch_step"]):
            dt_mpc = g["dt_mpc1"].astype(F).copy()
        assert np.array_equal(phase, g["phase"][t], equal_nan=True), t
        lib.srbd_emul_gait(fp(phase), fp(dt_mpc), fp(ssp), fp(dsp), ctypes.c_double(float(g["dt"])), 1, H, fp(c), fp(s), fp(tab), None, B)
        assert np.array_equal(c, g["contact_sub"][t], equal_nan=True) and np.array_equal(s, g["swing_sub"][t], equal_nan=True), t
        assert np.array_equal(tab, g["table"][t]), t
    assert np.array_equal(phase, g["phase"][T], equal_nan=True)


@pytest.mark.parametrize("tick", [0, 1])
def test_state_builder_body_matches_the_reference_controller(tick):
    g = _load("srbd_ref_controller_v1.npz")
    lib = stack_emul_lib()
    t = f"t{tick}_"
    B = g[t + "x0"].shape[0]
    ins = [np.ascontiguousarray(g[t + n], F) for n in ("root_euler", "root_position", "root_angular_velocity_w", "root_velocity_w",
                                                        "rotation_body", "desired_velocity_b", "desired_angular_velocity_b",
                                                        "desired_height", "dt_mpc")]
    arr = (ctypes.c_void_p * 9)(*[a.ctypes.data for a in ins])
    wpd, yaw = g[t + "world_position_desired_in"].astype(F).copy(), g[t + "yaw_desired_in"].astype(F).copy()
    first = g[t + "first_run_in"].astype(np.uint8).copy()
    x0, x_ref = np.zeros((B, 12), F), np.zeros((B, 10, 12), F)
    lib.srbd_emul_reference_trajectory(arr, ctypes.c_double(float(g["decimation"]) * float(g["dt"])), fp(wpd), fp(yaw), fp(first),
                                       fp(x0), fp(x_ref), B)
    assert np.array_equal(x0, g[t + "x0"]) and not first.any()
    assert np.array_equal(wpd, g[t + "world_position_desired_out"]) and np.array_equal(yaw, g[t + "yaw_desired_out"])
    want = g[t + "x_ref"]
    exact = [k for k in range(12) if k not in (3, 4, 9, 10)]               # the others carry R @ v_b (a torch.bmm there)
    assert np.array_equal(x_ref[:, :, exact], want[:, :, exact])
    assert np.abs(x_ref - want).max() <= 2.4e-7 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("fixture", ["srbd_ref_loop_v1.npz", "srbd_ref_loop_v1_world.npz", "srbd_ref_loop_v1_standing.npz",
                                     "srbd_ref_loop_v1_events.npz"])
def test_control_step_bodies_match_the_reference_stack(fixture):
    """update_state + lowlevel bodies over the recorded sensor stream, the recorded MPC wrench injected (the solve itself is
    covered by the PDIPM emulator tests): every intermediate quantity of the reference stack, flags exactly, fp32
    arithmetic to a few ulp."""
    g = _load(fixture)
    lib = stack_emul_lib()
    T, B = g["state"].shape[:2]
    gait2 = int(g.get("gait_id", 2)) == 2
    ssp = np.tile(np.array([int(g["ssp"])] * 2 if gait2 else [0, 0], I32), (B, 1))
    dsp = np.tile(np.array([int(g["dsp"])] * 2 if gait2 else [5, 5], I32), (B, 1))
    z = lambda *s: np.zeros(s, F)  # noqa: E731
    t_ = dict(gait_phase=g["phase0"].astype(F).copy(), dt_mpc=np.full(B, float(g["dt_mpc"]), F), ssp_durations=ssp, dsp_durations=dsp,
              swing_time_remaining=(ssp * np.float32(g["dt_mpc"])).astype(F), first_swing=np.ones((B, 2), np.uint8), p0=z(B, 2, 3),
              foot_height=np.full(B, float(g["swing_height"]), F), cp1=np.full(B, 1 / 3, F), cp2=np.full(B, 2 / 3, F),
              desired_velocity_b=z(B, 3), root_position=z(B, 3), root_euler=z(B, 3), rotation_body=z(B, 3, 3), root_velocity_b=z(B, 3),
              root_angular_velocity_b=z(B, 3), root_velocity_w=z(B, 3), root_angular_velocity_w=z(B, 3), foot_position=z(B, 2, 3),
              q=z(B, 2, 5), qd=z(B, 2, 5), p=z(B, 2, 3), v=z(B, 2, 3), J=z(B, 2, 6, 5), contact_phase=z(B, 2), swing_phase=z(B, 2),
              contact_bool=z(B, 2), swing_bool=z(B, 2), feedforward_force=z(B, 2, 6), foot_placement=z(B, 2, 3), foot_placement_b=z(B, 2, 3),
              pDes=z(B, 2, 3), vDes=z(B, 2, 3), tau=z(B, 2, 5), qDes=z(B, 2, 5), qdDes=z(B, 2, 5), kpjoint=z(B, 2, 5), kdjoint=z(B, 2, 5),
              action=z(B, 10))
    buf = bp.biped_controller._LoopBuffers()
    for n, a in t_.items():
        setattr(buf, n, a.ctypes.data)
    H = bp.biped_controller.HectorModel
    kp, kd, lim = [(ctypes.c_float * len(v))(*v) for v in (H.kp, H.kd, H.torque_limit)]
    err = {}

    def cmp(name, got, want):
        got = got.reshape(want.shape).astype(np.float32)
        fin = np.isfinite(want)
        assert np.array_equal(fin, np.isfinite(got)), name
        err[name] = max(err.get(name, 0.0), float(np.abs(got[fin] - want[fin]).max()) if fin.any() else 0.0)

    for t in range(T):
        if "reset_step" in g and t == int(g["reset_step"]):
            ids = g["reset_envs"]
            t_["gait_phase"][ids] = 0.0
            t_["first_swing"][ids] = 1
        if "knob_step" in g and t == int(g["knob_step"]):
            t_["dt_mpc"][:] = g["knob_dt_mpc"]; t_["foot_height"][:] = g["knob_foot_height"]
            t_["cp1"][:] = g["knob_cp1"]; t_["cp2"][:] = g["knob_cp2"]
        t_["desired_velocity_b"][:, :2] = g["twist"][t][:, :2]
        state = np.ascontiguousarray(g["state"][t], F)
        assert lib.srbd_emul_update_state(ctypes.byref(buf), fp(state), state.shape[1], B) == 0
        for n in ("root_euler", "rotation_body", "root_velocity_w", "root_angular_velocity_w", "foot_position"):
            cmp(n, t_[n], g[n][t])
        for n in ("p", "v", "J", "contact_bool", "swing_bool"):
            cmp("leg_" + n, t_[n], g["leg_" + n][t])
        cmp("contact_phase", t_["contact_phase"], g["contact_phase"][t]); cmp("swing_phase", t_["swing_phase"], g["swing_phase"][t])
        t_["feedforward_force"][:] = g["wrench"][t]
        assert lib.srbd_emul_lowlevel(ctypes.byref(buf), kp, kd, lim, ctypes.c_double(float(g["dt"])), int(g.get("world_frame", 0)), 1, B) == 0
        for n in ("swing_time_remaining", "first_swing", "foot_placement", "foot_placement_b", "p0"):
            cmp(n, t_[n], g[n][t].astype(np.float32))
        for n in ("pDes", "vDes", "tau", "qDes", "qdDes", "kpjoint", "kdjoint"):
            cmp("cmd_" + n, t_[n], g["cmd_" + n][t])
        cmp("action", t_["action"], g["action"][t])
        cmp("gait_phase_after", t_["gait_phase"], g["gait_phase_after"][t])
    scale = lambda n: max(1.0, float(np.abs(g[n][np.isfinite(g[n])]).max())) if np.isfinite(g[n]).any() else 1.0  # noqa: E731
    exact = ("leg_contact_bool", "leg_swing_bool", "contact_phase", "swing_phase", "swing_time_remaining", "first_swing", "cmd_kpjoint",
             "cmd_kdjoint", "gait_phase_after")
    tol = {"cmd_vDes": 5e-6, "cmd_tau": 4e-6, "action": 1.5e-5}
    for n, v in err.items():
        if n in exact:
            assert v == 0.0, (n, v)
        else:
            assert v <= tol.get(n, 2e-6) * scale(n), (n, v, scale(n))
