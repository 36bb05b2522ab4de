"""CPU tests of the host-side mirrors: no compute call is made; the product refuses to run without a CUDA device
instead of falling back to anything (there is no CPU path and nothing under oracle/ is ever imported by it)."""
import os
import re

import pytest

import biped_pympc_b200 as bp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_control_stack_mirrors_refuse_to_run_on_the_cpu():
    import torch
    with pytest.raises(bp.SrbdError):
        bp.BipedControllerB200(4, "cpu")
    with pytest.raises(bp.SrbdError):
        bp.GaitGeneratorB200(4, 10, 0.001, torch.full((4,), 0.025), torch.zeros(4, 2), torch.full((4, 2), 5), "cpu")
    with pytest.raises(TypeError):      # host tensors are rejected by the kernel wrappers, not silently copied
        bp._native.reference_trajectory([torch.zeros(4, 3)] * 9, 0.01, torch.zeros(4, 3), torch.zeros(4), torch.ones(4, dtype=torch.bool))
    with pytest.raises(TypeError):
        bp.marshal([torch.zeros(4, 12)] * 12, 40.0, 1.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "biped-pympc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "srbd_oracle" not in text, f


def test_mirror_classes_expose_the_reference_surface():
    """Method names of the reference classes the mirrors stand in for (mpc_wrapper.py:3-67, biped_controller.py:83-176,
    gait_generator.py:75-252, base_controller.py:164-266, mpc_controller_cusadi.py:23-43)."""
    for name in ("set_command", "update_state", "run_mpc", "run_lowlevel", "get_action", "reset", "update_mpc_sampling_time",
                 "set_swing_parameters", "set_srbd_accel", "set_desired_state", "command_joint_torque", "ground_reaction_wrench",
                 "contact_state", "contact_phase", "swing_state", "swing_phase", "foot_placement", "foot_placement_b",
                 "ref_foot_pos_b", "ref_foot_vel_b", "foot_pos_b", "foot_vel_b"):
        assert hasattr(bp.BipedControllerB200, name), name
    for name in ("reset", "update_phase", "update_sampling_time", "get_contact_sub_phase", "get_swing_sub_phase", "contact_bool",
                 "swing_bool", "mpc_gait"):
        assert hasattr(bp.GaitGeneratorB200, name), name
    for name in ("init_solver", "run", "reset", "build_state"):
        assert hasattr(bp.MPCControllerB200, name), name


@pytest.mark.parametrize("B,sms", [(1, 148), (100, 148), (148, 148), (149, 148), (2368, 148), (2369, 148), (4096, 148), (65536, 148), (5000, 132)])
def test_persistent_kernel_env_numbering_covers_every_env_once(B, sms):
    """The fp64 fast kernel is persistent (csrc/srbd_pdipm.cu): G = min(SMs, B) CTAs of 16 warps, warp w of CTA s solves
    envs e = w * G + s, then e += 16 G.  Every env exactly once; within a round the envs fill warps 0.. of ALL CTAs first,
    so a partial round loads every SM (and every scheduler: warp w runs on scheduler w % 4) evenly."""
    G, W = min(sms, B), 16
    seen = []
    per_cta_last_round = {}
    for s_ in range(G):
        for w in range(W):
            e = w * G + s_
            while e < B:
                seen.append(e)
                if e >= (B // (W * G)) * W * G:
                    per_cta_last_round[s_] = per_cta_last_round.get(s_, 0) + 1
                e += W * G
    assert sorted(seen) == list(range(B))
    if per_cta_last_round:
        assert max(per_cta_last_round.values()) - min(per_cta_last_round.get(s_, 0) for s_ in range(G)) <= 1
