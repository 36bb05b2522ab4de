"""Batched T1 kinematics: the device code (csrc/srbd_t1_impl.h) against the numpy restatement of the URDF
chain (oracle/t1_kinematics.py).  Parity is UNPINNED to Pinocchio / the serialized CasADi functions (absent offline): what is
checked is (a) the chain constants against the reference's own URDF, (b) the device code against the independent
homogeneous-matrix formulation, (c) the Jacobian against finite differences, (d) known answers of the zero pose, (e) FK o IK."""
import ctypes
import os

import numpy as np
import pytest

from helpers import stack_emul_lib
from oracle import t1_kinematics as t1o

URDF = "/root/reference/model/t1_serial.urdf"


def _emul_fk(q):
    lib = stack_emul_lib()
    q = np.ascontiguousarray(q, np.float32)
    B = q.shape[0]
    p = np.zeros((B, 2, 3), np.float32); J = np.zeros((B, 2, 6, 6), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.srbd_emul_t1_kinematics(q.ctypes.data_as(fp), p.ctypes.data_as(fp), J.ctypes.data_as(fp), B)
    return p, J


def _emul_ik(p):
    lib = stack_emul_lib()
    p = np.ascontiguousarray(p, np.float32)
    q = np.zeros((p.shape[0], 2, 6), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.srbd_emul_t1_ik(p.ctypes.data_as(fp), q.ctypes.data_as(fp), p.shape[0])
    return q


def _random_q(B, seed):
    rng = np.random.default_rng(seed)
    lim = np.array([1.2, 0.5, 0.6, 2.0, 0.8, 0.4])
    return (rng.uniform(-1, 1, (B, 2, 6)) * lim).astype(np.float32)


@pytest.mark.skipif(not os.path.exists(URDF), reason="reference URDF only exists in the build container")
def test_chain_constants_are_the_urdfs():
    u = t1o.chain_from_urdf(URDF)
    assert u["rpy_all_zero"] and u["waist"] == t1o.TRUNK_TO_WAIST
    assert u["chains"]["Left"] == t1o.CHAIN
    assert u["chains"]["Right"] == tuple((n, (o[0], -o[1], o[2]), a) for n, o, a in t1o.CHAIN)
    assert u["sole"] == t1o.SOLE


def test_zero_pose_known_answers():
    """All joints at zero: the sole hangs straight below the hip chain -- the sum of the URDF offsets."""
    p, J = _emul_fk(np.zeros((1, 2, 6), np.float32))
    z = -0.1155 - 0.02 - 0.081854 - 0.134 - 0.28 - 0.012
    assert np.allclose(p[0, 0], [0.0625 - 0.014, 0.106 + 0.00025, z - 0.035192], atol=1e-7)
    assert np.allclose(p[0, 1], [0.0625 - 0.014, -0.106 - 0.00025, z - 0.03519], atol=1e-7)
    # angular rows = joint axes in the trunk frame: y x z y y x
    axes = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0]], np.float32).T
    assert np.array_equal(J[0, 0, 3:], axes) and np.array_equal(J[0, 1, 3:], axes)
    # a hip pitch rotation moves the sole forward by the leg length: dp_x / dq0 = -(p_z - z_hip)
    assert np.isclose(J[0, 0, 0, 0], p[0, 0, 2] - (-0.1155), atol=1e-6)


def test_device_code_matches_the_homogeneous_matrix_restatement():
    q = _random_q(64, 1)
    p, J = _emul_fk(q)
    pr, Jr = t1o.fk_jac_batch(q)
    assert np.abs(p - pr.astype(np.float32)).max() <= 6e-8          # fp64 inside, one fp32 rounding at the end
    assert np.abs(J - Jr.astype(np.float32)).max() <= 6e-8


def test_jacobian_is_the_derivative_of_the_position():
    q = _random_q(8, 2).astype(np.float64)
    _, J = t1o.fk_jac_batch(q)
    h = 1e-6
    for j in range(6):
        dq = np.zeros_like(q); dq[:, :, j] = h
        pp, _ = t1o.fk_jac_batch(q + dq); pm, _ = t1o.fk_jac_batch(q - dq)
        assert np.abs((pp - pm) / (2 * h) - J[:, :, :3, j]).max() < 1e-8


def test_inverse_kinematics_round_trip_and_restatement():
    rng = np.random.default_rng(3)
    B = 32
    foot = np.zeros((B, 2, 3), np.float32)
    foot[:, :, 0] = 0.0485 + rng.uniform(-0.15, 0.15, (B, 2))
    foot[:, 0, 1] = 0.10625 + rng.uniform(-0.05, 0.08, B); foot[:, 1, 1] = -0.10625 - rng.uniform(-0.05, 0.08, B)
    foot[:, :, 2] = -0.60 + rng.uniform(-0.05, 0.05, (B, 2))
    q = _emul_ik(foot)
    for leg in range(2):
        assert np.abs(q[:, leg] - t1o.analytical_ik(foot[:, leg], leg)).max() < 5e-6
    assert np.all(q[:, :, 2] == 0) and np.all(q[:, :, 5] == 0)
    # The reference's IK measures the pitch angles with the opposite sign of the URDF's y axes (hip_pitch = atan2(x, -z) - beta
    # is positive for a foot in FRONT of the hip, while a positive rotation about +y moves it BACK): it is mirrored here as
    # it is (t1.py:139-213), so the round trip closes through the URDF chain with the three pitch joints negated -- to within ~2
    # centimetre, which is the accuracy of the reference's closed form itself: it solves the hip roll as if it preceded the
    # hip pitch (the URDF chain is pitch -> roll -> yaw), lumps the 2 cm / 8 cm hip offsets into the thigh length, and uses
    # the left sole offset for both legs.  The exact check of the kernel is the restatement above.
    p, _ = _emul_fk(q * np.array([-1, 1, 1, -1, -1, 1], np.float32))
    assert np.abs(p - foot).max() < 2.5e-2


@pytest.mark.gpu
def test_t1_kernels_on_the_gpu():
    torch = pytest.importorskip("torch")
    import biped_pympc_b200 as bp
    q = _random_q(4099, 5)
    qd = torch.as_tensor(q, device="cuda:0")
    p, J = bp.t1_kinematics(qd)
    pe, Je = _emul_fk(q)
    assert np.abs(p.cpu().numpy() - pe).max() <= 6e-8 and np.abs(J.cpu().numpy() - Je).max() <= 6e-8
    pr, Jr = t1o.fk_jac_batch(q[:200])
    assert np.abs(p[:200].cpu().numpy() - pr).max() <= 1e-7 and np.abs(J[:200].cpu().numpy() - Jr).max() <= 1e-7
    qi = bp.t1_inverse_kinematics(p)
    assert np.abs(qi.cpu().numpy() - _emul_ik(pe)).max() < 2e-5
    # the host mirror of the reference's T1 class: same calls, every env instead of env 0
    t1 = bp.T1B200(4099, "cuda:0")
    for leg in (0, 1):
        t1.forward_kinematics(qd[:, leg], leg)
        assert torch.equal(t1.get_p0e(leg), p[:, leg]) and torch.equal(t1.contact_jacobian(leg), J[:, leg])
        assert tuple(t1.contact_jacobian(leg).shape) == (4099, 6, 6)
    assert tuple(t1.analytical_IK(p[:, 0], 0).shape) == (4099, 6)
    assert bp._native.load().srbd_t1_kinematics(None, None, None, 4, None) == -1
