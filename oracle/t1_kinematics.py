"""CPU restatement of the T1 leg kinematics.  TEST INFRASTRUCTURE ONLY (tests/ may import it; the product never does).

What it restates: the four functions biped_pympc/casadi/t1_kinematics.py:39-75 generates with Pinocchio from
model/t1_serial.urdf -- position of `<side>_foot_sole_link` in the trunk frame (`fk_func`, :39-54) and its
LOCAL_WORLD_ALIGNED frame Jacobian restricted to the six leg joints (`jacobian_func`, :56-75) with every other joint at
zero -- which core/robot/t1.py:96-138 evaluates on the CPU, and T1.analytical_IK (t1.py:139-213).

PARITY UNPINNED: Pinocchio, CasADi and the serialized .casadi functions are absent (pyproject.toml lists pinocchio
unpinned; the files are downloaded by script/setup.sh).  The published algorithm is restated instead: a URDF serial
chain is a product of homogeneous transforms T = prod_j [Trans(origin_j) Rot(rpy_j) Rot(axis_j, q_j)], the frame
Jacobian column of a revolute joint is [a_j x (p - o_j); a_j] in world-aligned axes.  The chain constants below are the
URDF's; tests/test_t1_kinematics.py re-reads them from /root/reference/model/t1_serial.urdf when it is present, and
checks the Jacobian against finite differences of the position.

Formulated with 4x4 homogeneous matrices (numpy), i.e. not the way the CUDA code accumulates (R, p)."""
from __future__ import annotations

import numpy as np

# (joint name without side, origin xyz for the LEFT leg, axis); right leg: y components of the origins change sign
TRUNK_TO_WAIST = (0.0625, 0.0, -0.1155)      # joint "Waist" (axis z, held at zero)
CHAIN = (
    ("Hip_Pitch", (0.0, 0.106, 0.0), (0, 1, 0)),
    ("Hip_Roll", (0.0, 0.0, -0.02), (1, 0, 0)),
    ("Hip_Yaw", (0.0, 0.0, -0.081854), (0, 0, 1)),
    ("Knee_Pitch", (-0.014, 0.0, -0.134), (0, 1, 0)),
    ("Ankle_Pitch", (0.0, 0.0, -0.28), (0, 1, 0)),
    ("Ankle_Roll", (0.0, 0.00025, -0.012), (1, 0, 0)),
)
SOLE = {"Left": (0.0, 0.0, -0.035192), "Right": (0.0, 0.0, -0.03519)}


def _trans(v):
    T = np.eye(4); T[:3, 3] = v
    return T


def _rot(axis, q):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)      # Rodrigues
    return T


def chain_for(leg):
    """[(origin, axis)] of the six joints and the sole offset for leg 0 (left) / 1 (right)."""
    side = -1.0 if leg else 1.0
    joints = [((o[0], side * o[1], o[2]), ax) for _, o, ax in CHAIN]
    return joints, SOLE["Right" if leg else "Left"]


def fk_jac(q, leg):
    """q (6,) -> (p (3,), J (6,6)) for one leg, float64."""
    joints, sole = chain_for(leg)
    T = _trans(TRUNK_TO_WAIST)
    org, axs = [], []
    for (o, ax), qi in zip(joints, np.asarray(q, float)):
        T = T @ _trans(o)
        org.append(T[:3, 3].copy()); axs.append(T[:3, :3] @ np.asarray(ax, float))
        T = T @ _rot(ax, qi)
    T = T @ _trans(sole)
    p = T[:3, 3].copy()
    J = np.zeros((6, 6))
    for j in range(6):
        J[:3, j] = np.cross(axs[j], p - org[j]); J[3:, j] = axs[j]
    return p, J


def fk_jac_batch(q):
    """q (B,2,6) -> p (B,2,3), J (B,2,6,6)"""
    q = np.asarray(q, float)
    B = q.shape[0]
    p, J = np.zeros((B, 2, 3)), np.zeros((B, 2, 6, 6))
    for e in range(B):
        for leg in range(2):
            p[e, leg], J[e, leg] = fk_jac(q[e, leg], leg)
    return p, J


def analytical_ik(p, leg):
    """t1.py:139-213 restated in float32 numpy (the reference's torch code is fp32).  p (B,3) -> q (B,6)."""
    f = np.float32
    p = np.asarray(p, f)
    side = f(-1.0 if leg else 1.0)
    hip = np.array([0.0625 + 0.0, side * (0.0 + 0.106), -0.1155 + 0.0], f)
    ee = np.array([0.0 + 0.0, side * (0.00025 + 0.0), -0.035192], f)
    L1 = f(-(-0.02 + -0.081854 + -0.134)); L2 = f(-(-0.28 + -0.012)); kx = f(-0.014)
    v = p - hip - ee
    roll = np.arctan2(v[:, 1], -v[:, 2]).astype(f)
    c, s = np.cos(roll).astype(f), np.sin(roll).astype(f)
    xs = v[:, 0] - kx
    zs = -v[:, 1] * s + v[:, 2] * c
    d = np.sqrt(xs * xs + zs * zs).astype(f)
    cb = np.clip((L1 * L1 + d * d - L2 * L2) / (f(2) * L1 * d + f(1e-6)), -1, 1).astype(f)
    beta = np.arccos(cb).astype(f)
    ck = np.clip((L1 * L1 + L2 * L2 - d * d) / (f(2) * L1 * L2 + f(1e-6)), -1, 1).astype(f)
    knee = (f(np.pi) - np.arccos(ck)).astype(f)
    alpha = np.arctan2(xs, -zs).astype(f)
    hp = alpha - beta
    q = np.zeros((p.shape[0], 6), f)
    q[:, 0] = hp; q[:, 1] = roll; q[:, 3] = knee; q[:, 4] = -(hp + knee)
    return q


def chain_from_urdf(path):
    """Re-derive (TRUNK_TO_WAIST, CHAIN, SOLE) from the URDF (used by the tests to pin the constants above)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    joints = {j.get("name"): j for j in root.findall("joint")}
    xyz = lambda j: tuple(float(v) for v in j.find("origin").get("xyz").split())  # noqa: E731
    rpy = lambda j: tuple(float(v) for v in (j.find("origin").get("rpy") or "0 0 0").split())  # noqa: E731
    axis = lambda j: tuple(int(float(v)) for v in j.find("axis").get("xyz").split())  # noqa: E731
    out = {"waist": xyz(joints["Waist"]), "chains": {}, "sole": {}, "rpy_all_zero": True}
    for side in ("Left", "Right"):
        ch = []
        for name, _, _ in CHAIN:
            j = joints[f"{side}_{name}"]
            ch.append((name, xyz(j), axis(j)))
            out["rpy_all_zero"] &= all(v == 0 for v in rpy(j))
        out["chains"][side] = tuple(ch)
        out["sole"][side] = xyz(joints[f"{side.lower()}_foot_sole_joint"])
        # the chain must be serial: each joint's parent is the previous joint's child
        prev = "Waist"
        for name, _, _ in CHAIN:
            j = joints[f"{side}_{name}"]
            assert j.find("parent").get("link") == prev, (side, name)
            prev = j.find("child").get("link")
        assert joints[f"{side.lower()}_foot_sole_joint"].find("parent").get("link") == prev
    return out
