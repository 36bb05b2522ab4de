"""CPU tests of the oracle itself (no GPU, no product code on the compute path)."""
import numpy as np
import pytest

import biped_pympc_b200 as bp
from helpers import oracle_ip_batch, oracle_qp_batch, rel_err
from oracle import srbd_oracle as so
from oracle import structured_proto as sp


def test_dimensions_match_reference_filename():
    # mpc_multiple_iter_5_solver_240v_140eq_160ineq.casadi (script/setup.sh:12)
    assert (so.NZ, so.NEQ, so.NINEQ, so.NKKT) == (240, 140, 160, 700)
    assert (so.NNZ_H, so.NNZ_A, so.NNZ_G) == (240, 1196, 280)
    assert sum(so.QP_FORMER_NNZ_IN) == 452 and sum(so.QP_FORMER_NNZ_OUT) == 2256
    assert sum(so.PDIPM_NNZ_IN) == 2956 and sum(so.PDIPM_NNZ_OUT) == 705


def test_rk4_is_exact_closed_form():
    inp, _ = bp.workload.make_batch(8, seed=3)
    for e in range(8):
        a = [t[e] for t in inp]
        Rm, Iw = a[7].reshape(3, 3, order="F"), a[8].reshape(3, 3, order="F")
        Ac, Bc, c = so.continuous_affine(Rm, Iw, a[9], a[10], a[11], a[5][0], a[15], a[16])
        assert np.abs(Ac @ Ac).max() == 0.0                        # nilpotent
        lit = so.rk4_affine(Ac, Bc, c, a[4][0])
        cf = so.discrete_closed_form(Ac, Bc, c, a[4][0])
        for l, f in zip(lit, cf):
            assert np.abs(l - f).max() <= 4e-16 * max(1.0, np.abs(l).max())
        # literal numeric RK4 step agrees with the affine-map form
        x = np.random.default_rng(e).normal(size=12); u = np.random.default_rng(e + 100).normal(size=12)
        assert np.abs(so.rk4_step(Ac, Bc, c, x, u, a[4][0]) - (lit[0] @ x + lit[1] @ u + lit[2])).max() < 1e-13


def test_qp_former_nominal_standing_fixture():
    """The reference's own fixture (srbd_constraints.py:243-288): analytic known answers."""
    a = [t[0] for t in bp.workload.nominal_standing_inputs()]
    Hv, f, Av, b, Gv, d = so.qp_former(*a)
    Q, R = a[13], a[14]
    assert np.array_equal(Hv, np.concatenate([np.tile(Q, 10), np.tile(R, 10)]))
    x_ref = a[3]
    assert np.abs(f[:120] + np.tile(Q, 10) * x_ref).max() < 1e-10 and np.abs(f[120:]).max() < 1e-12
    dd = d.reshape(10, 16)
    assert np.allclose(dd[:, [7, 15]], 500.0, atol=1e-9) and np.abs(np.delete(dd, [7, 15], axis=1)).max() < 1e-12
    A = so.dense_from_ccs(Av, so.A_COLPTR, so.A_ROW, (140, 240))
    G = so.dense_from_ccs(Gv, so.G_COLPTR, so.G_ROW, (160, 240))
    assert np.count_nonzero(A[:120, :120].diagonal() == 1.0) == 120
    # mass * gravity: with x0 at rest and u = m g / 2 per foot, the dynamics rows balance in z
    m, dt = a[5][0], a[4][0]
    Bd_vz = -A[11, 120 + 2]                                    # d v_z / d f_Lz = dt / m
    assert abs(Bd_vz - dt / m) < 1e-15
    assert abs(b[11] - dt * so.GRAV) < 1e-12                   # c_d of v_z (x0 v = 0)
    mu = a[6][0]
    assert np.allclose(G[0, [120, 122]], [-1, -mu]) and np.allclose(G[4, [122, 127]], [-so.LT, -1])


def test_contact_and_rotation_layout_quirks():
    """Kernel contract = CasADi column-major nnz order."""
    inp, _ = bp.workload.make_batch(1, seed=11)
    a = [t[0].copy() for t in inp]
    ct = np.zeros(20); ct[3] = 1.0; ct[10 + 6] = 1.0          # stage 3 left, stage 6 right
    a[12] = ct
    d = so.qp_former(*a)[5].reshape(10, 16)
    assert d[3, 7] == 500.0 and d[6, 15] == 500.0 and d.sum() == 1000.0


def test_direction_satisfies_reference_kkt():
    """The structured solve reproduces the reference's regularised 700x700 KKT system."""
    inp, _ = bp.workload.make_batch(3, seed=5)
    qp = oracle_qp_batch(inp)
    Hv, f, Av, b, Gv, d = [q[1] for q in qp]
    x, s, z, y = so.cold_start(Gv, d)
    x, s, z, y, _, _ = so.pdipm_iterations(Hv, Gv, Av, f, d, b, x, s, z, y, num_iters=2, linsolve="lu")
    Hm = so.dense_from_ccs(Hv, so.H_COLPTR, so.H_ROW, (240, 240))
    Gm = so.dense_from_ccs(Gv, so.G_COLPTR, so.G_ROW, (160, 240))
    Am = so.dense_from_ccs(Av, so.A_COLPTR, so.A_ROW, (140, 240))
    K = so.build_kkt(Hm, Gm, Am, s, z)
    rng = np.random.default_rng(0)
    rhs = rng.normal(size=700)
    st = sp.StructuredStep(Hv, Gv, Av, s, z)
    sol = np.concatenate(st.solve(rhs[:240], rhs[240:400], rhs[400:560], rhs[560:]))
    ref = np.linalg.solve(K, rhs)
    assert rel_err(sol, ref) < 1e-8
    assert np.linalg.norm(K @ sol - rhs) / np.linalg.norm(rhs) < 1e-6


@pytest.mark.parametrize("iters", [1, 5])
def test_structured_matches_dense_oracle(iters):
    inp, _ = bp.workload.make_batch(3, seed=21)
    qp = oracle_qp_batch(inp)
    ref = oracle_ip_batch(qp, iters, dtype=np.longdouble)
    got = oracle_ip_batch(qp, iters, structured=True)
    for i in range(4):
        assert rel_err(got[i], ref[i]) < 1e-9


def test_oracle_against_golden(golden):
    names = bp.workload.INPUT_NAMES
    inputs = [golden[f"in_{n}"] for n in names]
    qp = oracle_qp_batch(inputs)
    for n, v in zip(("H", "f", "A", "b", "G", "d"), qp):
        assert np.array_equal(v, golden[f"qp_{n}"])
    got = oracle_ip_batch([g[:2] for g in qp], 5)
    for i, n in enumerate(("x", "s", "z", "y")):
        assert np.array_equal(got[i], golden[f"ip5f_{n}"][:2])             # float64 oracle is deterministic
        assert rel_err(got[i], golden[f"ip5_{n}"][:2]) < 5e-9              # vs long double evaluation


def test_standing_solution_approaches_weight_support(golden):
    """Physics known-answer: after 20 iterations on the reference's nominal standing pose
    (at rest, at the reference height), fz_L + fz_R ~ m g and the feet share it equally."""
    x = golden["ip20_x"][0]                  # env 0 = nominal standing fixture
    m = golden["in_m"][0, 0]
    assert abs(x[120 + 2] - x[120 + 5]) < 1e-6 * m * 9.81
    fz = x[120 + 2] + x[120 + 5]
    assert abs(fz - m * 9.81) / (m * 9.81) < 0.05
