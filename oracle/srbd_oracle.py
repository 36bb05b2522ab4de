"""CPU oracle (numpy) for the batched SRBD-MPC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``biped-pympc_b200``) never imports anything under ``oracle/``.

PARITY PINNED TO THE REFERENCE'S SOURCE, NOT TO CASADI.  The reference defines this path as
two CasADi ``SX`` functions (``qp_former`` and ``sparse_pdipm_multiple_iterations``).
CasADi (unpinned in the reference's ``pyproject.toml:17``; CusADi was built against
3.6.5) is not installed in this image and cannot be installed (no network); the
reference ships no test, golden vector or serialized function for the path (``test/``
is empty).  What pins this file: ``tests/golden/make_golden_ref.py`` imports the
reference's unmodified ``srbd_centroidal_model.py`` / ``srbd_constraints.py`` /
``sparse_pdipm_solver.py`` from ``/root/reference`` and executes them over a numeric
stand-in for the ``casadi`` module (``tests/golden/casadi_shim/casadi.py``); the
committed fixture ``tests/golden/srbd_ref_shim_v1.npz`` holds the patterns and outputs
that code produces, and ``tests/test_reference_pin.py`` checks this restatement against
it (patterns identical, assembly <= 1e-12, iterates <= 1e-9 after the reference's 5
iterations).  What stays unpinned is CasADi's own floating-point operation order
(AMD ordering inside ``ca.ldl``, QR inside ``ca.inv``): the stand-in restates those
calls' published semantics, not their code, so agreement with real CasADi is expected
at the round-off level the reference itself quotes for GPU-vs-CPU ("~1e-10 or
smaller", ``biped_pympc/cusadi/README.md:126``).  ``tests/test_reference_pin.py`` re-runs the
generator over real CasADi whenever ``import casadi`` succeeds (it never does here).

This file is a *restatement* of the published algorithm, written by reading

  biped_pympc/casadi/srbd_centroidal_model.py:101-166   (dynamics, RK4)
  biped_pympc/casadi/srbd_constraints.py:20-227         (qp_former: H,f,A,b,G,d)
  biped_pympc/casadi/sparse_pdipm_solver.py:357-534     (PDIPM, every constant)
  biped_pympc/casadi/sparse_pdipm_solver.py:537-558 and
  biped_pympc/convex_mpc/mpc_controller_cusadi.py:138-141 (cold start, y = 1)
  biped_pympc/convex_mpc/mpc_controller_cusadi.py:144-203 (4x loop, wrench)

It keeps the reference's variable order, CCS (column-major) nonzero order,
constants and operation *semantics* (full 700x700 regularised KKT, LDL^T without
pivoting, two solves sharing one factorisation, ratio tests, clamps).

Everything here is per environment, float64 (or ``np.longdouble`` where noted).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# Problem dimensions (srbd_constraints.py:29-30,97,221-222; script/setup.sh:12)
# ----------------------------------------------------------------------------
N = 10          # horizon
NX = 12         # states  [roll,pitch,yaw, px,py,pz, wx,wy,wz, vx,vy,vz]
NU = 12         # inputs  [f_L(3), f_R(3), m_L(3), m_R(3)]
NZ = (NX + NU) * N          # 240 decision variables  z = [x_1..x_N, u_0..u_{N-1}]
NEQ = NX * N + 2 * N        # 140 = 120 dynamics + 20 "Mx = 0"
NINEQ = 16 * N              # 160
NKKT = NZ + 2 * NINEQ + NEQ  # 700

F_MAX = 500.0   # srbd_constraints.py:31
LT = 0.07       # srbd_constraints.py:161
LH = 0.04       # srbd_constraints.py:162
GRAV = -9.81    # srbd_centroidal_model.py:136

BETA = 1e-8     # generate_solver_function.py:112
DELTA = 1e-8    # sparse_pdipm_solver.py:416

QP_FORMER_NNZ_IN = (12, 120, 120, 120, 1, 1, 1, 9, 9, 3, 3, 3, 20, 12, 12, 3, 3)
QP_FORMER_NAMES_IN = ("x0", "x", "u", "x_ref", "dt", "m", "mu", "R_body", "I_world",
                      "body_pos", "left_foot_pos", "right_foot_pos", "contact_table",
                      "Q", "R", "residual_lin_accel", "residual_ang_accel")


# ----------------------------------------------------------------------------
# Structural sparsity (the "type" of the solver's matrix arguments)
# ----------------------------------------------------------------------------
def _structural_masks():
    """Boolean masks of structural nonzeros of A (140x240), G (160x240), H (240x240).

    Structure = symbolic dependency of CasADi's jacobian (srbd_constraints.py:139,225),
    not numerical nonzero-ness: e.g. R_body entries stay structural when they are 0.
    """
    # continuous dynamics f(x,u) = Ac x + Bc u + c  (srbd_centroidal_model.py:152-166)
    Ac = np.zeros((NX, NX), bool)
    Ac[0:3, 6:9] = True                      # euler_dot = R_body @ omega   (dense 3x3)
    Ac[3:6, 9:12] = np.eye(3, dtype=bool)    # p_dot = v
    Bc = np.zeros((NX, NU), bool)
    Bc[6:9, :] = True                        # I^-1 (skew(r) f + m): dense in all 12 inputs
    Bc[9:12, 0:3] = np.eye(3, dtype=bool)    # f_L / m
    Bc[9:12, 3:6] = np.eye(3, dtype=bool)    # f_R / m
    # RK4 of an affine map with Ac^2 = 0: Ad = I + dt Ac, Bd = dt Bc + dt^2/2 Ac Bc
    Ad = np.eye(NX, dtype=bool) | Ac
    Bd = Bc | ((Ac.astype(int) @ Bc.astype(int)) > 0)

    A = np.zeros((NEQ, NZ), bool)
    for i in range(N):
        r = slice(NX * i, NX * (i + 1))
        A[r, NX * i:NX * (i + 1)] |= np.eye(NX, dtype=bool)          # x_{i+1}
        if i >= 1:
            A[r, NX * (i - 1):NX * i] |= Ad                           # -Ad x_i
        A[r, NX * N + NU * i:NX * N + NU * (i + 1)] |= Bd             # -Bd u_i
        A[NX * N + 2 * i, NX * N + NU * i + 6] = True                 # Mx_left
        A[NX * N + 2 * i + 1, NX * N + NU * i + 9] = True             # Mx_right

    G = np.zeros((NINEQ, NZ), bool)
    for i in range(N):
        u0 = NX * N + NU * i
        for foot, (fx, fy, fz, my) in enumerate(((0, 1, 2, 7), (3, 4, 5, 10))):
            r0 = 16 * i + 8 * foot
            G[r0 + 0, [u0 + fx, u0 + fz]] = True
            G[r0 + 1, [u0 + fx, u0 + fz]] = True
            G[r0 + 2, [u0 + fy, u0 + fz]] = True
            G[r0 + 3, [u0 + fy, u0 + fz]] = True
            G[r0 + 4, [u0 + fz, u0 + my]] = True
            G[r0 + 5, [u0 + fz, u0 + my]] = True
            G[r0 + 6, u0 + fz] = True
            G[r0 + 7, u0 + fz] = True
    H = np.eye(NZ, dtype=bool)
    return A, G, H


def _ccs(mask):
    """CCS (colptr, row) of a boolean mask: column-major, rows ascending within a column
    -- the order of CasADi's ``nonzeros()`` / ``sparsity().get_ccs()``
    (sparse_pdipm_solver.py:561-568)."""
    colptr = [0]
    rows = []
    for j in range(mask.shape[1]):
        r = np.nonzero(mask[:, j])[0]
        rows.extend(r.tolist())
        colptr.append(len(rows))
    return np.asarray(colptr, np.int32), np.asarray(rows, np.int32)


A_MASK, G_MASK, H_MASK = _structural_masks()
A_COLPTR, A_ROW = _ccs(A_MASK)
G_COLPTR, G_ROW = _ccs(G_MASK)
H_COLPTR, H_ROW = _ccs(H_MASK)
NNZ_A = int(A_COLPTR[-1])   # 1196
NNZ_G = int(G_COLPTR[-1])   # 280
NNZ_H = int(H_COLPTR[-1])   # 240
A_COL = np.repeat(np.arange(NZ, dtype=np.int32), np.diff(A_COLPTR))
G_COL = np.repeat(np.arange(NZ, dtype=np.int32), np.diff(G_COLPTR))
assert (NNZ_A, NNZ_G, NNZ_H) == (1196, 280, 240)

QP_FORMER_NNZ_OUT = (NNZ_H, NZ, NNZ_A, NEQ, NNZ_G, NINEQ)        # 240,240,1196,140,280,160
PDIPM_NNZ_IN = (NNZ_H, NNZ_G, NNZ_A, NZ, NINEQ, NEQ, NZ, NINEQ, NINEQ, NEQ)
PDIPM_NNZ_OUT = (NZ, NINEQ, NINEQ, NEQ, 4, 1)


def dense_from_ccs(val, colptr, row, shape, dtype=np.float64):
    M = np.zeros(shape, dtype)
    col = np.repeat(np.arange(shape[1]), np.diff(colptr))
    M[row, col] = val
    return M


# ----------------------------------------------------------------------------
# a1: continuous dynamics + RK4 (srbd_centroidal_model.py:101-166)
# ----------------------------------------------------------------------------
def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]], dtype=v.dtype)


def continuous_affine(R_body, I_world, body_pos, left_foot_pos, right_foot_pos, m,
                      residual_lin_accel, residual_ang_accel, dtype=np.float64):
    """f(x,u) = Ac x + Bc u + c   (srbd_centroidal_model.py:123-166).

    ``R_body`` and ``I_world`` are 3x3 matrices already un-flattened from CasADi's
    column-major nnz order.  ``I_world`` is the *world inertia* (the reference's
    symbol is misleadingly called 'I_world_inv', srbd_constraints.py:40); it is
    inverted here as ``casadi.inv`` does at srbd_centroidal_model.py:152."""
    R_body = np.asarray(R_body, dtype)
    I_inv = np.linalg.inv(np.asarray(I_world, np.float64)).astype(dtype) if dtype == np.float64 \
        else _inv3(np.asarray(I_world, dtype))
    rl = np.asarray(left_foot_pos, dtype) - np.asarray(body_pos, dtype)
    rr = np.asarray(right_foot_pos, dtype) - np.asarray(body_pos, dtype)
    Ac = np.zeros((NX, NX), dtype)
    Ac[0:3, 6:9] = R_body
    Ac[3:6, 9:12] = np.eye(3, dtype=dtype)
    Bc = np.zeros((NX, NU), dtype)
    Bc[6:9, 0:3] = I_inv @ _skew(rl)
    Bc[6:9, 3:6] = I_inv @ _skew(rr)
    Bc[6:9, 6:9] = I_inv
    Bc[6:9, 9:12] = I_inv
    Bc[9:12, 0:3] = np.eye(3, dtype=dtype) / dtype(m)
    Bc[9:12, 3:6] = np.eye(3, dtype=dtype) / dtype(m)
    c = np.zeros(NX, dtype)
    c[6:9] = np.asarray(residual_ang_accel, dtype)
    c[9:12] = np.array([0.0, 0.0, GRAV], dtype) + np.asarray(residual_lin_accel, dtype)
    return Ac, Bc, c


def _inv3(M):
    """3x3 inverse by adjugate, in the dtype of M (used for the long-double oracle)."""
    a, b, c = M[0]; d, e, f = M[1]; g, h, i = M[2]
    A = e * i - f * h; B = -(d * i - f * g); C = d * h - e * g
    det = a * A + b * B + c * C
    adj = np.array([[A, -(b * i - c * h), b * f - c * e],
                    [B, a * i - c * g, -(a * f - c * d)],
                    [C, -(a * h - b * g), a * e - b * d]], dtype=M.dtype)
    return adj / det


def rk4_affine(Ac, Bc, c, dt):
    """Literal RK4 (srbd_centroidal_model.py:101-121) of x' = Ac x + Bc u + c carried
    out on the affine-map coefficients, i.e. what differentiating the RK4 expression
    gives.  Returns (Ad, Bd, cd) with x+ = Ad x + Bd u + cd."""
    dtype = Ac.dtype
    I = np.eye(NX, dtype=dtype)
    dt = dtype.type(dt)
    half = dt / dtype.type(2)
    # k_j = Kx_j x + Ku_j u + kc_j
    Kx1, Ku1, kc1 = Ac, Bc, c
    Kx2, Ku2, kc2 = Ac @ (I + half * Kx1), Ac @ (half * Ku1) + Bc, Ac @ (half * kc1) + c
    Kx3, Ku3, kc3 = Ac @ (I + half * Kx2), Ac @ (half * Ku2) + Bc, Ac @ (half * kc2) + c
    Kx4, Ku4, kc4 = Ac @ (I + dt * Kx3), Ac @ (dt * Ku3) + Bc, Ac @ (dt * kc3) + c
    s = dt / dtype.type(6)
    two = dtype.type(2)
    Ad = I + s * (Kx1 + two * Kx2 + two * Kx3 + Kx4)
    Bd = s * (Ku1 + two * Ku2 + two * Ku3 + Ku4)
    cd = s * (kc1 + two * kc2 + two * kc3 + kc4)
    return Ad, Bd, cd


def rk4_step(Ac, Bc, c, x, u, dt):
    """Literal numerical RK4 step (srbd_centroidal_model.py:110-115)."""
    f = lambda xx: Ac @ xx + Bc @ u + c
    k1 = f(x)
    k2 = f(x + dt / 2 * k1)
    k3 = f(x + dt / 2 * k2)
    k4 = f(x + dt * k3)
    return x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def discrete_closed_form(Ac, Bc, c, dt):
    """Closed form used by the CUDA kernel: Ac is nilpotent (Ac^2 = 0) so RK4 is exact:
    Ad = I + dt Ac, Bd = dt Bc + dt^2/2 Ac Bc, cd = dt c + dt^2/2 Ac c  (SURVEY F5)."""
    I = np.eye(NX, dtype=Ac.dtype)
    return I + dt * Ac, dt * Bc + (dt * dt / 2) * (Ac @ Bc), dt * c + (dt * dt / 2) * (Ac @ c)


# ----------------------------------------------------------------------------
# a2-a4: qp_former (srbd_constraints.py:20-227)
# ----------------------------------------------------------------------------
def qp_former(x0, x, u, x_ref, dt, m, mu, R_body, I_world, body_pos, left_foot_pos,
              right_foot_pos, contact_table, Q, R, residual_lin_accel, residual_ang_accel,
              dtype=np.float64):
    """One environment.  All arguments are flat vectors in CasADi nnz order (column-major
    for the 3x3 matrices and the 10x2 contact table), exactly the rows of the (B, nnz)
    tensors the reference passes at mpc_controller_cusadi.py:77-95.

    Returns (H_val[240], f[240], A_val[1196], b[140], G_val[280], d[160]): the six
    outputs of ``Function 'qp_former'`` (srbd_constraints.py:75-79) as CCS value vectors.
    """
    t = dtype
    x0 = np.asarray(x0, t).reshape(NX)
    x = np.asarray(x, t).reshape(NX * N)
    u = np.asarray(u, t).reshape(NU * N)
    x_ref = np.asarray(x_ref, t).reshape(NX * N)
    dt = t(np.asarray(dt).reshape(()))
    m = t(np.asarray(m).reshape(()))
    mu = t(np.asarray(mu).reshape(()))
    Rb = np.asarray(R_body, t).reshape(3, 3, order="F")       # column-major nnz order
    Iw = np.asarray(I_world, t).reshape(3, 3, order="F")
    contact = np.asarray(contact_table, t).reshape(N, 2, order="F")  # flat[k] = c[k%10, k//10]
    Q = np.asarray(Q, t).reshape(NX)
    R = np.asarray(R, t).reshape(NU)
    z = np.concatenate([x, u])

    # --- cost: H = hessian, f = grad - H z   (srbd_constraints.py:65-69)
    Hdiag = np.concatenate([np.tile(Q, N), np.tile(R, N)])
    grad = np.concatenate([np.tile(Q, N) * (x - x_ref), np.tile(R, N) * u])
    f = grad - Hdiag * z

    # --- equalities (srbd_constraints.py:83-142)
    Ac, Bc, c = continuous_affine(Rb, Iw, body_pos, left_foot_pos, right_foot_pos, m,
                                  residual_lin_accel, residual_ang_accel, dtype=t)
    Ad, Bd, cd = rk4_affine(Ac, Bc, c, dt)
    A = np.zeros((NEQ, NZ), t)
    eq = np.zeros(NEQ, t)
    x_cat = np.concatenate([x0, x])
    for i in range(N):
        r = slice(NX * i, NX * (i + 1))
        A[r, NX * i:NX * (i + 1)] += np.eye(NX, dtype=t)
        if i >= 1:
            A[r, NX * (i - 1):NX * i] -= Ad
        A[r, NX * N + NU * i:NX * N + NU * (i + 1)] -= Bd
        x_i = x_cat[NX * i:NX * (i + 1)]
        u_i = u[NU * i:NU * (i + 1)]
        x_next = x_cat[NX * (i + 1):NX * (i + 2)]
        eq[r] = x_next - rk4_step(Ac, Bc, c, x_i, u_i, dt)
        A[NX * N + 2 * i, NX * N + NU * i + 6] = 1
        A[NX * N + 2 * i + 1, NX * N + NU * i + 9] = 1
        eq[NX * N + 2 * i] = u_i[6]
        eq[NX * N + 2 * i + 1] = u_i[9]
    b = A @ z - eq                                            # srbd_constraints.py:140

    # --- inequalities (srbd_constraints.py:144-227)
    G = np.zeros((NINEQ, NZ), t)
    ineq = np.zeros(NINEQ, t)
    lt, lh, fmax = t(LT), t(LH), t(F_MAX)
    for i in range(N):
        u0 = NX * N + NU * i
        u_i = u[NU * i:NU * (i + 1)]
        for foot, (fx, fy, fz, my) in enumerate(((0, 1, 2, 7), (3, 4, 5, 10))):
            r0 = 16 * i + 8 * foot
            rows = (
                ((fx, -1), (fz, -mu)), ((fx, 1), (fz, -mu)),
                ((fy, -1), (fz, -mu)), ((fy, 1), (fz, -mu)),
                ((fz, -lt), (my, -1)), ((fz, -lh), (my, 1)),
                ((fz, -1),), ((fz, 1),),
            )
            for j, terms in enumerate(rows):
                val = t(0)
                for col, coef in terms:
                    G[r0 + j, u0 + col] = coef
                    val = val + t(coef) * u_i[col]
                if j == 7:
                    val = val - fmax * contact[i, foot]
                ineq[r0 + j] = val
    d = G @ z - ineq                                          # srbd_constraints.py:226

    H_val = Hdiag.copy()
    A_val = A[A_ROW, A_COL]
    G_val = G[G_ROW, G_COL]
    return H_val, f, A_val, b, G_val, d


# ----------------------------------------------------------------------------
# a6: sparse_pdipm_multiple_iterations (sparse_pdipm_solver.py:357-534)
# ----------------------------------------------------------------------------
def ldl_nopivot(K):
    """Dense LDL^T without pivoting (the semantics of ``ca.ldl``: no numerical pivoting,
    sparse_pdipm_solver.py:451).  The KKT matrix is symmetric quasi-definite so the
    factorisation exists for every symmetric permutation.  Returns (L unit-lower, D)."""
    K = K.copy()
    n = K.shape[0]
    D = np.zeros(n, K.dtype)
    for j in range(n):
        dj = K[j, j]
        D[j] = dj
        col = K[j + 1:, j] / dj
        K[j + 1:, j + 1:] -= np.outer(col, K[j + 1:, j])
        K[j + 1:, j] = col
    L = np.tril(K, -1) + np.eye(n, dtype=K.dtype)
    return L, D


def ldl_solve(L, D, rhs):
    import scipy.linalg as sla
    if L.dtype == np.float64:
        y = sla.solve_triangular(L, rhs, lower=True, unit_diagonal=True)
        y = y / D
        return sla.solve_triangular(L.T, y, lower=False, unit_diagonal=True)
    # generic dtype (long double): plain substitution
    n = L.shape[0]
    y = rhs.astype(L.dtype).copy()
    for i in range(n):
        y[i] -= L[i, :i] @ y[:i]
    y /= D
    for i in range(n - 1, -1, -1):
        y[i] -= L[i + 1:, i] @ y[i + 1:]
    return y


def _ratio(v, dv):
    """alpha = max(min(1, 0.99 * min_i{-v_i/dv_i if dv_i < 0 else 1}), 1e-12)
    (sparse_pdipm_solver.py:460-467, 495-502)."""
    t = v.dtype.type
    with np.errstate(divide="ignore", invalid="ignore"):
        cand = np.where(dv < 0, -v / dv, t(1.0))
    a = min(t(1.0), t(0.99) * cand.min())
    return max(a, t(1e-12))


def build_kkt(Hm, Gm, Am, s, z, beta=BETA, delta=DELTA):
    """The regularised 4-block KKT in variable order [dx, ds, dz, dy]
    (sparse_pdipm_solver.py:406-439)."""
    t = Hm.dtype
    K = np.zeros((NKKT, NKKT), t)
    o1, o2, o3 = NZ, NZ + NINEQ, NZ + 2 * NINEQ
    K[:o1, :o1] = Hm + t.type(beta) * np.eye(NZ, dtype=t)
    K[:o1, o2:o3] = Gm.T
    K[:o1, o3:] = Am.T
    K[o1:o2, o1:o2] = np.diag((t.type(1.0) / s) * z + t.type(delta))
    K[o1:o2, o2:o3] = np.eye(NINEQ, dtype=t)
    K[o2:o3, :o1] = Gm
    K[o2:o3, o1:o2] = np.eye(NINEQ, dtype=t)
    K[o2:o3, o2:o3] = -t.type(delta) * np.eye(NINEQ, dtype=t)
    K[o3:, :o1] = Am
    K[o3:, o3:] = -t.type(delta) * np.eye(NEQ, dtype=t)
    return K


def pdipm_iterations(Q_val, G_val, A_val, f, h, b, x, s, z, y, num_iters=5,
                     beta=BETA, delta=DELTA, dtype=np.float64, linsolve="ldl", trace=None):
    """``sparse_pdipm_multiple_iterations`` for one environment.

    Inputs are the ten value vectors of sparse_pdipm_solver.py:532-534 (CCS values
    against the frozen patterns A_*/G_*/H_* above).  Returns
    (x, s, z, y, residuals[4] = (|rx|, |rs|, |re|, mu_new), mu_new).
    ``linsolve``: "ldl" = dense LDL^T without pivoting (reference semantics),
    "lu" = numpy's pivoted LU (cross-check only)."""
    t = np.dtype(dtype)
    one = t.type(1.0)
    Hm = dense_from_ccs(np.asarray(Q_val, t), H_COLPTR, H_ROW, (NZ, NZ), t)
    Gm = dense_from_ccs(np.asarray(G_val, t), G_COLPTR, G_ROW, (NINEQ, NZ), t)
    Am = dense_from_ccs(np.asarray(A_val, t), A_COLPTR, A_ROW, (NEQ, NZ), t)
    f = np.asarray(f, t).reshape(NZ); h = np.asarray(h, t).reshape(NINEQ)
    b = np.asarray(b, t).reshape(NEQ)
    x = np.asarray(x, t).reshape(NZ).copy(); s = np.asarray(s, t).reshape(NINEQ).copy()
    z = np.asarray(z, t).reshape(NINEQ).copy(); y = np.asarray(y, t).reshape(NEQ).copy()
    o1, o2, o3 = NZ, NZ + NINEQ, NZ + 2 * NINEQ
    rx = rs = re = None
    mu_new = None
    for it in range(num_iters):
        rx = Hm @ x + f + Gm.T @ z + Am.T @ y                 # :392-399
        re = Am @ x - b
        rs = Gm @ x + s - h                                   # :401
        mu = (s @ z) / t.type(NINEQ)                          # :402
        K = build_kkt(Hm, Gm, Am, s, z, beta, delta)
        s_inv = one / s
        rhs = np.concatenate([-rx, -(s_inv * (s * z)), -rs, -re])   # :443-448
        if linsolve == "ldl":
            L, D = ldl_nopivot(K)
            solve = lambda r: ldl_solve(L, D, r)
        else:
            import scipy.linalg as sla
            lu = sla.lu_factor(K.astype(np.float64))
            solve = lambda r: sla.lu_solve(lu, r.astype(np.float64)).astype(t)
        aff = solve(rhs)
        dx_a, ds_a, dz_a, dy_a = aff[:o1], aff[o1:o2], aff[o2:o3], aff[o3:]
        a_pri = _ratio(s, ds_a)                               # :460-467
        a_dual = _ratio(z, dz_a)
        s_aff = s + a_pri * ds_a
        z_aff = z + a_dual * dz_a
        mu_aff = (s_aff @ z_aff) / t.type(NINEQ)
        sigma = (mu_aff / mu) ** 3                            # :474
        rc = s * z + ds_a * dz_a - sigma * mu * np.ones(NINEQ, t)   # :478
        rhs_c = np.concatenate([np.zeros(NZ, t), -(s_inv * rc), np.zeros(NINEQ, t), np.zeros(NEQ, t)])
        cor = solve(rhs_c)                                    # :482
        dx = dx_a + cor[:o1]; ds = ds_a + cor[o1:o2]
        dz = dz_a + cor[o2:o3]; dy = dy_a + cor[o3:]
        alpha_pri = _ratio(s, ds)                             # :495-502
        alpha_dual = _ratio(z, dz)
        if trace is not None:
            trace.append(dict(rx=rx.copy(), rs=rs.copy(), re=re.copy(), mu=mu, dx_aff=dx_a.copy(),
                              ds_aff=ds_a.copy(), dz_aff=dz_a.copy(), dy_aff=dy_a.copy(),
                              a_pri_aff=a_pri, a_dual_aff=a_dual, sigma=sigma, dx=dx.copy(),
                              ds=ds.copy(), dz=dz.copy(), dy=dy.copy(),
                              alpha_pri=alpha_pri, alpha_dual=alpha_dual))
        x = x + alpha_pri * dx                                # :505-508
        s = s + alpha_pri * ds
        z = z + alpha_dual * dz
        y = y + alpha_dual * dy
        s = np.maximum(s, t.type(1e-8))                       # :511-515
        z = np.maximum(z, t.type(1e-8))
        mu_new = (s @ z) / t.type(NINEQ)                      # :524
    res = np.array([np.sqrt(rx @ rx), np.sqrt(rs @ rs), np.sqrt(re @ re), mu_new], t)  # :527-530
    return x, s, z, y, res, np.array([mu_new], t)


# ----------------------------------------------------------------------------
# a7: cold start as used on the GPU path (mpc_controller_cusadi.py:138-141)
# ----------------------------------------------------------------------------
def cold_start(G_val, d, dtype=np.float64):
    """x = 0, s = max(d - G x, 1), z = 1, y = 1  (note y = 1, not 0)."""
    t = dtype
    x = np.zeros(NZ, t)
    Gm = dense_from_ccs(np.asarray(G_val, t), G_COLPTR, G_ROW, (NINEQ, NZ), t)
    s = np.maximum(np.asarray(d, t).reshape(NINEQ) - Gm @ x, t(1.0))
    return x, s, np.ones(NINEQ, t), np.ones(NEQ, t)


# ----------------------------------------------------------------------------
# a8: the reference tick: form -> cold start -> calls x iters -> u0 -> body wrench
# ----------------------------------------------------------------------------
def mpc_tick(inputs17, rotation_body, num_calls=4, iters_per_call=5, dtype=np.float64):
    """mpc_controller_cusadi.py:97-203 for one environment.

    ``rotation_body`` is the 3x3 row-major state-estimate rotation used for the final
    world->body transform (mpc_controller_cusadi.py:197-200); it is the same matrix
    whose row-major flattening the reference feeds to ``R_body`` (quirk Q1)."""
    Hv, f, Av, b, Gv, d = qp_former(*inputs17, dtype=dtype)
    x, s, z, y = cold_start(Gv, d, dtype=dtype)
    res = None
    for _ in range(num_calls):
        x, s, z, y, res, _mu = pdipm_iterations(Hv, Gv, Av, f, d, b, x, s, z, y,
                                                num_iters=iters_per_call, dtype=dtype)
    u0 = x[NX * N:NX * N + NU].astype(np.float64)
    Rb = np.asarray(rotation_body, np.float64).reshape(3, 3)
    # the reference casts to fp32 *before* the rotation (mpc_controller_cusadi.py:189-200)
    fl, fr = u0[0:3].astype(np.float32), u0[3:6].astype(np.float32)
    ml, mr = u0[6:9].astype(np.float32), u0[9:12].astype(np.float32)
    ml[0] = 0.0
    mr[0] = 0.0
    Rt = Rb.astype(np.float32).T
    wrench = np.concatenate([-(Rt @ fl), -(Rt @ ml), -(Rt @ fr), -(Rt @ mr)]).reshape(2, 6)
    return wrench.astype(np.float32), (x, s, z, y, res)
