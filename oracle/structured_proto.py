"""numpy prototype of the *structured* Newton step the CUDA kernel implements.
TEST INFRASTRUCTURE ONLY (same rules as srbd_oracle.py: never imported by the product).

It exists to check, on the CPU, that the exact block elimination used on the GPU
(DESIGN.md "PDIPM kernel") reproduces the reference's full 700x700 regularised KKT
solve (sparse_pdipm_solver.py:406-452) including the beta/delta terms:

  W = z/s + delta ; D = 1/(1/W + delta)
  Phi = H + beta I + G' D G        (diagonal on x; per stage two "arrow" 4x4 blocks on
                                    (fx,fy,my | fz) plus diagonal mx,mz)
  (A Phi^-1 A' + delta I) dy = A Phi^-1 r1~ - r4      block tridiagonal, 12x12 blocks,
                                    Mx rows eliminated in closed form
  block-Thomas with explicit SPD pivot inverses, "twisted" (both ends -> middle)

The kernel-side data flow (per stage quantities, twisted order) is mirrored so that
intermediate values can be compared when debugging the kernel.
"""
from __future__ import annotations

import numpy as np

from . import srbd_oracle as so

N, NX, NU, NZ, NEQ, NINEQ = so.N, so.NX, so.NU, so.NZ, so.NEQ, so.NINEQ


# -- CCS offset helpers (must match csrc/srbd_pattern.h) ---------------------------------
def a_xcol_off(c):
    return 2 * c if c < 6 else (12 + 5 * (c - 6) if c < 9 else 27 + 3 * (c - 9))


def a_xcol_start(k, c):
    return 36 * k + a_xcol_off(c) if k < 9 else 324 + c


_UCOL_OFF = (0, 8, 16, 24, 32, 40, 48, 55, 61, 67, 74, 80)


def a_ucol_start(k, c):
    return 336 + 86 * k + _UCOL_OFF[c]


def a_ucol_pos(r, c):
    if r < 3:
        return r
    if 6 <= r < 9:
        return r - 2 if c < 6 else r - 3
    if c >= 6:
        return -1
    if r == 3 + c % 3:
        return 3
    if r == 9 + c % 3:
        return 7
    return -1


def unpack_stage_matrices(A_val):
    """E[k][12], Abar[k][12,12] (k>=1), Bbar[k][12,12], mx[k][2] from the CCS values."""
    t = A_val.dtype
    E = np.zeros((N, NX), t)
    Ab = np.zeros((N, NX, NX), t)
    Bb = np.zeros((N, NX, NU), t)
    mx = np.zeros((N, 2), t)
    for k in range(N):
        for c in range(NX):
            st = a_xcol_start(k, c)
            E[k, c] = A_val[st]
            if k < 9:
                # entries of -Ad column c live in row block k+1 -> Abar[k+1][:, c]
                if c < 6:
                    Ab[k + 1, c, c] = A_val[st + 1]
                elif c < 9:
                    Ab[k + 1, 0:3, c] = A_val[st + 1:st + 4]
                    Ab[k + 1, c, c] = A_val[st + 4]
                else:
                    Ab[k + 1, c - 6, c] = A_val[st + 1]
                    Ab[k + 1, c, c] = A_val[st + 2]
        for c in range(NU):
            st = a_ucol_start(k, c)
            for r in range(NX):
                p = a_ucol_pos(r, c)
                if p >= 0:
                    Bb[k, r, c] = A_val[st + p]
            if c == 6:
                mx[k, 0] = A_val[st + 6]
            if c == 9:
                mx[k, 1] = A_val[st + 6]
    return E, Ab, Bb, mx


def unpack_G(G_val):
    """Per stage k, foot ft, row j: (coef on fx|fy|fz(first col), coef on fz|my(second))."""
    t = G_val.dtype
    Gs = np.zeros((N, 16, NU), t)
    for k in range(N):
        for ft, (fx, fy, fz, my) in enumerate(((0, 1, 2, 7), (3, 4, 5, 10))):
            gb = 28 * k + 12 * ft
            mb = 28 * k + 24 + 2 * ft
            r0 = 8 * ft
            Gs[k, r0 + 0, fx] = G_val[gb + 0]; Gs[k, r0 + 0, fz] = G_val[gb + 4]
            Gs[k, r0 + 1, fx] = G_val[gb + 1]; Gs[k, r0 + 1, fz] = G_val[gb + 5]
            Gs[k, r0 + 2, fy] = G_val[gb + 2]; Gs[k, r0 + 2, fz] = G_val[gb + 6]
            Gs[k, r0 + 3, fy] = G_val[gb + 3]; Gs[k, r0 + 3, fz] = G_val[gb + 7]
            Gs[k, r0 + 4, fz] = G_val[gb + 8]; Gs[k, r0 + 4, my] = G_val[mb + 0]
            Gs[k, r0 + 5, fz] = G_val[gb + 9]; Gs[k, r0 + 5, my] = G_val[mb + 1]
            Gs[k, r0 + 6, fz] = G_val[gb + 10]
            Gs[k, r0 + 7, fz] = G_val[gb + 11]
    return Gs


def spd_inverse_gauss_jordan(M):
    """SPD inverse by the symmetric sweep operator, as each half-warp does it: every lane
    keeps one full row, the (unscaled) pivot row is broadcast, all pivots 0..11 are swept
    (giving -M^-1), and only the lower triangle is kept.  A plain Gauss-Jordan inverse is
    not symmetric in floating point and loses ~1.5 digits once its upper half is dropped."""
    A = M.copy()
    n = A.shape[0]
    for p in range(n):
        v = A[p, :].copy()
        di = 1.0 / v[p]
        for r in range(n):
            if r == p:
                A[p, :] = v * di
                A[p, p] = -di
            else:
                t = v[r] * di
                row = A[r, :] - t * v
                row[p] = t
                A[r, :] = row
    X = -A
    return np.tril(X) + np.tril(X, -1).T


class StructuredStep:
    """Factorisation of one Newton step's KKT in reduced form + solve()."""

    def __init__(self, Q_val, G_val, A_val, s, z, beta=so.BETA, delta=so.DELTA, twisted=True):
        t = np.float64
        self.delta = delta
        self.E, self.Ab, self.Bb, self.mx = unpack_stage_matrices(np.asarray(A_val, t))
        self.Gs = unpack_G(np.asarray(G_val, t))
        hx = np.asarray(Q_val[:120], t).reshape(N, NX)
        hu = np.asarray(Q_val[120:], t).reshape(N, NU)
        self.phix_inv = 1.0 / (hx + beta)            # for x_{k+1}
        s = np.asarray(s, t).reshape(N, 16)
        z = np.asarray(z, t).reshape(N, 16)
        self.W = z / s + delta
        self.Wi = 1.0 / self.W
        self.D = 1.0 / (self.Wi + delta)
        # ---- Psi_k = Phi_u^-1 in "diag + rank-1 per foot" form (cancellation-free Schur)
        self.psi_diag = np.zeros((N, NU), t)          # 1/a_i for fx,fy,my ; 0 for fz ; mz ; mx'
        self.tq = np.zeros((N, 2, NU), t)             # q vector per foot (t_i on fx,fy,my ; -1 on fz)
        self.sig_inv = np.zeros((N, 2), t)
        self.mx_den = np.zeros((N, 2), t)             # m^2 + delta*phi
        for k in range(N):
            eps = hu[k] + beta
            for ft, (fx, fy, fz, my) in enumerate(((0, 1, 2, 7), (3, 4, 5, 10))):
                Dk = self.D[k, 8 * ft:8 * ft + 8]
                Gk = self.Gs[k, 8 * ft:8 * ft + 8]
                sigma = eps[fz] + Gk[6, fz] ** 2 * Dk[6] + Gk[7, fz] ** 2 * Dk[7]
                q = np.zeros(NU, t)
                q[fz] = -1.0
                for (col, j0) in ((fx, 0), (fy, 2), (my, 4)):
                    j1 = j0 + 1
                    gx0, gx1 = Gk[j0, col], Gk[j1, col]
                    gz0, gz1 = Gk[j0, fz], Gk[j1, fz]
                    D0, D1 = Dk[j0], Dk[j1]
                    a = eps[col] + gx0 * gx0 * D0 + gx1 * gx1 * D1
                    bb = gx0 * gz0 * D0 + gx1 * gz1 * D1
                    cross = gz0 * gx1 - gz1 * gx0
                    sigma += (eps[col] * (gz0 * gz0 * D0 + gz1 * gz1 * D1) + D0 * D1 * cross * cross) / a
                    self.psi_diag[k, col] = 1.0 / a
                    q[col] = bb / a
                self.tq[k, ft] = q
                self.sig_inv[k, ft] = 1.0 / sigma
            # mz: plain diagonal ; mx: eliminated together with its equality row
            self.psi_diag[k, 8] = 1.0 / eps[8]
            self.psi_diag[k, 11] = 1.0 / eps[11]
            for i, c in enumerate((6, 9)):
                den = self.mx[k, i] ** 2 + delta * eps[c]
                self.mx_den[k, i] = den
                self.psi_diag[k, c] = delta / den       # psi'
            self.eps_u = getattr(self, "eps_u", np.zeros((N, NU), t))
            self.eps_u[k] = eps
        # ---- Schur blocks
        self.S = np.zeros((N, NX, NX), t)
        self.C = np.zeros((N, NX, NX), t)             # C[k] = S_{k,k-1}  (k>=1)
        for k in range(N):
            Sk = np.diag(self.E[k] ** 2 * self.phix_inv[k]) + delta * np.eye(NX)
            if k >= 1:
                Sk += (self.Ab[k] * self.phix_inv[k - 1][None, :]) @ self.Ab[k].T
                self.C[k] = self.Ab[k] * (self.phix_inv[k - 1] * self.E[k - 1])[None, :]
            Sk += (self.Bb[k] * self.psi_diag[k][None, :]) @ self.Bb[k].T
            for ft in range(2):
                p = self.Bb[k] @ self.tq[k, ft]
                Sk += self.sig_inv[k, ft] * np.outer(p, p)
            self.S[k] = Sk
        # ---- block Thomas with explicit pivot inverses
        self.twisted = twisted
        self.P = np.zeros((N, NX, NX), t)
        if not twisted:
            for k in range(N):
                St = self.S[k].copy()
                if k >= 1:
                    St -= self.C[k] @ self.P[k - 1] @ self.C[k].T
                self.P[k] = spd_inverse_gauss_jordan(St)
        else:
            # top: 0..3 ascending ; bottom: 9..6 descending ; then 4, then 5
            for k in (0, 1, 2, 3):
                St = self.S[k].copy()
                if k >= 1:
                    St -= self.C[k] @ self.P[k - 1] @ self.C[k].T
                self.P[k] = spd_inverse_gauss_jordan(St)
            for k in (9, 8, 7, 6):
                St = self.S[k].copy()
                if k <= 8:
                    St -= self.C[k + 1].T @ self.P[k + 1] @ self.C[k + 1]
                self.P[k] = spd_inverse_gauss_jordan(St)
            St = self.S[4] - self.C[4] @ self.P[3] @ self.C[4].T
            self.P[4] = spd_inverse_gauss_jordan(St)
            St = self.S[5] - self.C[5] @ self.P[4] @ self.C[5].T - self.C[6].T @ self.P[6] @ self.C[6]
            self.P[5] = spd_inverse_gauss_jordan(St)

    def psi_apply(self, k, v):
        out = self.psi_diag[k] * v
        for ft in range(2):
            q = self.tq[k, ft]
            out = out + q * (self.sig_inv[k, ft] * (q @ v))
        return out

    def solve(self, r1, r2, r3, r4):
        """Solve the full KKT for rhs blocks (r1[240], r2[160], r3[160], r4[140]).
        Returns dx, ds, dz, dy."""
        t = np.float64
        delta = self.delta
        r2 = r2.reshape(N, 16); r3 = r3.reshape(N, 16)
        g = r2 * self.Wi - r3
        Dg = self.D * g
        r1x = r1[:120].reshape(N, NX).copy()
        r1u = r1[120:].reshape(N, NU).copy()
        for k in range(N):
            r1u[k] -= self.Gs[k].T @ Dg[k]
        r4d = r4[:120].reshape(N, NX)
        r4m = r4[120:].reshape(N, 2)
        # v_u = Psi' r1u (+ mx rhs coupling)
        vu = np.zeros((N, NU), t)
        rho = np.zeros((N, NX), t)
        for k in range(N):
            vu[k] = self.psi_apply(k, r1u[k])
            for i, c in enumerate((6, 9)):
                vu[k, c] += self.mx[k, i] / self.mx_den[k, i] * r4m[k, i]
            rho[k] = self.E[k] * self.phix_inv[k] * r1x[k] + self.Bb[k] @ vu[k] - r4d[k]
            if k >= 1:
                rho[k] += self.Ab[k] @ (self.phix_inv[k - 1] * r1x[k - 1])
        # block Thomas
        tt = rho.copy()
        dy = np.zeros((N, NX), t)
        if not self.twisted:
            for k in range(1, N):
                tt[k] -= self.C[k] @ (self.P[k - 1] @ tt[k - 1])
            dy[9] = self.P[9] @ tt[9]
            for k in range(8, -1, -1):
                dy[k] = self.P[k] @ (tt[k] - self.C[k + 1].T @ dy[k + 1])
        else:
            for k in (1, 2, 3):
                tt[k] -= self.C[k] @ (self.P[k - 1] @ tt[k - 1])
            for k in (8, 7, 6):
                tt[k] -= self.C[k + 1].T @ (self.P[k + 1] @ tt[k + 1])
            tt[4] -= self.C[4] @ (self.P[3] @ tt[3])
            tt[5] -= self.C[5] @ (self.P[4] @ tt[4]) + self.C[6].T @ (self.P[6] @ tt[6])
            dy[5] = self.P[5] @ tt[5]
            dy[4] = self.P[4] @ (tt[4] - self.C[5].T @ dy[5])
            for k in (3, 2, 1, 0):
                dy[k] = self.P[k] @ (tt[k] - self.C[k + 1].T @ dy[k + 1])
            for k in (6, 7, 8, 9):
                dy[k] = self.P[k] @ (tt[k] - self.C[k] @ dy[k - 1])
        # recovery
        dxx = np.zeros((N, NX), t)
        duu = np.zeros((N, NU), t)
        dym = np.zeros((N, 2), t)
        for k in range(N):
            w = r1x[k] - self.E[k] * dy[k]
            if k <= 8:
                w -= self.Ab[k + 1].T @ dy[k + 1]
            dxx[k] = self.phix_inv[k] * w
            qv = self.Bb[k].T @ dy[k]
            wv = r1u[k] - qv
            duu[k] = self.psi_apply(k, wv)
            for i, c in enumerate((6, 9)):
                duu[k, c] = (delta * wv[c] + self.mx[k, i] * r4m[k, i]) / self.mx_den[k, i]
                dym[k, i] = (self.mx[k, i] * wv[c] - self.eps_u[k, c] * r4m[k, i]) / self.mx_den[k, i]
        dz = np.zeros((N, 16), t)
        for k in range(N):
            dz[k] = self.D[k] * (self.Gs[k] @ duu[k]) + Dg[k]
        ds = (r2 - dz) * self.Wi
        dx = np.concatenate([dxx.ravel(), duu.ravel()])
        dyv = np.concatenate([dy.ravel(), dym.ravel()])
        return dx, ds.ravel(), dz.ravel(), dyv


def pdipm_iterations_structured(Q_val, G_val, A_val, f, h, b, x, s, z, y, num_iters=5,
                                beta=so.BETA, delta=so.DELTA, twisted=True, trace=None):
    """Same contract as srbd_oracle.pdipm_iterations, structured linear algebra."""
    t = np.float64
    Hm = so.dense_from_ccs(np.asarray(Q_val, t), so.H_COLPTR, so.H_ROW, (NZ, NZ))
    Gm = so.dense_from_ccs(np.asarray(G_val, t), so.G_COLPTR, so.G_ROW, (NINEQ, NZ))
    Am = so.dense_from_ccs(np.asarray(A_val, t), so.A_COLPTR, so.A_ROW, (NEQ, NZ))
    f = np.asarray(f, t).reshape(NZ); h = np.asarray(h, t).reshape(NINEQ); b = np.asarray(b, t).reshape(NEQ)
    x = np.asarray(x, t).reshape(NZ).copy(); s = np.asarray(s, t).reshape(NINEQ).copy()
    z = np.asarray(z, t).reshape(NINEQ).copy(); y = np.asarray(y, t).reshape(NEQ).copy()
    for _ in range(num_iters):
        rx = Hm @ x + f + Gm.T @ z + Am.T @ y
        re = Am @ x - b
        rs = Gm @ x + s - h
        mu = (s @ z) / NINEQ
        step = StructuredStep(Q_val, G_val, A_val, s, z, beta, delta, twisted)
        s_inv = 1.0 / s
        dx_a, ds_a, dz_a, dy_a = step.solve(-rx, -(s_inv * (s * z)), -rs, -re)
        a_pri = so._ratio(s, ds_a); a_dual = so._ratio(z, dz_a)
        mu_aff = ((s + a_pri * ds_a) @ (z + a_dual * dz_a)) / NINEQ
        sigma = (mu_aff / mu) ** 3
        rc = s * z + ds_a * dz_a - sigma * mu
        dx_c, ds_c, dz_c, dy_c = step.solve(np.zeros(NZ), -(s_inv * rc), np.zeros(NINEQ), np.zeros(NEQ))
        dx = dx_a + dx_c; ds = ds_a + ds_c; dz = dz_a + dz_c; dy = dy_a + dy_c
        alpha_pri = so._ratio(s, ds); alpha_dual = so._ratio(z, dz)
        if trace is not None:
            trace.append(dict(dx_aff=dx_a, ds_aff=ds_a, dz_aff=dz_a, dy_aff=dy_a, dx=dx, ds=ds, dz=dz, dy=dy,
                              alpha_pri=alpha_pri, alpha_dual=alpha_dual, sigma=sigma, mu=mu))
        x = x + alpha_pri * dx; s = s + alpha_pri * ds
        z = z + alpha_dual * dz; y = y + alpha_dual * dy
        s = np.maximum(s, 1e-8); z = np.maximum(z, 1e-8)
        mu_new = (s @ z) / NINEQ
    res = np.array([np.linalg.norm(rx), np.linalg.norm(rs), np.linalg.norm(re), mu_new])
    return x, s, z, y, res, np.array([mu_new])
