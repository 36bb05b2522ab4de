"""Developer study (CPU, numpy): what an fp32 factorisation of the ten Schur pivot blocks costs in accuracy, and what
iterative refinement against the fp64 Schur operator buys back.  Uses the structured numpy prototype of the kernel's
Newton step (oracle/structured_proto.py); the reference for the error is the 80-bit C oracle.

    python tests/tools/fp32_mode_study.py [envs] [iters]

Prints, per number of refinement steps, the worst norm-wise relative error of x, s, z, y over the batch after `iters`
Newton iterations.  The numbers are quoted in DESIGN.md section 8 (fp32 mode)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import biped_pympc_b200 as bp  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import srbd_oracle as so  # noqa: E402
from oracle import structured_proto as sp  # noqa: E402

N, NX = sp.N, sp.NX


class Step32(sp.StructuredStep):
    """StructuredStep with the pivot blocks formed, coupled and inverted in fp32 (what PdipmFastT<true> does) and
    `refine` steps of iterative refinement of dy against the fp64 block-tridiagonal operator."""
    refine = 0

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        f = np.float32
        S32 = self.S.astype(f); C32 = self.C.astype(f)
        P = np.zeros((N, NX, NX), f)
        inv = lambda M: sp.spd_inverse_gauss_jordan(M.astype(f)).astype(f)
        for k in (0, 1, 2, 3):
            St = S32[k].copy()
            if k >= 1:
                St -= C32[k] @ P[k - 1] @ C32[k].T
            P[k] = inv(St)
        for k in (9, 8, 7, 6):
            St = S32[k].copy()
            if k <= 8:
                St -= C32[k + 1].T @ P[k + 1] @ C32[k + 1]
            P[k] = inv(St)
        P[4] = inv(S32[4] - C32[4] @ P[3] @ C32[4].T)
        P[5] = inv(S32[5] - C32[5] @ P[4] @ C32[5].T - C32[6].T @ P[6] @ C32[6])
        self.P = P.astype(np.float64)

    def thomas(self, rho):
        tt = rho.copy(); dy = np.zeros_like(rho)
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
        return dy

    def apply_S(self, dy):
        out = np.zeros_like(dy)
        for k in range(N):
            out[k] = self.S[k] @ dy[k]
            if k >= 1:
                out[k] += self.C[k] @ dy[k - 1]
            if k <= 8:
                out[k] += self.C[k + 1].T @ dy[k + 1]
        return out

    def solve(self, r1, r2, r3, r4):
        # same reduction / recovery as the parent; only the dy solve differs
        outer = self

        class _Hook:
            pass
        # re-implement by temporarily replacing the Thomas part: compute rho as the parent does
        t = np.float64
        delta = self.delta
        r2 = r2.reshape(N, 16); r3 = r3.reshape(N, 16)
        g = r2 * self.Wi - r3
        Dg = self.D * g
        r1x = r1[:120].reshape(N, NX).copy(); r1u = r1[120:].reshape(N, sp.NU).copy()
        for k in range(N):
            r1u[k] -= self.Gs[k].T @ Dg[k]
        r4d = r4[:120].reshape(N, NX); r4m = r4[120:].reshape(N, 2)
        vu = np.zeros((N, sp.NU), t); rho = np.zeros((N, NX), t)
        for k in range(N):
            vu[k] = self.psi_apply(k, r1u[k])
            for i, c in enumerate((6, 9)):
                vu[k, c] += self.mx[k, i] / self.mx_den[k, i] * r4m[k, i]
            rho[k] = self.E[k] * self.phix_inv[k] * r1x[k] + self.Bb[k] @ vu[k] - r4d[k]
            if k >= 1:
                rho[k] += self.Ab[k] @ (self.phix_inv[k - 1] * r1x[k - 1])
        dy = self.thomas(rho)
        for _ in range(self.refine):
            dy = dy + self.thomas(rho - self.apply_S(dy))
        dxx = np.zeros((N, NX), t); duu = np.zeros((N, sp.NU), t); dym = np.zeros((N, 2), t)
        for k in range(N):
            w = r1x[k] - self.E[k] * dy[k]
            if k <= 8:
                w -= self.Ab[k + 1].T @ dy[k + 1]
            dxx[k] = self.phix_inv[k] * w
            wv = r1u[k] - self.Bb[k].T @ dy[k]
            duu[k] = self.psi_apply(k, wv)
            for i, c in enumerate((6, 9)):
                duu[k, c] = (delta * wv[c] + self.mx[k, i] * r4m[k, i]) / self.mx_den[k, i]
                dym[k, i] = (self.mx[k, i] * wv[c] - self.eps_u[k, c] * r4m[k, i]) / self.mx_den[k, i]
        dz = np.zeros((N, 16), t)
        for k in range(N):
            dz[k] = self.D[k] * (self.Gs[k] @ duu[k]) + Dg[k]
        ds = (r2 - dz) * self.Wi
        return np.concatenate([dxx.ravel(), duu.ravel()]), ds.ravel(), dz.ravel(), np.concatenate([dy.ravel(), dym.ravel()])


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    inp, _ = bp.workload.make_batch(B, seed=3)
    qp = co.qp_former(inp)
    ref = co.pdipm(qp, iters, long_double=True)
    Hv, f, Av, b, Gv, d = qp
    orig = sp.StructuredStep
    print("refinement_steps,err_x,err_s,err_z,err_y   (%d envs, %d Newton iterations, vs the 80-bit oracle)" % (B, iters))
    for refine in (None, 0, 1, 2, 3):
        if refine is None:
            sp.StructuredStep = orig
        else:
            Step32.refine = refine
            sp.StructuredStep = Step32
        errs = np.zeros(4)
        for e in range(B):
            x0, s0, z0, y0 = so.cold_start(Gv[e], d[e])
            r = sp.pdipm_iterations_structured(Hv[e], Gv[e], Av[e], f[e], d[e], b[e], x0, s0, z0, y0, num_iters=iters)
            for i in range(4):
                errs[i] = max(errs[i], np.linalg.norm(r[i] - ref[i][e]) / np.linalg.norm(ref[i][e]))
        print(("fp64" if refine is None else str(refine)) + "," + ",".join("%.1e" % v for v in errs))
    sp.StructuredStep = orig


if __name__ == "__main__":
    main()
