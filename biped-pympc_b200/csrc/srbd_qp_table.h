// Per-environment "value table" of the QP assembly: every distinct number that appears in the
// reference's `qp_former` outputs (biped_pympc/casadi/srbd_constraints.py:20-227 with the dynamics of
// biped_pympc/casadi/srbd_centroidal_model.py:101-166), computed by one warp from closed forms
// (A_c is nilpotent, so RK4 is exact: A_d = I + dt A_c, B_d = dt B_c + dt^2/2 A_c B_c,
// c_d = dt c + dt^2/2 A_c c).  Shared by the assembly kernel (srbd_qp_form.cu), which streams the table
// out through the frozen CCS patterns, and by the fused tick of the PDIPM kernel (srbd_pdipm_fast.h),
// which builds its stage data from it without going through the CCS vectors.
#pragma once
#include "srbd_pattern.h"
#include "srbd_warp.h"

namespace srbd {

// ---- value-table layout -----------------------------------------------------------------
enum : int {
  V_ZERO = 0, V_ONE = 1, V_NEG1 = 2, V_NEGMU = 3, V_NEGLT = 4, V_NEGLH = 5,
  V_Q = 6, V_R = 18,
  V_NADTR = 30,   // -dt * R[i][j]  at 30 + 3i + j
  V_NADT = 39,    // -dt
  V_NBD = 40,     // -B_d[r][c]     at 40 + 12r + c
  V_CD = 184,     // c_d
  V_B0 = 196,     // c_d + A_d x0
  V_FC = 208,     // F_max * contact[stage][foot] at 208 + 2*stage + foot
  V_COUNT = 228
};
static_assert(V_COUNT <= 256, "tables are uint8");

// raw per-env inputs staged in shared memory
enum : int { RW_RB = 0, RW_IW = 9, RW_BODY = 18, RW_LF = 21, RW_RF = 24, RW_RL = 27, RW_RA = 30,
             RW_X0 = 33, RW_CT = 45, RW_DT = 65, RW_M = 66, RW_MU = 67, RW_IINV = 68, RW_BCW = 77,
             RW_COUNT = 113 };   // BCW: 3 x 12 angular-acceleration map  I^-1 [skew(rL) skew(rR) I I]


// V: V_COUNT doubles, W: RW_COUNT doubles of warp-private scratch; in[] = the 17 qp_former inputs
// (device pointers, env-major); all 32 lanes of the warp must call
// `lane` in [0, nl): nl cooperating lanes (32 = a whole warp per env, 16 = half a warp per env)
SRBD_DEV void qp_value_table(const double* const* in, long long e, int lane0, int nl, const Warp& w, double* V, double* W) {
  // ---- stage raw inputs (nnz order = CasADi column-major) ---------------------------
  if (nl == 32) {
    // every load of the lane is issued before the first store: when the inputs are mapped host memory (host entry,
    // read in place over PCIe) the env then pays one round trip, not one per input array
    const int l = lane0;
    double t_rb = 0, t_iw = 0, t_b0 = 0, t_b1 = 0, t_b2 = 0, t_b3 = 0, t_b4 = 0, t_x0 = 0, t_q = 0, t_r = 0, t_ct = 0, t_dt = 0, t_m = 0, t_mu = 0;
    if (l < 9) { t_rb = in[IN_RBODY][e * 9 + l]; t_iw = in[IN_IWORLD][e * 9 + l]; }
    if (l < 3) {
      t_b0 = in[IN_BODY][e * 3 + l]; t_b1 = in[IN_LFOOT][e * 3 + l]; t_b2 = in[IN_RFOOT][e * 3 + l];
      t_b3 = in[IN_RESLIN][e * 3 + l]; t_b4 = in[IN_RESANG][e * 3 + l];
    }
    if (l < 12) { t_x0 = in[IN_X0][e * 12 + l]; t_q = in[IN_Q][e * 12 + l]; t_r = in[IN_R][e * 12 + l]; }
    if (l < 20) t_ct = in[IN_CONTACT][e * 20 + l];
    if (l == 0) { t_dt = in[IN_DT][e]; t_m = in[IN_M][e]; t_mu = in[IN_MU][e]; }
    if (l < 9) { W[RW_RB + l] = t_rb; W[RW_IW + l] = t_iw; }
    if (l < 3) { W[RW_BODY + l] = t_b0; W[RW_LF + l] = t_b1; W[RW_RF + l] = t_b2; W[RW_RL + l] = t_b3; W[RW_RA + l] = t_b4; }
    if (l < 12) { W[RW_X0 + l] = t_x0; V[V_Q + l] = t_q; V[V_R + l] = t_r; }
    if (l < 20) W[RW_CT + l] = t_ct;
    if (l == 0) { W[RW_DT] = t_dt; W[RW_M] = t_m; W[RW_MU] = t_mu; }
  } else {
  for (int lane = lane0; lane < 9; lane += nl) { W[RW_RB + lane] = in[IN_RBODY][e * 9 + lane]; W[RW_IW + lane] = in[IN_IWORLD][e * 9 + lane]; }
  for (int lane = lane0; lane < 3; lane += nl) {
    W[RW_BODY + lane] = in[IN_BODY][e * 3 + lane];
    W[RW_LF + lane] = in[IN_LFOOT][e * 3 + lane];
    W[RW_RF + lane] = in[IN_RFOOT][e * 3 + lane];
    W[RW_RL + lane] = in[IN_RESLIN][e * 3 + lane];
    W[RW_RA + lane] = in[IN_RESANG][e * 3 + lane];
  }
  for (int lane = lane0; lane < 12; lane += nl) {
    W[RW_X0 + lane] = in[IN_X0][e * 12 + lane];
    V[V_Q + lane] = in[IN_Q][e * 12 + lane];
    V[V_R + lane] = in[IN_R][e * 12 + lane];
  }
  for (int lane = lane0; lane < 20; lane += nl) W[RW_CT + lane] = in[IN_CONTACT][e * 20 + lane];
  if (lane0 == 0) {
    W[RW_DT] = in[IN_DT][e]; W[RW_M] = in[IN_M][e]; W[RW_MU] = in[IN_MU][e];
  }
  }
  w.sync();
  const double dt = W[RW_DT], m = W[RW_M], mu = W[RW_MU];
  const double hdt2 = dt * dt / 2;
  // ---- I_world^-1 (adjugate / det); M(i,j) = Iw_flat[i + 3j] ------------------------
  for (int lane = lane0; lane < 9; lane += nl) {
    const double* M = W + RW_IW;
    auto Mx = [&](int i, int j) { return M[i + 3 * j]; };
    const double det = Mx(0, 0) * (Mx(1, 1) * Mx(2, 2) - Mx(1, 2) * Mx(2, 1))
                     + Mx(0, 1) * (Mx(1, 2) * Mx(2, 0) - Mx(1, 0) * Mx(2, 2))
                     + Mx(0, 2) * (Mx(1, 0) * Mx(2, 1) - Mx(1, 1) * Mx(2, 0));
    const int i = lane / 3, j = lane % 3;
    const int j1 = (j + 1) % 3, j2 = (j + 2) % 3, i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    W[RW_IINV + 3 * i + j] = (Mx(j1, i1) * Mx(j2, i2) - Mx(j1, i2) * Mx(j2, i1)) / det;
  }
  w.sync();
  // ---- angular-acceleration map: I^-1 [skew(rL) | skew(rR) | I | I]   (3 x 12) ------
  for (int lane = lane0; lane < 27; lane += nl) {
   if (lane < 18) {
    const int i = lane / 6, c = lane % 6, ft = c / 3, j = c % 3;
    const double* foot = W + (ft == 0 ? RW_LF : RW_RF);
    const double rx = foot[0] - W[RW_BODY + 0], ry = foot[1] - W[RW_BODY + 1], rz = foot[2] - W[RW_BODY + 2];
    const double* Ii = W + RW_IINV + 3 * i;
    double v;
    if (j == 0) v = Ii[1] * rz - Ii[2] * ry;
    else if (j == 1) v = Ii[2] * rx - Ii[0] * rz;
    else v = Ii[0] * ry - Ii[1] * rx;
    W[RW_BCW + 12 * i + c] = v;
   } else {
    const int i = (lane - 18) / 3, j = (lane - 18) % 3;
    const double v = W[RW_IINV + 3 * i + j];
    W[RW_BCW + 12 * i + 6 + j] = v;
    W[RW_BCW + 12 * i + 9 + j] = v;
   }
  }
  w.sync();
  // ---- constants, -dt R, contact ----------------------------------------------------
  if (lane0 == 0) {
    V[V_ZERO] = 0.0; V[V_ONE] = 1.0; V[V_NEG1] = -1.0; V[V_NEGMU] = -mu;
    V[V_NEGLT] = -LT; V[V_NEGLH] = -LH; V[V_NADT] = -dt;
  }
  for (int lane = lane0; lane < 9; lane += nl) {   // R(i,j) = Rb_flat[i + 3j]
    const int i = lane / 3, j = lane % 3;
    V[V_NADTR + 3 * i + j] = -(dt * W[RW_RB + i + 3 * j]);
  }
  for (int lane = lane0; lane < 20; lane += nl) {  // contact(stage, foot) = flat[stage + 10*foot]; F_max * contact
    const int st = lane % 10, ft = lane / 10;
    V[V_FC + 2 * st + ft] = F_MAX * W[RW_CT + lane];
  }
  // ---- -B_d (12 x 12) ----------------------------------------------------------------
  for (int idx = lane0; idx < 144; idx += nl) {
    const int r = idx / 12, c = idx % 12;
    double v = 0.0;
    if (r < 3) {          // euler rows: dt^2/2 * R * Bc_omega
      const double* Rb = W + RW_RB;
      v = hdt2 * (Rb[r] * W[RW_BCW + c] + Rb[r + 3] * W[RW_BCW + 12 + c] + Rb[r + 6] * W[RW_BCW + 24 + c]);
    } else if (r < 6) {   // position rows: dt^2/2 / m on f_L[r-3], f_R[r-3]
      if (c < 6 && c % 3 == r - 3) v = hdt2 * (1.0 / m);
    } else if (r < 9) {   // omega rows
      v = dt * W[RW_BCW + 12 * (r - 6) + c];
    } else {              // velocity rows
      if (c < 6 && c % 3 == r - 9) v = dt * (1.0 / m);
    }
    V[V_NBD + idx] = -v;
  }
  // ---- c_d and b_0 = c_d + A_d x0 -----------------------------------------------------
  for (int lane = lane0; lane < 12; lane += nl) {
    const int r = lane, i = r % 3;
    const double* Rb = W + RW_RB;
    const double* x0 = W + RW_X0;
    const double gl = (i == 2 ? GRAV_Z : 0.0) + W[RW_RL + i];   // gravity + residual lin accel
    double cd, ax;
    if (r < 3) {
      cd = hdt2 * (Rb[i] * W[RW_RA] + Rb[i + 3] * W[RW_RA + 1] + Rb[i + 6] * W[RW_RA + 2]);
      ax = x0[r] + dt * (Rb[i] * x0[6] + Rb[i + 3] * x0[7] + Rb[i + 6] * x0[8]);
    } else if (r < 6) {
      cd = hdt2 * gl;
      ax = x0[r] + dt * x0[r + 6];
    } else if (r < 9) {
      cd = dt * W[RW_RA + i];
      ax = x0[r];
    } else {
      cd = dt * gl;
      ax = x0[r];
    }
    V[V_CD + r] = cd;
    V[V_B0 + r] = cd + ax;
  }
  w.sync();
}

}  // namespace srbd
