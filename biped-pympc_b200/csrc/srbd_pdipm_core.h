// PDIPM Newton iterations for one environment, executed by one warp.
//
// B200-native replacement of the reference's CusADi-generated
// `sparse_pdipm_multiple_iterations` (biped_pympc/casadi/sparse_pdipm_solver.py:357-534):
// Mehrotra predictor-corrector on the 240-variable / 140-equality / 160-inequality
// SRBD-MPC QP.  The reference factors the full 700x700 regularised KKT with a symbolic
// sparse LDL^T; here the same linear system (including beta / delta) is solved by the
// exact block elimination described in DESIGN.md ("PDIPM kernel"):
//
//   W = z/s + delta, D = 1/(1/W + delta)           (ds, dz eliminated row-wise)
//   Phi_u = H_u + beta + G' D G                    per stage: two "arrow" 4x4 blocks on
//                                                  (fx,fy,my | fz) + diagonal mx, mz;
//                                                  inverse = diag + rank-1 per foot, with the
//                                                  Schur term computed cancellation-free
//   S = A Phi^-1 A' + delta I                      block tridiagonal, ten 12x12 blocks;
//                                                  Mx rows eliminated in closed form
//   block-Thomas with explicit SPD pivot inverses  "twisted": the two half-warps eliminate
//                                                  from both ends of the horizon
//
// Thread mapping: lane = 16*h + r.  Half h works on stage k = t (h = 0) or 9 - t (h = 1) in
// round t; r is a row (state / input / dynamics-row / inequality-row index) of that stage.
//
// The same source is compiled for the GPU (srbd_pdipm.cu) and, for tests only, for the host
// under the fiber-based warp emulator of srbd_warp.h.
#pragma once

#include "srbd_args.h"
#include "srbd_pattern.h"
#include "srbd_warp.h"

namespace srbd {

// ---- per-warp shared-memory layout (doubles) --------------------------------------------
enum : int {
  SB_B = 0,        // Bbar_k dense [12][14]
  SB_E = 168,      // E_k[12]
  SB_AP = 180,     // Abar_k     compact: diag[12] | R[3][3] (rows 0..2, cols 6..8) | pv[3]
  SB_AN = 204,     // Abar_{k+1} compact
  SB_MX = 228,     // Mx-row coefficients [2]
  SB_G = 230,      // G values of the stage [28]
  SB_HX = 258,     // hx_k[12]   (Q_val of x_{k+1})
  SB_PHX = 270,    // 1/(hx_k + beta)
  SB_HU = 282,     // hu_k[12]
  SB_PHXP = 294,   // 1/(hx_{k-1} + beta)
  SB_EP = 306,     // E_{k-1}[12]
  SB_DER = 318,    // derived: psi_mz[2], psi_mx'[2], 1/mx_den[2], eps_mx[2]
  SB_SIZE = 328
};
enum : int { ST_ZERO = 0, ST_AU, ST_E, ST_AXPREV, ST_AXCUR, ST_G, ST_Q, ST_QINV, ST_QINVPREV, ST_EPREV };
struct StageSlot { unsigned char kind; short rel; };

template <bool kUniform>
struct PdipmSmem {
  static constexpr int NSB = kUniform ? 1 : 2;
  static constexpr int OFF_P = 0;          // 10 x 78 packed lower; scratch for iterates in phase A
  static constexpr int OFF_VX = 780;       // 240  r1~ -> dx
  static constexpr int OFF_VY = 1020;      // 140  re -> rho -> t -> dy
  static constexpr int OFF_D = 1160;       // 160
  static constexpr int OFF_WI = 1320;      // 160
  static constexpr int OFF_DG = 1480;      // 160  D*g of the current solve ; later ds
  static constexpr int OFF_R2 = 1640;      // 160  r2 of the current solve  ; later dz
  static constexpr int OFF_PSI = 1800;     // 10 x 16
  static constexpr int OFF_BC = 1960;      // 2 x 36 broadcast scratch per half
  static constexpr int OFF_CM = 2032;      // 2 x 36 coupling coefficient slots
  static constexpr int OFF_SB = 2104;      // NSB x SB_SIZE
  static constexpr int SIZE = OFF_SB + NSB * SB_SIZE;
};

SRBD_DEV int pk(int i, int j) { return i >= j ? ((i * (i + 1)) >> 1) + j : ((j * (j + 1)) >> 1) + i; }

template <bool kUniform>
struct PdipmWarp {
  using SM = PdipmSmem<kUniform>;
  Warp w;
  double* sm;                    // this warp's shared memory
  const StageSlot* slots;        // gather table (SB_SIZE entries)
  // per-env global pointers
  const double *gQ, *gG, *gA, *gF, *gH, *gB;
  double *oX, *oS, *oZ, *oY, *oRes, *oMu, *ws;
  int lane, h, r;

  SRBD_DEV double* P(int k) const { return sm + SM::OFF_P + 78 * k; }
  SRBD_DEV double* sb(int hh) const { return sm + SM::OFF_SB + (kUniform ? 0 : hh * SB_SIZE); }
  SRBD_DEV double* bc(int hh) const { return sm + SM::OFF_BC + 36 * hh; }
  SRBD_DEV double* cm(int slot) const { return sm + SM::OFF_CM + 36 * slot; }
  SRBD_DEV double* psi(int k) const { return sm + SM::OFF_PSI + 16 * k; }

  // ------------------------------------------------------------------------------------
  // gather the stage-k matrices from the CCS value vectors into the half's stage buffer
  // ------------------------------------------------------------------------------------
  SRBD_DEV void load_stage_raw(int hh, int k, int first, int step) const {
    double* s = sb(hh);
    for (int i = first; i < SB_DER; i += step) {
      const StageSlot sl = slots[i];
      double v = 0.0;
      switch (sl.kind) {
        case ST_AU: v = ldg(gA + 336 + 86 * k + sl.rel); break;
        case ST_E: v = ldg(gA + a_xcol_start(k, sl.rel)); break;
        case ST_AXPREV: if (k >= 1) v = ldg(gA + 36 * (k - 1) + sl.rel); break;
        case ST_AXCUR: if (k <= 8) v = ldg(gA + 36 * k + sl.rel); break;
        case ST_G: v = ldg(gG + 28 * k + sl.rel); break;
        case ST_Q: v = ldg(gQ + 12 * k + sl.rel); break;
        case ST_QINV: v = 1.0 / (ldg(gQ + 12 * k + sl.rel) + IP_BETA); break;
        case ST_QINVPREV: if (k >= 1) v = 1.0 / (ldg(gQ + 12 * (k - 1) + sl.rel) + IP_BETA); break;
        case ST_EPREV: if (k >= 1) v = ldg(gA + a_xcol_start(k - 1, sl.rel)); break;
        default: break;
      }
      s[i] = v;
    }
  }
  SRBD_DEV void load_stage_derived(int hh, int first) const {
    double* s = sb(hh);
    if (first < 2) {              // mz: plain diagonal of Phi_u
      s[SB_DER + first] = 1.0 / (s[SB_HU + 8 + 3 * first] + IP_BETA);
    } else if (first < 4) {       // mx: eliminated together with its equality row
      const int i = first - 2;
      const double eps = s[SB_HU + 6 + 3 * i] + IP_BETA;
      const double m = s[SB_MX + i];
      const double iden = 1.0 / (m * m + IP_DELTA * eps);
      s[SB_DER + 2 + i] = IP_DELTA * iden;   // psi'
      s[SB_DER + 4 + i] = iden;
      s[SB_DER + 6 + i] = eps;
    }
  }
  // per-round (general) or once (stage-invariant) stage load; all lanes must call
  SRBD_DEV void stage(int k) const {
    if (!kUniform) {
      w.sync();
      load_stage_raw(h, k, r, 16);
      w.sync();
      load_stage_derived(h, r);
      w.sync();
    }
  }
  SRBD_DEV void stage_once() const {
    if (kUniform) {
      load_stage_raw(0, 1, lane, 32);
      w.sync();
      load_stage_derived(0, lane);
      w.sync();
    }
  }

  // ------------------------------------------------------------------------------------
  // sparse coupling rows.  type 0 ("LP"): row r of C_k = Abar_k diag(phix_{k-1} E_{k-1});
  //                        type 1 ("LN"): row r of C_{k+1}' , C_{k+1} = Abar_{k+1} diag(phix_k E_k)
  // ------------------------------------------------------------------------------------
  SRBD_DEV void coupling_row(int type, const double* s, int idx[4], double val[4]) const {
    const int rr = r < 12 ? r : 0;
    idx[0] = rr; idx[1] = rr; idx[2] = rr; idx[3] = rr;
    val[0] = 0.0; val[1] = 0.0; val[2] = 0.0; val[3] = 0.0;
    if (r >= 12) return;
    if (type == 0) {
      const double* A = s + SB_AP;
      const double* ph = s + SB_PHXP;
      const double* e = s + SB_EP;
      val[0] = A[rr] * (ph[rr] * e[rr]);
      if (rr < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { idx[1 + j] = 6 + j; val[1 + j] = A[12 + 3 * rr + j] * (ph[6 + j] * e[6 + j]); }
      } else if (rr < 6) {
        idx[1] = rr + 6; val[1] = A[21 + rr - 3] * (ph[rr + 6] * e[rr + 6]);
      }
    } else {
      const double* A = s + SB_AN;
      const double wr = s[SB_PHX + rr] * s[SB_E + rr];
      val[0] = A[rr] * wr;
      if (rr >= 6 && rr < 9) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { idx[1 + i] = i; val[1 + i] = A[12 + 3 * i + (rr - 6)] * wr; }
      } else if (rr >= 9) {
        idx[1] = rr - 6; val[1] = A[21 + rr - 9] * wr;
      }
    }
  }
  // publish row r's coupling coefficients into the 36 static slots used by the congruence
  SRBD_DEV void publish_cm(double* c, int type, const double val[4]) const {
    if (r >= 12) return;
    c[r] = val[0];
    if (r < 3) { for (int j = 0; j < 3; ++j) c[12 + 3 * r + j] = type == 0 ? val[1 + j] : 0.0; }
    else if (r < 6) { c[21 + r - 3] = type == 0 ? val[1] : 0.0; }
    else if (r < 9) { for (int i = 0; i < 3; ++i) c[24 + 3 * (r - 6) + i] = type == 1 ? val[1 + i] : 0.0; }
    else { c[33 + r - 9] = type == 1 ? val[1] : 0.0; }
  }

  // ------------------------------------------------------------------------------------
  // Psi_k applied to a 12-vector held in smem at `v`; returns entry c
  // ------------------------------------------------------------------------------------
  SRBD_DEV double psi_apply(const double* ps, const double* der, const double* v, int c) const {
    // ps: [ft*7 + {0..2: 1/a (fx,fy,my), 3..5: t (fx,fy,my), 6: 1/sigma}]
    if (c == 8 || c == 11) return der[c == 8 ? 0 : 1] * v[c];
    if (c == 6 || c == 9) return der[2 + (c == 9)] * v[c];   // psi' (mx coupling added by caller)
    const int ft = (c < 6) ? c / 3 : (c == 10);
    const double* p = ps + 7 * ft;
    const int fx = 3 * ft, fy = fx + 1, fz = fx + 2, my = 7 + 3 * ft;
    const double dot = p[3] * v[fx] + p[4] * v[fy] + p[5] * v[my] - v[fz];
    double diag, q;
    if (c == fz) { diag = 0.0; q = -1.0; }
    else { const int i = (c == fx) ? 0 : (c == fy ? 1 : 2); diag = p[i]; q = p[3 + i]; }
    return diag * v[c] + q * (p[6] * dot);
  }

  // ------------------------------------------------------------------------------------
  // build rho_k (into vy) from r1~ (vx) and the Mx/dynamics right-hand sides (vy)
  // on entry: vy[12k+r] = -r4_dyn (i.e. `re`, or 0 for the corrector), vy[120+2k+i] = -r4_mx
  // ------------------------------------------------------------------------------------
  SRBD_DEV void build_rho(int k) const {
    const double* s = sb(h);
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    if (r < 12) {
      double v = psi_apply(psi(k), s + SB_DER, vx + 120 + 12 * k, r);
      if (r == 6 || r == 9) {
        const int i = (r == 9);
        v += s[SB_MX + i] * s[SB_DER + 4 + i] * (-vy[120 + 2 * k + i]);   // (m/den) * r4_mx
      }
      b[r] = v;
    }
    w.sync();
    if (r < 12) {
      double rho = s[SB_E + r] * s[SB_PHX + r] * vx[12 * k + r] + vy[12 * k + r];
      const double* Br = s + SB_B + 14 * r;
#pragma unroll
      for (int c = 0; c < 12; ++c) rho += Br[c] * b[c];
      if (k >= 1) {
        const double* A = s + SB_AP;
        const double* ph = s + SB_PHXP;
        const double* xp = vx + 12 * (k - 1);
        rho += A[r] * ph[r] * xp[r];
        if (r < 3) { for (int j = 0; j < 3; ++j) rho += A[12 + 3 * r + j] * ph[6 + j] * xp[6 + j]; }
        else if (r < 6) rho += A[21 + r - 3] * ph[r + 6] * xp[r + 6];
      }
      vy[12 * k + r] = rho;
    }
    w.sync();
  }

  // ------------------------------------------------------------------------------------
  // form S_kk (packed lower) into P[k]
  // ------------------------------------------------------------------------------------
  SRBD_DEV void form_S(int k) const {
    const double* s = sb(h);
    const double* ps = psi(k);
    const double* der = s + SB_DER;
    double* b = bc(h);
    const int rr = r < 12 ? r : 11;
    double wv[12];
    const double* Br = s + SB_B + 14 * rr;
    double pL, pR;
    {
      double Brow[12];
#pragma unroll
      for (int c = 0; c < 12; ++c) Brow[c] = Br[c];
      pL = ps[3] * Brow[0] + ps[4] * Brow[1] + ps[5] * Brow[7] - Brow[2];
      pR = ps[10] * Brow[3] + ps[11] * Brow[4] + ps[12] * Brow[10] - Brow[5];
      wv[0] = ps[0] * Brow[0]; wv[1] = ps[1] * Brow[1]; wv[2] = 0.0;
      wv[3] = ps[7] * Brow[3]; wv[4] = ps[8] * Brow[4]; wv[5] = 0.0;
      wv[6] = der[2] * Brow[6]; wv[7] = ps[2] * Brow[7]; wv[8] = der[0] * Brow[8];
      wv[9] = der[3] * Brow[9]; wv[10] = ps[9] * Brow[10]; wv[11] = der[1] * Brow[11];
    }
    if (r < 12) { b[12 + r] = pL; b[24 + r] = pR; }
    w.sync();
    const double sL = ps[6] * pL, sR = ps[13] * pR;
    // x-part of row rr: sparse row of Abar_k scaled by phix_{k-1}
    int ai[4]; double av[4];
    {
      const double* A = s + SB_AP;
      const double* ph = s + SB_PHXP;
      ai[0] = rr; ai[1] = rr; ai[2] = rr; ai[3] = rr;
      av[0] = 0.0; av[1] = 0.0; av[2] = 0.0; av[3] = 0.0;
      if (k >= 1) {
        av[0] = A[rr] * ph[rr];
        if (rr < 3) { for (int j = 0; j < 3; ++j) { ai[1 + j] = 6 + j; av[1 + j] = A[12 + 3 * rr + j] * ph[6 + j]; } }
        else if (rr < 6) { ai[1] = rr + 6; av[1] = A[21 + rr - 3] * ph[rr + 6]; }
      }
    }
    double* Pk = P(k);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const double* Bj = s + SB_B + 14 * j;
      double acc = sL * b[12 + j] + sR * b[24 + j];
#pragma unroll
      for (int c = 0; c < 12; ++c) acc += wv[c] * Bj[c];
      // Abar_k(j, c) for this lane's (<= 4) columns c
      const double* A = s + SB_AP;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int c = ai[a];
        double ajc = 0.0;
        if (c == j) ajc = A[j];
        else if (j < 3 && c >= 6 && c < 9) ajc = A[12 + 3 * j + (c - 6)];
        else if (j >= 3 && j < 6 && c == j + 6) ajc = A[21 + j - 3];
        acc += av[a] * ajc;
      }
      if (j == rr) acc += s[SB_E + rr] * s[SB_E + rr] * s[SB_PHX + rr] + IP_DELTA;
      if (r < 12 && j <= r) Pk[pk(r, j)] = acc;
    }
    w.sync();
  }

  // ------------------------------------------------------------------------------------
  // one elimination round: a = row rr of S_kk - M P_nb M'  (M = coupling of `type`);
  // mode 0: invert (symmetric sweep) and store P_k ; mode 1: store the updated S~ row only
  // ------------------------------------------------------------------------------------
  SRBD_DEV void factor_round(int k, int nb, int type, int mode, bool act) const {
    const int rr = r < 12 ? r : 11;
    double a[12];
    double* Pk = P(k);
#pragma unroll
    for (int c = 0; c < 12; ++c) a[c] = Pk[pk(rr, c)];
    w.sync();   // everyone has read S_kk before anyone overwrites it
    if (nb >= 0) {
      int idx[4]; double val[4];
      coupling_row(type, sb(h), idx, val);
      double* c = cm(h);
      publish_cm(c, type, val);
      w.sync();
      const double* Pn = P(nb);
      double U[12];
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) u += val[q] * Pn[pk(idx[q], m)];
        U[m] = u;
      }
#pragma unroll
      for (int cc = 0; cc < 12; ++cc) {
        double o = U[cc] * c[cc];
        if (cc < 3) o += U[6] * c[12 + 3 * cc] + U[7] * c[12 + 3 * cc + 1] + U[8] * c[12 + 3 * cc + 2];
        else if (cc < 6) o += U[cc + 6] * c[21 + cc - 3];
        else if (cc < 9) o += U[0] * c[24 + 3 * (cc - 6)] + U[1] * c[24 + 3 * (cc - 6) + 1] + U[2] * c[24 + 3 * (cc - 6) + 2];
        else o += U[cc - 6] * c[33 + cc - 9];
        a[cc] -= o;
      }
    }
    if (mode == 1) {
      if (act && r < 12) {
#pragma unroll
        for (int c = 0; c < 12; ++c) if (c <= r) Pk[pk(r, c)] = a[c];
      }
      // fall through: the sweep below is executed (results discarded) to keep the warp convergent
    }
    // symmetric sweep on all 12 pivots: every lane keeps one full row, the unscaled pivot row
    // is broadcast through shared memory; the result is -S~^-1 (negated on store).  Unlike a
    // plain Gauss-Jordan inverse this stays symmetric, so keeping only the lower triangle
    // costs no accuracy.
    double* b = bc(h);
#pragma unroll
    for (int p = 0; p < 12; ++p) {
      if (r == p) {
#pragma unroll
        for (int c = 0; c < 12; ++c) b[c] = a[c];
      }
      w.sync();
      const double di = 1.0 / b[p];
      const double t = b[rr] * di;
      if (r == p) {
#pragma unroll
        for (int c = 0; c < 12; ++c) a[c] = (c == p) ? -di : b[c] * di;
      } else {
#pragma unroll
        for (int c = 0; c < 12; ++c) a[c] = (c == p) ? t : a[c] - t * b[c];
      }
      w.sync();
    }
    if (mode == 0 && act && r < 12) {
#pragma unroll
      for (int c = 0; c < 12; ++c) if (c <= r) Pk[pk(r, c)] = -a[c];
    }
    w.sync();
  }

  SRBD_DEV void factor() const {
    // rounds 0..3: top k = t (nb k-1, LP) | bottom k = 9-t (nb k+1, LN)
    for (int t = 0; t < 4; ++t) {
      const int k = h ? 9 - t : t;
      stage(k);
      factor_round(k, t == 0 ? -1 : (h ? k + 1 : k - 1), h, 0, true);
    }
    // round 4: top factors k=4 ; bottom applies the k=6 update to S_55 (store only)
    { const int k = h ? 5 : 4; stage(k); factor_round(k, h ? 6 : 3, h, h ? 1 : 0, true); }
    // round 5: top factors k=5 (update from k=4) ; bottom idle
    { stage(5); factor_round(5, 4, 0, 0, h == 0); }
  }

  // ------------------------------------------------------------------------------------
  // block-Thomas solve on vy[0..119]  (rho -> t -> dy), twisted
  // ------------------------------------------------------------------------------------
  SRBD_DEV void fwd_round(int k, int nb, int type, bool act) const {
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    const int rr = r < 12 ? r : 11;
    const double* Pn = P(nb);
    double q = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) q += Pn[pk(rr, m)] * vy[12 * nb + m];
    if (r < 12) b[r] = q;
    int idx[4]; double val[4];
    coupling_row(type, sb(h), idx, val);
    w.sync();
    if (act && r < 12) {
      double t = vy[12 * k + r];
#pragma unroll
      for (int a = 0; a < 4; ++a) t -= val[a] * b[idx[a]];
      vy[12 * k + r] = t;
    }
    w.sync();
  }
  SRBD_DEV void bwd_round(int k, int nb, int type, bool act) const {
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    const int rr = r < 12 ? r : 11;
    double t = vy[12 * k + rr];
    if (nb >= 0) {
      int idx[4]; double val[4];
      coupling_row(type, sb(h), idx, val);
#pragma unroll
      for (int a = 0; a < 4; ++a) t -= val[a] * vy[12 * nb + idx[a]];
    }
    if (r < 12) b[r] = t;
    w.sync();
    const double* Pk = P(k);
    double dy = 0.0;
#pragma unroll
    for (int m = 0; m < 12; ++m) dy += Pk[pk(rr, m)] * b[m];
    if (act && r < 12) vy[12 * k + r] = dy;
    w.sync();
  }
  SRBD_DEV void solve_thomas() const {
    for (int t = 1; t < 4; ++t) {
      const int k = h ? 9 - t : t;
      stage(k);
      fwd_round(k, h ? k + 1 : k - 1, h, true);
    }
    { const int k = h ? 5 : 4; stage(k); fwd_round(k, h ? 6 : 3, h, true); }
    { stage(5); fwd_round(5, 4, 0, h == 0); }
    // backward: dy_5 ; then 4|6 ; 3|7 ; 2|8 ; 1|9 ; 0|-
    { stage(5); bwd_round(5, -1, 0, h == 0); }
    for (int t = 1; t <= 5; ++t) {
      const int k = h ? 5 + t : 5 - t;
      const bool act = k <= 9;
      const int kk = act ? k : 9;
      stage(kk);
      // top: t_k - C_{k+1}' dy_{k+1} (LN) ; bottom: t_k - C_k dy_{k-1} (LP)
      bwd_round(kk, h ? kk - 1 : kk + 1, h ? 0 : 1, act);
    }
  }

  // ------------------------------------------------------------------------------------
  // recover dx (into vx), dy_mx (into vy), ds/dz (into OFF_DG / OFF_R2 temporaries returned
  // through registers of the owning lane): one stage
  // ------------------------------------------------------------------------------------
  SRBD_DEV void recover(int k, double& ds_out, double& dz_out) const {
    const double* s = sb(h);
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    const double* dyk = vy + 12 * k;
    if (r < 12) {
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < 12; ++i) q += s[SB_B + 14 * i + r] * dyk[i];
      b[r] = vx[120 + 12 * k + r] - q;
    }
    w.sync();
    if (r < 12) {
      double du;
      if (r == 6 || r == 9) {
        const int i = (r == 9);
        const double m = s[SB_MX + i], iden = s[SB_DER + 4 + i], eps = s[SB_DER + 6 + i];
        const double r4 = -vy[120 + 2 * k + i];
        du = (IP_DELTA * b[r] + m * r4) * iden;
        vy[120 + 2 * k + i] = (m * b[r] - eps * r4) * iden;     // dy_mx
      } else {
        du = psi_apply(psi(k), s + SB_DER, b, r);
      }
      // states: dx = phix (r1x - E dy_k - Abar_{k+1}' dy_{k+1})
      double wx = vx[12 * k + r] - s[SB_E + r] * dyk[r];
      if (k <= 8) {
        const double* A = s + SB_AN;
        const double* dyn = vy + 12 * (k + 1);
        wx -= A[r] * dyn[r];
        if (r >= 6 && r < 9) { for (int i = 0; i < 3; ++i) wx -= A[12 + 3 * i + (r - 6)] * dyn[i]; }
        else if (r >= 9) wx -= A[21 + r - 9] * dyn[r - 6];
      }
      vx[12 * k + r] = s[SB_PHX + r] * wx;
      vx[120 + 12 * k + r] = du;
    }
    w.sync();
    {
      // inequality row j = r of the stage
      const int ft = r >> 3, jj = r & 7;
      const double gb = s[SB_G + 12 * ft + 4 + jj];
      double ga = 0.0; int ca = 3 * ft + 2;
      if (jj < 4) { ga = s[SB_G + 12 * ft + jj]; ca = 3 * ft + (jj >> 1); }
      else if (jj < 6) { ga = s[SB_G + 24 + 2 * ft + jj - 4]; ca = 7 + 3 * ft; }
      const double* du = vx + 120 + 12 * k;
      const double gdu = ga * du[ca] + gb * du[3 * ft + 2];
      const int j = 16 * k + r;
      const double dz = sm[SM::OFF_D + j] * gdu + sm[SM::OFF_DG + j];
      dz_out = dz;
      ds_out = (sm[SM::OFF_R2 + j] - dz) * sm[SM::OFF_WI + j];
    }
  }

  SRBD_DEV double warp_sum(double v) const {
    for (int m = 16; m >= 1; m >>= 1) v += w.shfl_xor(v, m);
    return v;
  }
  SRBD_DEV double warp_min(double v) const {
    for (int m = 16; m >= 1; m >>= 1) v = fmin(v, w.shfl_xor(v, m));
    return v;
  }
  SRBD_DEV static double ratio(double v, double dv) { return dv < 0.0 ? -v / dv : 1.0; }
  SRBD_DEV static double step_len(double mn) { return fmax(fmin(1.0, 0.99 * mn), 1e-12); }

  // ------------------------------------------------------------------------------------
  // the whole call: num_iters Newton iterations for this warp's environment
  // ------------------------------------------------------------------------------------
  SRBD_DEV int run(const PdipmArgs& A, long long e) {
    lane = w.lane(); h = lane >> 4; r = lane & 15;
    gQ = A.in[IP_Q] + e * NNZ_H; gG = A.in[IP_G] + e * NNZ_G; gA = A.in[IP_A] + e * NNZ_A;
    gF = A.in[IP_F] + e * NZ; gH = A.in[IP_H] + e * NINEQ; gB = A.in[IP_B] + e * NEQ;
    oX = A.out[IPO_X] + e * NZ; oS = A.out[IPO_S] + e * NINEQ; oZ = A.out[IPO_Z] + e * NINEQ;
    oY = A.out[IPO_Y] + e * NEQ; oRes = A.out[IPO_RES] + e * 4; oMu = A.out[IPO_MU] + e;
    ws = A.work + e * (long long)PDIPM_WORK_DOUBLES;
    double* wdx = ws; double* wdy = ws + 240; double* wds = ws + 380; double* wdz = ws + 540;

    // ---- initial iterate -> output buffers (updated in place from here on) ---------------
    if (A.cold_start) {
      for (int i = lane; i < NZ; i += 32) oX[i] = 0.0;
      for (int i = lane; i < NINEQ; i += 32) { oS[i] = fmax(ldg(gH + i), 1.0); oZ[i] = 1.0; }
      for (int i = lane; i < NEQ; i += 32) oY[i] = 1.0;
    } else {
      const double* ix = A.in[IP_X] + e * NZ; const double* is = A.in[IP_S] + e * NINEQ;
      const double* iz = A.in[IP_Z] + e * NINEQ; const double* iy = A.in[IP_Y] + e * NEQ;
      for (int i = lane; i < NZ; i += 32) oX[i] = ix[i];
      for (int i = lane; i < NINEQ; i += 32) { oS[i] = is[i]; oZ[i] = iz[i]; }
      for (int i = lane; i < NEQ; i += 32) oY[i] = iy[i];
    }
    w.sync();
    stage_once();

    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* sx = sm + SM::OFF_P;         // iterate scratch (phase A only)
    double* ss = sx + 240; double* sz = sx + 400; double* sy = sx + 560;
    int it = 0;
    double n_rx = 0, n_rs = 0, n_re = 0, mu_new = 0;
    for (; it < A.num_iters; ++it) {
      // ======================= phase A: residuals, weights, Psi, r1~ =====================
      for (int i = lane; i < NZ; i += 32) sx[i] = oX[i];
      for (int i = lane; i < NINEQ; i += 32) { ss[i] = oS[i]; sz[i] = oZ[i]; }
      for (int i = lane; i < NEQ; i += 32) sy[i] = oY[i];
      w.sync();
      double a_sz = 0, a_rs = 0, a_rx = 0, a_re = 0;
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        const double* s = sb(h);
        double* ps = psi(k);
        {   // A1: inequality row j = r
          const int ft = r >> 3, jj = r & 7, j = 16 * k + r;
          const double gb = s[SB_G + 12 * ft + 4 + jj];
          double ga = 0.0; int ca = 3 * ft + 2;
          if (jj < 4) { ga = s[SB_G + 12 * ft + jj]; ca = 3 * ft + (jj >> 1); }
          else if (jj < 6) { ga = s[SB_G + 24 + 2 * ft + jj - 4]; ca = 7 + 3 * ft; }
          const double* u = sx + 120 + 12 * k;
          const double sv = ss[j], zv = sz[j];
          const double rs = ga * u[ca] + gb * u[3 * ft + 2] + sv - ldg(gH + j);
          const double si = 1.0 / sv;
          const double W = si * zv + IP_DELTA;
          const double wi = 1.0 / W;
          const double D = 1.0 / (wi + IP_DELTA);
          const double r2 = -(si * (sv * zv));
          const double Dg = D * (r2 * wi + rs);
          sm[SM::OFF_D + j] = D; sm[SM::OFF_WI + j] = wi; sm[SM::OFF_DG + j] = Dg; sm[SM::OFF_R2 + j] = r2;
          a_sz += sv * zv; a_rs += rs * rs;
          // Psi: pairs (0,1)->fx (2,3)->fy (4,5)->my ; (6,7) -> fz only
          const double pD = w.shfl_xor(D, 1), pga = w.shfl_xor(ga, 1), pgb = w.shfl_xor(gb, 1);
          double contrib, ia = 0.0, tq = 0.0;
          if (jj < 6) {
            const double eps = s[SB_HU + ca] + IP_BETA;
            const double aa = eps + ga * ga * D + pga * pga * pD;
            const double bb = ga * gb * D + pga * pgb * pD;
            const double cr = gb * pga - pgb * ga;
            ia = 1.0 / aa;
            tq = bb * ia;
            contrib = (eps * (gb * gb * D + pgb * pgb * pD) + D * pD * cr * cr) * ia;
          } else {
            contrib = gb * gb * D + pgb * pgb * pD;
          }
          if (jj & 1) contrib = 0.0;
          contrib += w.shfl_xor(contrib, 1);
          contrib += w.shfl_xor(contrib, 2);
          contrib += w.shfl_xor(contrib, 4);
          const double sigma = s[SB_HU + 3 * ft + 2] + IP_BETA + contrib;
          if (jj < 6 && !(jj & 1)) { ps[7 * ft + (jj >> 1)] = ia; ps[7 * ft + 3 + (jj >> 1)] = tq; }
          if (jj == 6) ps[7 * ft + 6] = 1.0 / sigma;
        }
        w.sync();
        if (r < 12) {   // A3: input column c = r ; A4: state column c = r
          const int c = r;
          const double* u = sx + 120 + 12 * k;
          const double* yk = sy + 12 * k;
          double acc = s[SB_HU + c] * u[c] + ldg(gF + 120 + 12 * k + c);
          double gdg = 0.0;
          int ft = -1, row0 = 0, cnt = 0, gofs = 0;
          if (c < 6) {
            ft = c / 3; const int q = c % 3;
            if (q == 0) { row0 = 8 * ft; cnt = 2; gofs = 12 * ft; }
            else if (q == 1) { row0 = 8 * ft + 2; cnt = 2; gofs = 12 * ft + 2; }
            else { row0 = 8 * ft; cnt = 8; gofs = 12 * ft + 4; }
          } else if (c == 7 || c == 10) { ft = (c == 10); row0 = 8 * ft + 4; cnt = 2; gofs = 24 + 2 * ft; }
          for (int i = 0; i < cnt; ++i) {
            const double gv = s[SB_G + gofs + i];
            acc += gv * sz[16 * k + row0 + i];
            gdg += gv * sm[SM::OFF_DG + 16 * k + row0 + i];
          }
#pragma unroll
          for (int i = 0; i < 12; ++i) acc += s[SB_B + 14 * i + c] * yk[i];
          if (c == 6) acc += s[SB_MX] * sy[120 + 2 * k];
          if (c == 9) acc += s[SB_MX + 1] * sy[121 + 2 * k];
          a_rx += acc * acc;
          vx[120 + 12 * k + c] = -acc - gdg;
          // state column
          double ax = s[SB_HX + c] * sx[12 * k + c] + ldg(gF + 12 * k + c) + s[SB_E + c] * yk[c];
          if (k <= 8) {
            const double* An = s + SB_AN;
            const double* yn = sy + 12 * (k + 1);
            ax += An[c] * yn[c];
            if (c >= 6 && c < 9) { for (int i = 0; i < 3; ++i) ax += An[12 + 3 * i + (c - 6)] * yn[i]; }
            else if (c >= 9) ax += An[21 + c - 9] * yn[c - 6];
          }
          a_rx += ax * ax;
          vx[12 * k + c] = -ax;
          // A5: dynamics row r
          double re = s[SB_E + r] * sx[12 * k + r] - ldg(gB + 12 * k + r);
#pragma unroll
          for (int cc = 0; cc < 12; ++cc) re += s[SB_B + 14 * r + cc] * u[cc];
          if (k >= 1) {
            const double* Ap = s + SB_AP;
            const double* xp = sx + 12 * (k - 1);
            re += Ap[r] * xp[r];
            if (r < 3) { for (int j2 = 0; j2 < 3; ++j2) re += Ap[12 + 3 * r + j2] * xp[6 + j2]; }
            else if (r < 6) re += Ap[21 + r - 3] * xp[r + 6];
          }
          a_re += re * re;
          vy[12 * k + r] = re;
        } else if (r < 14) {   // Mx rows
          const int i = r - 12;
          const double re = s[SB_MX + i] * sx[120 + 12 * k + 6 + 3 * i] - ldg(gB + 120 + 2 * k + i);
          a_re += re * re;
          vy[120 + 2 * k + i] = re;
        }
        w.sync();
      }
      a_sz = warp_sum(a_sz); a_rs = warp_sum(a_rs); a_rx = warp_sum(a_rx); a_re = warp_sum(a_re);
      const double mu = a_sz / (double)NINEQ;
      n_rx = sqrt(a_rx); n_rs = sqrt(a_rs); n_re = sqrt(a_re);

      // ======================= phase B: rho (affine) and S_kk ============================
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        build_rho(k);
        form_S(k);
      }
      factor();
      SRBD_TRACE("P", sm + SM::OFF_P, 780);
      SRBD_TRACE("rho_aff", vy, 140);
      solve_thomas();
      // ======================= affine recovery + step lengths ============================
      // min over ALL 160 candidates (1.0 where the direction is non-negative): may exceed 1
      double mn_s = 1e300, mn_z = 1e300;
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        double ds, dz;
        recover(k, ds, dz);
        const int j = 16 * k + r;
        wds[j] = ds; wdz[j] = dz;
        mn_s = fmin(mn_s, ratio(oS[j], ds));
        mn_z = fmin(mn_z, ratio(oZ[j], dz));
        w.sync();
      }
      for (int i = lane; i < NZ; i += 32) wdx[i] = vx[i];
      for (int i = lane; i < NEQ; i += 32) wdy[i] = vy[i];
      SRBD_TRACE("dx_aff", vx, 240);
      SRBD_TRACE("dy_aff", vy, 140);
      const double a_pri_aff = step_len(warp_min(mn_s));
      const double a_dual_aff = step_len(warp_min(mn_z));
      // mu_aff, sigma, corrector rhs
      double a_mu = 0.0;
      for (int t = 0; t < 5; ++t) {
        const int j = 16 * (h ? 9 - t : t) + r;
        a_mu += (oS[j] + a_pri_aff * wds[j]) * (oZ[j] + a_dual_aff * wdz[j]);
      }
      const double mu_aff = warp_sum(a_mu) / (double)NINEQ;
      const double sg = mu_aff / mu;
      const double sigma = sg * sg * sg;
      const double smu = sigma * mu;
      for (int t = 0; t < 5; ++t) {
        const int j = 16 * (h ? 9 - t : t) + r;
        const double sv = oS[j], zv = oZ[j];
        const double r2 = -((1.0 / sv) * (sv * zv + wds[j] * wdz[j] - smu));
        sm[SM::OFF_R2 + j] = r2;
        sm[SM::OFF_DG + j] = sm[SM::OFF_D + j] * (r2 * sm[SM::OFF_WI + j]);
      }
      for (int i = lane; i < NEQ; i += 32) vy[i] = 0.0;
      for (int i = lane; i < 120; i += 32) vx[i] = 0.0;
      w.sync();
      // corrector r1~ (inputs only) and rho
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        const double* s = sb(h);
        if (r < 12) {
          const int c = r;
          double gdg = 0.0;
          int ft = -1, row0 = 0, cnt = 0, gofs = 0;
          if (c < 6) {
            ft = c / 3; const int q = c % 3;
            if (q == 0) { row0 = 8 * ft; cnt = 2; gofs = 12 * ft; }
            else if (q == 1) { row0 = 8 * ft + 2; cnt = 2; gofs = 12 * ft + 2; }
            else { row0 = 8 * ft; cnt = 8; gofs = 12 * ft + 4; }
          } else if (c == 7 || c == 10) { ft = (c == 10); row0 = 8 * ft + 4; cnt = 2; gofs = 24 + 2 * ft; }
          for (int i = 0; i < cnt; ++i) gdg += s[SB_G + gofs + i] * sm[SM::OFF_DG + 16 * k + row0 + i];
          vx[120 + 12 * k + c] = -gdg;
        }
        w.sync();
        build_rho(k);
      }
      solve_thomas();
      // ======================= corrector recovery, combined direction =====================
      mn_s = 1e300; mn_z = 1e300;
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        double ds, dz;
        recover(k, ds, dz);
        const int j = 16 * k + r;
        ds += wds[j]; dz += wdz[j];
        w.sync();   // all lanes done with DG/R2 of this stage before they are overwritten
        sm[SM::OFF_DG + j] = ds; sm[SM::OFF_R2 + j] = dz;
        mn_s = fmin(mn_s, ratio(oS[j], ds));
        mn_z = fmin(mn_z, ratio(oZ[j], dz));
      }
      w.sync();
      const double a_pri = step_len(warp_min(mn_s));
      const double a_dual = step_len(warp_min(mn_z));
#ifdef SRBD_HOST_EMUL
      { double dbg[8] = {a_pri_aff, a_dual_aff, sigma, mu, a_pri, a_dual, mu_aff, 0.0}; SRBD_TRACE("scalars", dbg, 8); }
      SRBD_TRACE("dx_cor", vx, 240);
      SRBD_TRACE("dy_cor", vy, 140);
      SRBD_TRACE("ds_tot", sm + SM::OFF_DG, 160);
      SRBD_TRACE("dz_tot", sm + SM::OFF_R2, 160);
#endif
      // ======================= update ====================================================
      double a_new = 0.0;
      for (int i = lane; i < NZ; i += 32) oX[i] = oX[i] + a_pri * (wdx[i] + vx[i]);
      for (int i = lane; i < NEQ; i += 32) oY[i] = oY[i] + a_dual * (wdy[i] + vy[i]);
      for (int i = lane; i < NINEQ; i += 32) {
        const double sn = fmax(oS[i] + a_pri * sm[SM::OFF_DG + i], 1e-8);
        const double zn = fmax(oZ[i] + a_dual * sm[SM::OFF_R2 + i], 1e-8);
        oS[i] = sn; oZ[i] = zn;
        a_new += sn * zn;
      }
      mu_new = warp_sum(a_new) / (double)NINEQ;
      w.sync();
      if (A.tol > 0.0 && n_rx < A.tol && n_rs < A.tol && n_re < A.tol && mu_new < A.tol) { ++it; break; }
    }
    if (lane == 0) {
      oRes[0] = n_rx; oRes[1] = n_rs; oRes[2] = n_re; oRes[3] = mu_new;
      oMu[0] = mu_new;
      if (A.iters_done) A.iters_done[e] = it;
    }
    return it;
  }
};

}  // namespace srbd
