// PDIPM Newton iterations, TWO environments per warp.
//
// Same mathematics, shared-memory layout (FastSmem, one copy per environment) and barrier
// structure as srbd_pdipm_fast.h; what changes is the thread mapping.  The fast kernel is bound
// by shared-memory bandwidth, and most of that traffic is broadcasts: a 12-vector (pivot row,
// right-hand side, a row of Bbar) delivered to the 12 lanes that work on one 12x12 block costs
// the same number of wavefronts whether one or eight blocks are served by the instruction.
// Here a warp serves FOUR blocks per instruction:
//
//   lane = 8 g + l,  g = 2 * (environment slot) + (half: top / bottom of the horizon),
//   lane l < 6 of a group owns rows l and l + 6 of the group's current stage block (12-row data)
//   and all eight lanes own inequality rows l (left foot) and l + 8 (right foot).
//
// Every broadcast load feeds two rows (twice the FMAs per load), the per-pivot overhead
// (reciprocal, barrier, address arithmetic) is shared by four blocks instead of two, and the two
// environments of a warp give the scheduler the same instruction-level parallelism two warps gave
// before.  7 warps x 2 environments per SM; up to 255 registers per thread.
//
// Compiled for the GPU (srbd_pdipm.cu) and, for tests only, for the host under the warp emulator.
#pragma once

#include "srbd_pdipm_fast.h"

namespace srbd {

struct PdipmPair {
  using SM = FastSmem;
  enum : int { FB_ZERO = FB_B + 12 };
  struct Row {                    // what a lane keeps per 12-row index it owns
    int rr;
    int prow[12];
    double c0[4], c1[4];
    int i0[3], i1[3];
    double phx;
    int psk, psi_i, psf, doff, moff;
    int g_row0, g_cnt, g_ofs;
    int bs0, bs1;
  };
  struct Ineq { int goa, gob, ca, cz; };

  Warp w;
  double* sm;                     // shared memory of THIS lane's environment
  const StageSlot* slots;
  const PdipmArgs* pa;
  long long e;
  int lane, l16, l, h;
  bool own;                       // l < 6: the lane really owns its two 12-row indices
  bool wr;                        // the lane's environment exists: it may write global memory (the idle
                                  // half of the last warp of an odd batch shadows env B-1 read-only)
  Row R[2];
  Ineq Q[2];

  SRBD_DEV const double* gin(int i, int nnz) const { return pa->in[i] + e * nnz; }
  SRBD_DEV double* gout(int i, int nnz) const { return pa->out[i] + e * nnz; }
  SRBD_DEV double* gws() const { return pa->work + e * (long long)PDIPM_WORK_DOUBLES; }
  SRBD_DEV double* P(int k) const { return sm + SM::OFF_P + 78 * k; }
  SRBD_DEV double* fb() const { return sm + SM::OFF_FB; }
  SRBD_DEV double* bc() const { return sm + SM::OFF_BC + 24 * h; }
  SRBD_DEV double* tr() const { return sm + SM::OFF_TR + 32 * h; }
  SRBD_DEV double* pr(int hh) const { return sm + SM::OFF_PR + 72 * hh; }
  SRBD_DEV double* psi(int k) const { return sm + SM::OFF_PSI + 14 * k; }

  // ---- reductions over the 16 lanes of one environment -------------------------------------
  SRBD_DEV double env_sum(double v) const {
    for (int m = 8; m >= 1; m >>= 1) v += w.shfl_xor(v, m);
    return v;
  }
  SRBD_DEV double env_min(double v) const {
    for (int m = 8; m >= 1; m >>= 1) { const double o = w.shfl_xor(v, m); v = o < v ? o : v; }
    return v;
  }
  SRBD_DEV static double ratio(double v, double dv) { return dv < 0.0 ? -v * rcp(dv) : 1.0; }
  SRBD_DEV static double step_len(double mn) { return fmax(fmin(1.0, 0.99 * mn), 1e-12); }
  SRBD_DEV static double dot12(const double a[12], const double b[12]) {
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int c = 0; c < 12; c += 2) { q0 = fma(a[c], b[c], q0); q1 = fma(a[c + 1], b[c + 1], q1); }
    return q0 + q1;
  }

  // ------------------------------------------------------------------------------------
  // set-up
  // ------------------------------------------------------------------------------------
  SRBD_DEV void gather_fused() {
    double* f = fb();
    double* V = sm + SM::OFF_P;
    double* Wr = V + 232;
    qp_value_table(pa->qp_in, e, l16, 16, w, V, Wr);
    for (int i = l16; i < 168; i += 16) { const int rw = i / 14, c = i - 14 * rw; f[FB_B + i] = c < 12 ? V[V_NBD + 12 * rw + c] : 0.0; }
    for (int t = l16; t < 12; t += 16) { f[FB_AB + t] = V[V_NEG1]; f[FB_HX + t] = V[V_Q + t]; f[FB_HU + t] = V[V_R + t]; }
    for (int t = l16; t < 9; t += 16) f[FB_AB + 12 + t] = V[V_NADTR + t];
    for (int t = l16; t < 3; t += 16) f[FB_AB + 21 + t] = V[V_NADT];
    for (int t = l16; t < 28; t += 16) {
      const int q = t < 24 ? t % 12 : t - 24;
      int v;
      if (t < 24) v = (q == 0 || q == 2 || q == 10) ? V_NEG1 : (q == 1 || q == 3 || q == 11) ? V_ONE : q < 8 ? V_NEGMU : (q == 8 ? V_NEGLT : V_NEGLH);
      else v = (q & 1) ? V_ONE : V_NEG1;
      f[FB_G + t] = V[v];
    }
    double* ws = gws();
    {
      const double* xr = pa->qp_in[IN_XREF] + e * 120;
      const double* xl = pa->qp_in[IN_X] ? pa->qp_in[IN_X] + e * 120 : nullptr;
      for (int i = l16; i < 120; i += 16) {
        const double xv = xl ? xl[i] : 1.0, q = V[V_Q + i % 12];
        if (wr) ws[FW_FX + i] = q * (xv - xr[i]) - q * xv;
      }
    }
    for (int t = l16; t < 24; t += 16) if (wr) ws[FW_CD + t] = V[V_CD + t];
    for (int t = l16; t < 20; t += 16) if (wr) ws[FW_FC + t] = V[V_FC + t];
    w.sync();
  }
  SRBD_DEV double h_of(int k, int j) const {
    if (!pa->fused) return ldg(gin(IP_H, NINEQ) + j);
    return (j & 7) == 7 ? gws()[FW_FC + 2 * k + ((j >> 3) & 1)] : 0.0;
  }

  SRBD_DEV void setup_row(Row& X, int q) {
    double* f = fb();
    X.rr = q;
#pragma unroll
    for (int m = 0; m < 12; ++m) X.prow[m] = pk(q, m);
    X.bs0 = 3 + q % 3; X.bs1 = 9 + q % 3;
    X.phx = f[FB_PHX + q];
    const double* A = f + FB_AB; const double* ph = f + FB_PHX;
#pragma unroll
    for (int a = 0; a < 4; ++a) { X.c0[a] = 0.0; X.c1[a] = 0.0; }
#pragma unroll
    for (int a = 0; a < 3; ++a) { X.i0[a] = 6 + a; X.i1[a] = a; }   // zero-coefficient slots share an address (srbd_pdipm_fast.h)
    X.c0[0] = A[q] * ph[q];
    X.c1[0] = A[q] * ph[q];
    if (q < 3) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { X.i0[j] = 6 + j; X.c0[1 + j] = A[12 + 3 * q + j] * ph[6 + j]; }
    } else if (q < 6) { X.i0[0] = q + 6; X.c0[1] = A[21 + q - 3] * ph[q + 6]; }
    else if (q < 9) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { X.i1[i] = i; X.c1[1 + i] = A[12 + 3 * i + (q - 6)] * ph[q]; }
    } else { X.i1[0] = q - 6; X.c1[1] = A[21 + q - 9] * ph[q]; }
    double* m0 = f + FB_CM; double* m1 = f + FB_CM + 36;
    m0[q] = X.c0[0]; m1[q] = X.c1[0];
    if (q < 3) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { m0[12 + 3 * q + j] = X.c0[1 + j]; m1[12 + 3 * q + j] = 0.0; }
    } else if (q < 6) { m0[21 + q - 3] = X.c0[1]; m1[21 + q - 3] = 0.0; }
    else if (q < 9) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { m0[24 + 3 * (q - 6) + i] = 0.0; m1[24 + 3 * (q - 6) + i] = X.c1[1 + i]; }
    } else { m0[33 + q - 9] = 0.0; m1[33 + q - 9] = X.c1[1]; }
    for (int j = 0; j <= q; ++j) {
      double acc = X.c0[0] * PdipmWarp<true>::abar(A, j, q);
#pragma unroll
      for (int a = 0; a < 3; ++a) acc += X.c0[1 + a] * PdipmWarp<true>::abar(A, j, X.i0[a]);
      f[FB_SX + pk(q, j)] = acc;
    }
    const int c = q;
    X.psf = (c < 6) ? c / 3 : (c >= 9 ? 1 : 0);
    X.doff = FB_ZERO; X.moff = FB_ZERO; X.psi_i = 0;
    if (c < 6) {
      X.psi_i = c % 3;
      if (X.psi_i == 2) { X.psk = 1; X.psi_i = 0; } else X.psk = 0;
    } else if (c == 7 || c == 10) { X.psk = 0; X.psi_i = 2; }
    else if (c == 8 || c == 11) { X.psk = 2; X.psi_i = (c == 11); X.doff = FB_DER + X.psi_i; }
    else { X.psk = 3; X.psi_i = (c == 9); X.doff = FB_DER + 2 + X.psi_i; X.moff = FB_DER + 4 + X.psi_i; }
    X.g_row0 = 0; X.g_cnt = 0; X.g_ofs = 0;
    if (c < 6) {
      const int ft = c / 3, qq = c % 3;
      if (qq == 0) { X.g_row0 = 8 * ft; X.g_cnt = 2; X.g_ofs = 12 * ft; }
      else if (qq == 1) { X.g_row0 = 8 * ft + 2; X.g_cnt = 2; X.g_ofs = 12 * ft + 2; }
      else { X.g_row0 = 8 * ft; X.g_cnt = 8; X.g_ofs = 12 * ft + 4; }
    } else if (c == 7 || c == 10) { const int ft = (c == 10); X.g_row0 = 8 * ft + 4; X.g_cnt = 2; X.g_ofs = 24 + 2 * ft; }
  }

  SRBD_DEV int setup() {
    double* f = fb();
    int bad = 0;
    if (pa->fused) gather_fused();
    else {
      const double* gQ = gin(IP_Q, NNZ_H); const double* gG = gin(IP_G, NNZ_G); const double* gA = gin(IP_A, NNZ_A);
      for (int i = l16; i < FB_NSLOT; i += 16) {
        const StageSlot sl = slots[i];
        double v = 0.0;
        switch (sl.kind) {
          case ST_AU: v = ldg(gA + 336 + 86 + sl.rel); break;
          case ST_E: v = ldg(gA + a_xcol_start(1, sl.rel)); break;
          case ST_AXPREV: v = ldg(gA + sl.rel); break;
          case ST_G: v = ldg(gG + 28 + sl.rel); break;
          case ST_Q: v = ldg(gQ + 12 + sl.rel); break;
          default: break;
        }
        if (i < FB_RAW) f[i] = v;
        else if (v != 1.0) bad = 1;
      }
      for (int m = 8; m >= 1; m >>= 1) bad |= w.shfl_i(bad, lane ^ m);
      w.sync();
    }
    if (l16 < 12) f[FB_PHX + l16] = rcp(f[FB_HX + l16] + IP_BETA);
    else if (l16 < 14) f[FB_DER + l16 - 12] = rcp(f[FB_HU + 8 + 3 * (l16 - 12)] + IP_BETA);   // mz
    else {                                                                                  // mx
      const int i = l16 - 14;
      const double eps = f[FB_HU + 6 + 3 * i] + IP_BETA;
      const double iden = rcp(1.0 + IP_DELTA * eps);
      f[FB_DER + 2 + i] = IP_DELTA * iden;
      f[FB_DER + 4 + i] = iden;
      f[FB_DER + 6 + i] = eps;
    }
    w.sync();
    const int la = l < 6 ? l : 5;
    setup_row(R[0], la);
    setup_row(R[1], la + 6);
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // inequality rows l (left foot) and l + 8 (right foot)
      const int ft = q, jj = l;
      Q[q].cz = 3 * ft + 2;
      Q[q].gob = FB_G + 12 * ft + 4 + jj;
      Q[q].goa = FB_ZERO; Q[q].ca = Q[q].cz;
      if (jj < 4) { Q[q].goa = FB_G + 12 * ft + jj; Q[q].ca = 3 * ft + (jj >> 1); }
      else if (jj < 6) { Q[q].goa = FB_G + 24 + 2 * ft + jj - 4; Q[q].ca = 7 + 3 * ft; }
    }
    w.sync();
    return bad;
  }

  // ------------------------------------------------------------------------------------
  SRBD_DEV double psi_apply(const Row& X, const double* ps, const double* v) const {
    const double* p = ps + 7 * X.psf;
    const int fx = 3 * X.psf;
    const double dot = p[3] * v[fx] + p[4] * v[fx + 1] + p[5] * v[7 + fx] - v[fx + 2];
    const double dg = X.psk == 0 ? p[X.psi_i] : fb()[X.doff];
    const double q = X.psk == 0 ? p[3 + X.psi_i] : (X.psk == 1 ? -1.0 : 0.0);
    return dg * v[X.rr] + q * (p[6] * dot);
  }
  SRBD_DEV void gather_col(const Row& X, const double* va, const double* vb, double& sa, double& sb_) const {
    const double* g = fb() + FB_G + X.g_ofs;
    double xa = 0.0, xb = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool on = i < X.g_cnt;
      const double gv = on ? g[i] : 0.0;
      const double a = on ? va[X.g_row0 + i] : 0.0;
      xa = fma(gv, a, xa);
      if (vb) { const double b = on ? vb[X.g_row0 + i] : 0.0; xb = fma(gv, b, xb); }
    }
    sa = xa; sb_ = xb;
  }
  SRBD_DEV double bcol_dot(const Row& X, const double* v) const {
    const double* B = fb() + FB_B + X.rr;
    const double q0 = B[0] * v[0] + B[14] * v[1] + B[28] * v[2] + B[14 * X.bs0] * v[X.bs0];
    const double q1 = B[84] * v[6] + B[98] * v[7] + B[112] * v[8] + B[14 * X.bs1] * v[X.bs1];
    return q0 + q1;
  }
  SRBD_DEV double brow_dot(const Row& X, const double v[12]) const {
    double Brow[12];
    ld12(fb() + FB_B + 14 * X.rr, Brow);
    return dot12(Brow, v);
  }

  // ------------------------------------------------------------------------------------
  SRBD_DEV void build_rho(int k, bool cor) const {
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const Row& X = R[q];
      double v = psi_apply(X, psi(k), vx + 120 + 12 * k);
      if (!cor) v = fma(fb()[X.moff], -vy[120 + 2 * k + X.psi_i], v);
      b[X.rr] = v;
    }
    w.sync();
    {
      double bv[12];
      ld12(b, bv);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const Row& X = R[q];
        double rho = brow_dot(X, bv);
        if (!cor) {
          rho += X.phx * vx[12 * k + X.rr] + vy[12 * k + X.rr];
          if (k >= 1) {
            const double* xp = vx + 12 * (k - 1);
            rho += X.c0[0] * xp[X.rr] + X.c0[1] * xp[X.i0[0]] + X.c0[2] * xp[X.i0[1]] + X.c0[3] * xp[X.i0[2]];
          }
        }
        if (own) vy[12 * k + X.rr] = rho;
      }
    }
    w.sync();
  }

  // rows R[0].rr, R[1].rr of S_kk
  SRBD_DEV void form_rows(int k, double a[2][12]) const {
    const double* f = fb();
    const double* ps = psi(k);
    const double* der = f + FB_DER;
    double* b = bc();
    double psv[14];
    {
      const dbl2* q = reinterpret_cast<const dbl2*>(ps);
#pragma unroll
      for (int i = 0; i < 7; ++i) { const dbl2 t = q[i]; psv[2 * i] = t.x; psv[2 * i + 1] = t.y; }
    }
    double wv[2][12], sL[2], sR[2];
    const double fk = (k == 0) ? 0.0 : 1.0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const Row& X = R[q];
      double Brow[12];
      ld12(f + FB_B + 14 * X.rr, Brow);
      const double pL = psv[3] * Brow[0] + psv[4] * Brow[1] + psv[5] * Brow[7] - Brow[2];
      const double pR = psv[10] * Brow[3] + psv[11] * Brow[4] + psv[12] * Brow[10] - Brow[5];
      b[X.rr] = pL; b[12 + X.rr] = pR;
      wv[q][0] = psv[0] * Brow[0]; wv[q][1] = psv[1] * Brow[1]; wv[q][2] = 0.0;
      wv[q][3] = psv[7] * Brow[3]; wv[q][4] = psv[8] * Brow[4]; wv[q][5] = 0.0;
      wv[q][6] = der[2] * Brow[6]; wv[q][7] = psv[2] * Brow[7]; wv[q][8] = der[0] * Brow[8];
      wv[q][9] = der[3] * Brow[9]; wv[q][10] = psv[9] * Brow[10]; wv[q][11] = der[1] * Brow[11];
      sL[q] = psv[6] * pL; sR[q] = psv[13] * pR;
      const double dg = X.phx + IP_DELTA;
#pragma unroll
      for (int j = 0; j < 12; ++j) a[q][j] = fma(fk, f[FB_SX + X.prow[j]], (X.rr == j) ? dg : 0.0);
    }
    w.sync();
    {
      double pl[12];
      ld12(b, pl);
#pragma unroll
      for (int j = 0; j < 12; ++j) { a[0][j] = fma(sL[0], pl[j], a[0][j]); a[1][j] = fma(sL[1], pl[j], a[1][j]); }
      ld12(b + 12, pl);
#pragma unroll
      for (int j = 0; j < 12; ++j) { a[0][j] = fma(sR[0], pl[j], a[0][j]); a[1][j] = fma(sR[1], pl[j], a[1][j]); }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const double* Bj = f + FB_B + 14 * j;
      if (j < 3 || (j >= 6 && j < 9)) {
        double bj[12];
        ld12(Bj, bj);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          double q0 = 0.0, q1 = 0.0;
#pragma unroll
          for (int c = 0; c < 12; c += 2) {
            if (c != 2) q0 = fma(wv[q][c], bj[c], q0);
            if (c + 1 != 5) q1 = fma(wv[q][c + 1], bj[c + 1], q1);
          }
          a[q][j] += q0 + q1;
        }
      } else {
        const int c1_ = (j < 6) ? j - 3 : j - 9, c2_ = c1_ + 3;
        const double b1 = Bj[c1_], b2 = Bj[c2_];
#pragma unroll
        for (int q = 0; q < 2; ++q) a[q][j] = fma(wv[q][c1_], b1, fma(wv[q][c2_], b2, a[q][j]));
      }
    }
  }

  // n -= M P M' for one row; own = row of P, hand = handed-over rows
  SRBD_DEV void couple(const Row& X, double n[12], const double own_[12], const double* hand, int type) const {
    const double cv0 = type ? X.c1[0] : X.c0[0], cv1 = type ? X.c1[1] : X.c0[1];
    const double cv2 = type ? X.c1[2] : X.c0[2], cv3 = type ? X.c1[3] : X.c0[3];
    const int base = type ? 0 : 6;
    int s1 = (type ? X.i1[0] : X.i0[0]) - base, s2 = (type ? X.i1[1] : X.i0[1]) - base, s3 = (type ? X.i1[2] : X.i0[2]) - base;
    s1 = s1 < 0 ? 0 : (s1 > 5 ? 5 : s1); s2 = s2 < 0 ? 0 : (s2 > 5 ? 5 : s2); s3 = s3 < 0 ? 0 : (s3 > 5 ? 5 : s3);
    double U[12];
    {
      double t1[12];
      ld12(hand + 12 * s1, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fma(cv1, t1[m], cv0 * own_[m]);
      ld12(hand + 12 * s2, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fma(cv2, t1[m], U[m]);
      ld12(hand + 12 * s3, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fma(cv3, t1[m], U[m]);
    }
    const double* cm = fb() + FB_CM + 36 * type;
#pragma unroll
    for (int cc = 0; cc < 12; ++cc) {
      double o = U[cc] * cm[cc];
      if (cc < 3) o += U[6] * cm[12 + 3 * cc] + U[7] * cm[12 + 3 * cc + 1] + U[8] * cm[12 + 3 * cc + 2];
      else if (cc < 6) o += U[cc + 6] * cm[21 + cc - 3];
      else if (cc < 9) o += U[0] * cm[24 + 3 * (cc - 6)] + U[1] * cm[24 + 3 * (cc - 6) + 1] + U[2] * cm[24 + 3 * (cc - 6) + 2];
      else o += U[cc - 6] * cm[33 + cc - 9];
      n[cc] -= o;
    }
  }

  // symmetric sweep (see srbd_pdipm_fast.h); each lane holds two rows, the pivot row is loaded once
  SRBD_DEV void sweep_invert(double a[2][12]) const {
    double* b = bc();
    w.sync();
#pragma unroll
    for (int p = 0; p < 12; ++p) {
      double* bp = b + 12 * (p & 1);
      bp[R[0].rr] = a[0][p];
      bp[R[1].rr] = a[1][p];
      w.sync();
      double bv[12];
      ld12(bp, bv);
      const double di = rcp(bv[p]);
      {   // the row that can be the pivot row: index p < 6 lives in R[0], p >= 6 in R[1]
        const int q = p < 6 ? 0 : 1;
        const bool piv = (R[q].rr == p);
        const double t = a[q][p] * di;
        if (piv) {
#pragma unroll
          for (int c = 0; c < 12; ++c) if (c != p) a[q][c] = bv[c] * di;
        } else {
#pragma unroll
          for (int c = 0; c < 12; ++c) if (c != p) a[q][c] = fma(-t, bv[c], a[q][c]);
        }
        a[q][p] = piv ? -di : t;
      }
      {   // the other row is never the pivot row at this step
        const int q = p < 6 ? 1 : 0;
        const double t = a[q][p] * di;
#pragma unroll
        for (int c = 0; c < 12; ++c) if (c != p) a[q][c] = fma(-t, bv[c], a[q][c]);
        a[q][p] = t;
      }
    }
#pragma unroll
    for (int c = 0; c < 12; ++c) { a[0][c] = -a[0][c]; a[1][c] = -a[1][c]; }
  }

  SRBD_DEV void factor() {
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      const int k = (t == 5) ? 5 : (h ? 9 - t : t);
      const bool act = ((t < 4) || (h == 0)) && own;
      double n[2][12];
      form_rows(k, n);
      if (t > 0) {
        const double* Pn = P((t == 5) ? 4 : (h ? k + 1 : k - 1));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          double o[12];
#pragma unroll
          for (int c = 0; c < 12; ++c) o[c] = Pn[R[q].prow[c]];
          couple(R[q], n[q], o, pr(h), h);
        }
      }
      if (t == 5) {
        const double* P6 = P(6);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          double o[12];
#pragma unroll
          for (int c = 0; c < 12; ++c) o[c] = P6[R[q].prow[c]];
          couple(R[q], n[q], o, pr(1), 1);
        }
      }
      sweep_invert(n);
      if (act) {
        double* Pk = P(k);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int c = 0; c < 12; ++c) if (c <= R[q].rr) Pk[R[q].prow[c]] = n[q][c];
        }
        // hand-over window: the top half leaves rows 6..11 (its R[1]), the bottom half rows 0..5 (R[0])
        if (h) st12(pr(1) + 12 * R[0].rr, n[0]);
        else st12(pr(0) + 12 * (R[1].rr - 6), n[1]);
      }
      w.sync();
    }
  }

  SRBD_DEV double prow_dot(const Row& X, const double* Pm, const double v[12]) const {
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
    for (int m = 0; m < 12; m += 3) {
      q0 = fma(Pm[X.prow[m]], v[m], q0); q1 = fma(Pm[X.prow[m + 1]], v[m + 1], q1); q2 = fma(Pm[X.prow[m + 2]], v[m + 2], q2);
    }
    return q0 + q1 + q2;
  }
  SRBD_DEV void solve_thomas() {
    double* vy = sm + SM::OFF_VY;
    double* b = bc();
#pragma unroll 1
    for (int t = 1; t < 6; ++t) {
      int k, nb;
      if (t < 4) { k = h ? 9 - t : t; nb = h ? k + 1 : k - 1; }
      else if (t == 4) { k = h ? 5 : 4; nb = h ? 6 : 3; }
      else { k = 5; nb = 4; }
      const bool act = ((t < 5) || (h == 0)) && own;
      double tv[12];
      ld12(vy + 12 * nb, tv);
#pragma unroll
      for (int q = 0; q < 2; ++q) b[R[q].rr] = prow_dot(R[q], P(nb), tv);
      w.sync();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const Row& X = R[q];
        const double cv0 = h ? X.c1[0] : X.c0[0], cv1 = h ? X.c1[1] : X.c0[1], cv2 = h ? X.c1[2] : X.c0[2], cv3 = h ? X.c1[3] : X.c0[3];
        const int j1 = h ? X.i1[0] : X.i0[0], j2 = h ? X.i1[1] : X.i0[1], j3 = h ? X.i1[2] : X.i0[2];
        const double d = cv0 * b[X.rr] + cv1 * b[j1] + cv2 * b[j2] + cv3 * b[j3];
        if (act) vy[12 * k + X.rr] -= d;
      }
      w.sync();
    }
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      int k = (t == 0) ? 5 : (h ? 5 + t : 5 - t);
      const bool act = ((t == 0) ? (h == 0) : (k <= 9)) && own;
      if (k > 9) k = 9;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const Row& X = R[q];
        double tk = vy[12 * k + X.rr];
        if (t > 0) {
          const double* yn = vy + 12 * (h ? k - 1 : k + 1);
          const double cv0 = h ? X.c0[0] : X.c1[0], cv1 = h ? X.c0[1] : X.c1[1], cv2 = h ? X.c0[2] : X.c1[2], cv3 = h ? X.c0[3] : X.c1[3];
          const int j1 = h ? X.i0[0] : X.i1[0], j2 = h ? X.i0[1] : X.i1[1], j3 = h ? X.i0[2] : X.i1[2];
          tk -= cv0 * yn[X.rr] + cv1 * yn[j1] + cv2 * yn[j2] + cv3 * yn[j3];
        }
        b[X.rr] = tk;
      }
      w.sync();
      double bv[12];
      ld12(b, bv);
      double dy[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) dy[q] = prow_dot(R[q], P(k), bv);
      if (act) { vy[12 * k + R[0].rr] = dy[0]; vy[12 * k + R[1].rr] = dy[1]; }
      w.sync();
    }
  }

  // ======================= phase A: residuals, weights, Psi, r1~ ===========================
  SRBD_DEV void phase_a(double& mu, double& n_rx, double& n_rs, double& n_re) const {
    const double* f = fb();
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* Dsm = sm + SM::OFF_D;
    double* sx = sm + SM::OFF_P;
    double* sy = sx + 240; double* xs = sx + 380;
    double* ws = gws();
    const double* oS = gout(IPO_S, NINEQ); const double* oZ = gout(IPO_Z, NINEQ);
    const bool fused = pa->fused != 0;
    const double* gF = gin(IP_F, NZ); const double* gB = gin(IP_B, NEQ);
    {
      const double* oX = gout(IPO_X, NZ); const double* oY = gout(IPO_Y, NEQ);
      double xv[15], yv[9];
#pragma unroll
      for (int m = 0; m < 15; ++m) xv[m] = oX[l16 + 16 * m];
#pragma unroll
      for (int m = 0; m < 9; ++m) { const int i = l16 + 16 * m; yv[m] = i < NEQ ? oY[i] : 0.0; }
#pragma unroll
      for (int m = 0; m < 15; ++m) {
        const int i = l16 + 16 * m;
        sx[i] = xv[m];
        if (i < 120) xs[i] = xv[m] * (f[FB_HX + i % 12] + IP_BETA);
      }
#pragma unroll
      for (int m = 0; m < 9; ++m) { const int i = l16 + 16 * m; if (i < NEQ) sy[i] = yv[m]; }
    }
    w.sync();
    const double valid = own ? 1.0 : 0.0;
    double a_sz = 0, a_rs = 0, a_rx = 0, a_re = 0;
#pragma unroll 1
    for (int t = 0; t < 5; ++t) {
      const int k = h ? 9 - t : t;
      const double* u = sx + 120 + 12 * k;
      double* ps = psi(k);
      double* trd = tr();
#pragma unroll
      for (int q = 0; q < 2; ++q) {   // inequality rows l (left foot) and 8 + l (right foot)
        const int jr = 8 * q + l, j = 16 * k + jr;
        const double sv = oS[j], zv = oZ[j], hv = h_of(k, j);
        const double ga = f[Q[q].goa], gb = f[Q[q].gob];
        const double pga = w.shfl_xor(ga, 1), pgb = w.shfl_xor(gb, 1);
        const double rs = ga * u[Q[q].ca] + gb * u[Q[q].cz] + sv - hv;
        const double si = rcp(sv);
        const double W = fma(si, zv, IP_DELTA);
        const double qq = rcp(fma(IP_DELTA, W, 1.0));
        const double D = W * qq;
        const double Dg = fma(D, rs, -(qq * zv));
        if (wr) { ws[FW_S + j] = fma(IP_DELTA, Dg, -rs); ws[FW_Z + j] = Dg; ws[FW_Q + j] = qq; }
        Dsm[j] = D;
        trd[jr] = Dg; trd[16 + jr] = zv;
        a_sz = fma(sv, zv, a_sz); a_rs = fma(rs, rs, a_rs);
        const double pD = w.shfl_xor(D, 1);
        const double gg = gb * gb * D + pgb * pgb * pD;
        double contrib = gg, ia = 0.0, tq = 0.0;
        if (l < 6) {
          const double eps_ca = f[FB_HU + Q[q].ca] + IP_BETA;
          const double cr = gb * pga - pgb * ga;
          const double aa = eps_ca + ga * ga * D + pga * pga * pD;
          const double bb = ga * gb * D + pga * pgb * pD;
          ia = rcp(aa);
          tq = bb * ia;
          contrib = (eps_ca * gg + D * pD * cr * cr) * ia;
        }
        if (l & 1) contrib = 0.0;
        contrib += w.shfl_xor(contrib, 1);
        contrib += w.shfl_xor(contrib, 2);
        contrib += w.shfl_xor(contrib, 4);
        if (l < 6 && !(l & 1)) { ps[7 * q + (l >> 1)] = ia; ps[7 * q + 3 + (l >> 1)] = tq; }
        if (l == 6) ps[7 * q + 6] = rcp(f[FB_HU + Q[q].cz] + IP_BETA + contrib);
      }
      w.sync();
      {
        const double* yk = sy + 12 * k;
        double uv[12];
        ld12(u, uv);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const Row& X = R[q];
          const int rr = X.rr;
          double fu = 0.0, fxv, bd;
          if (fused) { fxv = ws[FW_FX + 12 * k + rr]; bd = ws[FW_CD + (k == 0 ? 12 : 0) + rr]; }
          else { fu = ldg(gF + 120 + 12 * k + rr); fxv = ldg(gF + 12 * k + rr); bd = ldg(gB + 12 * k + rr); }
          double acc = fma(f[FB_HU + rr], u[rr], fu) + bcol_dot(X, yk);
          if (X.psk == 3) acc += sy[120 + 2 * k + X.psi_i];
          double gdg, gz;
          gather_col(X, trd, trd + 16, gdg, gz);
          acc += gz;
          a_rx = fma(valid * acc, acc, a_rx);
          if (own) vx[120 + 12 * k + rr] = -acc - gdg;
          const double hx_rr = f[FB_HX + rr];
          double ax = fma(hx_rr, sx[12 * k + rr], fxv) + yk[rr];
          if (k <= 8) {
            const double* yn = sy + 12 * (k + 1);
            ax = fma(hx_rr + IP_BETA, X.c1[0] * yn[rr] + X.c1[1] * yn[X.i1[0]] + X.c1[2] * yn[X.i1[1]] + X.c1[3] * yn[X.i1[2]], ax);
          }
          a_rx = fma(valid * ax, ax, a_rx);
          if (own) vx[12 * k + rr] = -ax;
          double re = sx[12 * k + rr] - bd + brow_dot(X, uv);
          if (k >= 1) {
            const double* xp = xs + 12 * (k - 1);
            re += X.c0[0] * xp[rr] + X.c0[1] * xp[X.i0[0]] + X.c0[2] * xp[X.i0[1]] + X.c0[3] * xp[X.i0[2]];
          }
          a_re = fma(valid * re, re, a_re);
          if (own) vy[12 * k + rr] = re;
        }
        if (l >= 6) {   // Mx rows: lanes 6, 7 of the group
          const int i = l - 6;
          const double bm = fused ? 0.0 : ldg(gB + 120 + 2 * k + i);
          const double rm = u[6 + 3 * i] - bm;
          a_re = fma(rm, rm, a_re);
          vy[120 + 2 * k + i] = rm;
        }
      }
      w.sync();
    }
    a_sz = env_sum(a_sz); a_rs = env_sum(a_rs); a_rx = env_sum(a_rx); a_re = env_sum(a_re);
    mu = a_sz / (double)NINEQ;
    n_rx = sqrt(a_rx); n_rs = sqrt(a_rs); n_re = sqrt(a_re);
  }

  SRBD_DEV void corrector_rhs(int k, double smu) const {
    double* vx = sm + SM::OFF_VX;
    double* ws = gws();
    const double* oS = gout(IPO_S, NINEQ); const double* oZ = gout(IPO_Z, NINEQ);
    double* trd = tr();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int jr = 8 * q + l, j = 16 * k + jr;
      const double sv = oS[j], zv = oZ[j], dsa = ws[FW_S + j], dza = ws[FW_Z + j], qq = ws[FW_Q + j];
      const double r2 = -(rcp(sv) * (sv * zv + dsa * dza - smu));
      const double Dg = qq * r2;
      if (wr) { ws[FW_S + j] = fma(IP_DELTA, Dg, dsa); ws[FW_Z + j] = dza + Dg; }
      trd[jr] = Dg;
    }
    w.sync();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      double gdg, gz;
      gather_col(R[q], trd, nullptr, gdg, gz);
      if (own) vx[120 + 12 * k + R[q].rr] = -gdg;
    }
    w.sync();
  }

  SRBD_DEV void recover(int k, int pass, double& mn_s, double& mn_z) const {
    const double* f = fb();
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* ws = gws();
    const double* oS = gout(IPO_S, NINEQ); const double* oZ = gout(IPO_Z, NINEQ);
    double sv[2], zv[2], sS[2], sZ[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { const int j = 16 * k + 8 * q + l; sv[q] = oS[j]; zv[q] = oZ[j]; sS[q] = ws[FW_S + j]; sZ[q] = ws[FW_Z + j]; }
    double* b = bc();
    double* vu = vx + 120 + 12 * k;
    const double* dyk = vy + 12 * k;
#pragma unroll
    for (int q = 0; q < 2; ++q) b[R[q].rr] = vu[R[q].rr] - bcol_dot(R[q], dyk);
    w.sync();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const Row& X = R[q];
      const int rr = X.rr;
      double du = psi_apply(X, psi(k), b);
      if (X.psk == 3) {
        const double r4 = pass ? 0.0 : -vy[120 + 2 * k + X.psi_i];
        const double eps = f[FB_DER + 6 + X.psi_i], iden = f[X.moff];
        du = fma(IP_DELTA, b[rr], r4) * iden;
        if (own) vy[120 + 2 * k + X.psi_i] = (b[rr] - eps * r4) * iden;
      }
      double wx = (pass ? 0.0 : vx[12 * k + rr]) - dyk[rr];
      if (k <= 8) {
        const double* dyn = vy + 12 * (k + 1);
        wx = fma(-(f[FB_HX + rr] + IP_BETA), X.c1[0] * dyn[rr] + X.c1[1] * dyn[X.i1[0]] + X.c1[2] * dyn[X.i1[1]] + X.c1[3] * dyn[X.i1[2]], wx);
      }
      if (own) { vx[12 * k + rr] = X.phx * wx; vu[rr] = du; }
    }
    w.sync();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = 16 * k + 8 * q + l;
      const double gdu = f[Q[q].goa] * vu[Q[q].ca] + f[Q[q].gob] * vu[Q[q].cz];
      const double Dg = sm[SM::OFF_D + j] * gdu;
      const double dz = sZ[q] + Dg;
      const double ds = fma(IP_DELTA, Dg, sS[q] - gdu);
      if (wr) { ws[FW_S + j] = ds; ws[FW_Z + j] = dz; }
      mn_s = fmin(mn_s, ratio(sv[q], ds));
      mn_z = fmin(mn_z, ratio(zv[q], dz));
    }
  }

  // ------------------------------------------------------------------------------------
  // the whole call for the environment of this lane (both environments of the warp run in lock
  // step; `live` = this lane's environment exists and has not converged yet)
  // ------------------------------------------------------------------------------------
  SRBD_DEV int run(const PdipmArgs& A, long long env, bool exists) {
    pa = &A; e = env;
    lane = w.lane(); l16 = lane & 15; l = lane & 7; h = (lane >> 3) & 1; own = l < 6; wr = exists;
    double* oX = gout(IPO_X, NZ); double* oS = gout(IPO_S, NINEQ); double* oZ = gout(IPO_Z, NINEQ); double* oY = gout(IPO_Y, NEQ);
    const bool keep_xy = A.cold_start && A.warm_xy && !(A.reset && A.reset[e] != 0);
    if (exists) {
      if (A.cold_start) {
        if (!keep_xy) {
          for (int i = l16; i < NZ; i += 16) oX[i] = 0.0;
          for (int i = l16; i < NEQ; i += 16) oY[i] = 1.0;
        }
        for (int i = l16; i < NINEQ; i += 16) oZ[i] = 1.0;
      } else {
        const double* ix = gin(IP_X, NZ); const double* is = gin(IP_S, NINEQ);
        const double* iz = gin(IP_Z, NINEQ); const double* iy = gin(IP_Y, NEQ);
        for (int i = l16; i < NZ; i += 16) oX[i] = ix[i];
        for (int i = l16; i < NINEQ; i += 16) { oS[i] = is[i]; oZ[i] = iz[i]; }
        for (int i = l16; i < NEQ; i += 16) oY[i] = iy[i];
      }
    }
    w.sync();
    const int bad = setup();
    bool live = exists;
    if (bad && exists) {
      const double nan = __builtin_nan("");
      for (int i = l16; i < NZ; i += 16) oX[i] = nan;
      for (int i = l16; i < NINEQ; i += 16) { oS[i] = nan; oZ[i] = nan; }
      for (int i = l16; i < NEQ; i += 16) oY[i] = nan;
      if (l16 == 0) {
        double* oRes = gout(IPO_RES, 4);
        oRes[0] = nan; oRes[1] = nan; oRes[2] = nan; oRes[3] = nan; gout(IPO_MU, 1)[0] = nan;
        if (A.iters_done) A.iters_done[e] = -1;
      }
      live = false;
    }
    if (live && A.cold_start) {
#pragma unroll
      for (int t = 0; t < 5; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = h ? 9 - t : t, j = 16 * k + 8 * q + l;
          const double gx = keep_xy ? fb()[Q[q].goa] * oX[120 + 12 * k + Q[q].ca] + fb()[Q[q].gob] * oX[120 + 12 * k + Q[q].cz] : 0.0;
          oS[j] = fmax(h_of(k, j) - gx, 1.0);
        }
    }
    w.sync();

    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    int it = 0;
    double n_rx = 0, n_rs = 0, n_re = 0, mu_new = 0;
#pragma unroll 1
    for (int iter = 0; iter < A.num_iters; ++iter) {
      if (!w.any(live)) break;
      double mu, c_rx, c_rs, c_re;
      phase_a(mu, c_rx, c_rs, c_re);
      double smu = 0.0, mn_s = 1e300, mn_z = 1e300;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t;
          if (pass == 1) corrector_rhs(k, smu);
          build_rho(k, pass == 1);
        }
        if (pass == 0) {
          SRBD_TRACE("rho_aff", vy, 140);
          factor();
          SRBD_TRACE("P", sm + SM::OFF_P, 780);
        }
        solve_thomas();
        if (pass == 0) SRBD_TRACE("dyd_aff", vy, 120);
        mn_s = 1e300; mn_z = 1e300;
#pragma unroll 1
        for (int t = 0; t < 5; ++t) recover(h ? 9 - t : t, pass, mn_s, mn_z);
        if (pass == 0) {
          w.sync();
          double* ws = gws();
          if (wr) {
            for (int i = l16; i < NZ; i += 16) ws[FW_DX + i] = vx[i];
            for (int i = l16; i < NEQ; i += 16) ws[FW_DY + i] = vy[i];
          }
          SRBD_TRACE("dx_aff", vx, 240);
          SRBD_TRACE("dy_aff", vy, 140);
          const double ap = step_len(env_min(mn_s));
          const double ad = step_len(env_min(mn_z));
          double a_mu = 0.0;
#pragma unroll
          for (int t = 0; t < 5; ++t)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = 16 * (h ? 9 - t : t) + 8 * q + l;
              a_mu = fma(fma(ap, ws[FW_S + j], oS[j]), fma(ad, ws[FW_Z + j], oZ[j]), a_mu);
            }
          const double mu_aff = env_sum(a_mu) / (double)NINEQ;
          const double sg = mu_aff * rcp(mu);
          smu = sg * sg * sg * mu;
          w.sync();
        }
      }
      const double a_pri = step_len(env_min(mn_s));
      const double a_dual = step_len(env_min(mn_z));
      w.sync();
      double a_new = 0.0;
      if (live) {
        double* ws = gws();
        for (int i = l16; i < NZ; i += 16) oX[i] = fma(a_pri, ws[FW_DX + i] + vx[i], oX[i]);
        for (int i = l16; i < NEQ; i += 16) oY[i] = fma(a_dual, ws[FW_DY + i] + vy[i], oY[i]);
#pragma unroll
        for (int t = 0; t < 5; ++t)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int j = 16 * (h ? 9 - t : t) + 8 * q + l;
            const double sn = fmax(fma(a_pri, ws[FW_S + j], oS[j]), 1e-8);
            const double zn = fmax(fma(a_dual, ws[FW_Z + j], oZ[j]), 1e-8);
            oS[j] = sn; oZ[j] = zn;
            a_new = fma(sn, zn, a_new);
          }
      }
      const double mu_it = env_sum(a_new) / (double)NINEQ;
      w.sync();
      if (live) {
        n_rx = c_rx; n_rs = c_rs; n_re = c_re; mu_new = mu_it; ++it;
        if (A.tol > 0.0 && n_rx < A.tol && n_rs < A.tol && n_re < A.tol && mu_new < A.tol) live = false;
      }
    }
    if (exists && !bad) {
      if (l16 == 0) {
        double* oRes = gout(IPO_RES, 4);
        oRes[0] = n_rx; oRes[1] = n_rs; oRes[2] = n_re; oRes[3] = mu_new;
        gout(IPO_MU, 1)[0] = mu_new;
        if (A.iters_done) A.iters_done[e] = it;
      }
      if (A.fused && A.wrench && l16 < 12)
        A.wrench[e * 12 + l16] = wrench_entry(gout(IPO_X, NZ) + NX * N, A.rot + e * 9, l16);
    }
    return it;
  }
};

}  // namespace srbd
