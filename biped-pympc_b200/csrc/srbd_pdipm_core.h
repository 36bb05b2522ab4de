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
//                                                  from both ends of the horizon; pivot blocks
//                                                  inverted by a symmetric sweep
//
// Thread mapping: lane = 16*h + r.  Half h works on stage k = t (h = 0) or 9 - t (h = 1) in
// round t; r is a row (state / input / dynamics-row / inequality-row index) of that stage.
//
// kUniform = true is the fast path for stage-invariant problem data (what `qp_former`
// always produces): the stage matrices, coupling coefficients and the state part of the
// Schur blocks are built once per call and stay in shared memory / registers.
// kUniform = false accepts arbitrary values on the frozen pattern and re-gathers the stage
// data from the CCS vectors every round.
//
// The same source is compiled for the GPU (srbd_pdipm.cu) and, for tests only, for the host
// under the fiber-based warp emulator of srbd_warp.h.
#pragma once

#include "srbd_args.h"
#include "srbd_pattern.h"
#include "srbd_warp.h"

namespace srbd {

// ---- stage buffer layout (doubles); all double2-accessed blocks start at even offsets -----
enum : int {
  SB_B = 0,        // Bbar_k dense [12][14] (row stride 14 keeps rows 16-byte aligned)
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
  SB_RAW_END = 318,
  SB_CM = 326,     // 2 x 36 coupling coefficient slots: [0] = rows of C_k, [1] = rows of C_{k+1}'
  SB_SX = 398,     // state part of S_kk, packed lower (78)
  SB_SX0 = 476,    // its k = 0 variant (diagonal only, 12)
  SB_SIZE = 488
};
enum : int { ST_ZERO = 0, ST_AU, ST_E, ST_AXPREV, ST_AXCUR, ST_G, ST_Q, ST_QINV, ST_QINVPREV, ST_EPREV };
struct StageSlot { unsigned char kind; short rel; };

// per-env scratch in global memory (L2 resident): affine direction + per-row barrier data
enum : int { WK_DX = 0, WK_DY = 240, WK_DS = 380, WK_DZ = 540, WK_D = 700, WK_WI = 860, WK_DG = 1020, WK_R2 = 1180,
             WK_STAGE = 1344, WK_STAGE_STRIDE = 488 };   // general kernel: the ten stage buffers, prepared once
static_assert(WK_STAGE + 10 * WK_STAGE_STRIDE <= PDIPM_WORK_DOUBLES, "work buffer too small");
static_assert(WK_R2 + 160 <= WK_STAGE, "work buffer layout");

template <bool kUniform>
struct PdipmSmem {
  static constexpr int NSB = kUniform ? 1 : 2;
  static constexpr int OFF_P = 0;          // 10 x 78 packed lower; scratch for iterates in phase A
  static constexpr int OFF_VX = 780;       // 240  r1~ -> dx
  static constexpr int OFF_VY = 1020;      // 140  re -> rho -> t -> dy
  static constexpr int OFF_PSI = 1160;     // 10 x 16
  static constexpr int OFF_BC = 1320;      // 2 x 36 broadcast scratch per half
  static constexpr int OFF_DGS = 1392;     // 2 x 16 D*g of the stage being processed
  static constexpr int OFF_SB = 1424;      // NSB x SB_SIZE
  static constexpr int SIZE = OFF_SB + NSB * SB_SIZE;
};

#if defined(__CUDACC__) && !defined(SRBD_HOST_EMUL)
typedef double2 dbl2;
// 1/x to ~1 ulp: 20-bit seed + one cubic step (no slow-path call, unlike the IEEE division)
SRBD_DEV double rcp(double x) {
#ifdef SRBD_IEEE_DIV
  return 1.0 / x;
#endif
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // one third-order step e <- e + e^2 (2^-20 -> 2^-60): three dependent FMAs.  Two Newton steps (four FMAs) are not more
  // accurate after rounding and sit on the critical path of every pivot step of the sweep (measured: -0.7 %)
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}
#else
struct alignas(16) dbl2 { double x, y; };
SRBD_DEV double rcp(double x) { return 1.0 / x; }
#endif

// single-precision counterparts (the optional fp32 pivot-block mode of srbd_pdipm_fast.h)
#if defined(__CUDACC__) && !defined(SRBD_HOST_EMUL)
typedef float4 flt4;
SRBD_DEV float rcp(float x) { return __frcp_rn(x); }
#else
struct alignas(16) flt4 { float x, y, z, w; };
SRBD_DEV float rcp(float x) { return 1.0f / x; }
#endif
SRBD_DEV double fmaR(double a, double b, double c) { return fma(a, b, c); }
SRBD_DEV float fmaR(float a, float b, float c) { return fmaf(a, b, c); }

SRBD_DEV int pk(int i, int j) { return i >= j ? ((i * (i + 1)) >> 1) + j : ((j * (j + 1)) >> 1) + i; }
// 12 doubles from a 16-byte aligned shared-memory address
SRBD_DEV void ld12(const double* p, double v[12]) {
  // (twelve 8-byte loads instead were measured: 20 % fewer shared-memory wavefronts, twice the
  //  load instructions, 1 % slower)
  const dbl2* q = reinterpret_cast<const dbl2*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i) { const dbl2 t = q[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}
SRBD_DEV void st12(double* p, const double v[12]) {
  dbl2* q = reinterpret_cast<dbl2*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i) { dbl2 t; t.x = v[2 * i]; t.y = v[2 * i + 1]; q[i] = t; }
}

SRBD_DEV void ld12(const float* p, float v[12]) {
  const flt4* q = reinterpret_cast<const flt4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i) { const flt4 t = q[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}
SRBD_DEV void st12(float* p, const float v[12]) {
  flt4* q = reinterpret_cast<flt4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i) { flt4 t; t.x = v[4 * i]; t.y = v[4 * i + 1]; t.z = v[4 * i + 2]; t.w = v[4 * i + 3]; q[i] = t; }
}

template <bool kUniform>
struct PdipmWarp {
  using SM = PdipmSmem<kUniform>;
  Warp w;
  double* sm;                    // this warp's shared memory
  const StageSlot* slots;        // gather table (SB_RAW_END entries used)
  const double *gQ, *gG, *gA, *gF, *gH, *gB;
  double *oX, *oS, *oZ, *oY, *oRes, *oMu, *ws;
  int lane, h, r, rr;
  int prow[12];                  // packed offsets of row rr: prow[m] = pk(rr, m)
  // sparse coupling rows of this lane: [0] = row r of C_k ("LP"), [1] = row r of C_{k+1}' ("LN")
  int cidx[2][4];
  double cval[2][4];
  int ps_kind, ps_ft, ps_i;      // role of input column rr in Psi: 0 fx/fy/my, 1 fz, 2 mz, 3 mx

  SRBD_DEV double* P(int k) const { return sm + SM::OFF_P + 78 * k; }
  SRBD_DEV double* sb(int hh) const { return sm + SM::OFF_SB + (kUniform ? 0 : hh * SB_SIZE); }
  SRBD_DEV double* bc(int hh) const { return sm + SM::OFF_BC + 36 * hh; }
  SRBD_DEV double* psi(int k) const { return sm + SM::OFF_PSI + 16 * k; }

  // ------------------------------------------------------------------------------------
  // stage data: gather from the CCS value vectors + derived (iteration-invariant) quantities
  // ------------------------------------------------------------------------------------
  // value of raw slot i of stage k, gathered from the CCS value vectors
  SRBD_DEV double gather_slot(int i, int k) const {
    const StageSlot sl = slots[i];
    double v = 0.0;
    switch (sl.kind) {
      case ST_AU: v = ldg(gA + 336 + 86 * k + sl.rel); break;
      case ST_E: v = ldg(gA + a_xcol_start(k, sl.rel)); break;
      case ST_AXPREV: if (k >= 1) v = ldg(gA + 36 * (k - 1) + sl.rel); break;
      case ST_AXCUR: if (k <= 8) v = ldg(gA + 36 * k + sl.rel); break;
      case ST_G: v = ldg(gG + 28 * k + sl.rel); break;
      case ST_Q: v = ldg(gQ + 12 * k + sl.rel); break;
      case ST_QINV: v = rcp(ldg(gQ + 12 * k + sl.rel) + IP_BETA); break;
      case ST_QINVPREV: if (k >= 1) v = rcp(ldg(gQ + 12 * (k - 1) + sl.rel) + IP_BETA); break;
      case ST_EPREV: if (k >= 1) v = ldg(gA + a_xcol_start(k - 1, sl.rel)); break;
      default: break;
    }
    return v;
  }
  SRBD_DEV void load_stage_raw(int hh, int k, int first, int step) const {
    double* s = sb(hh);
    for (int i = first; i < SB_RAW_END; i += step) s[i] = gather_slot(i, k);
  }
  // Abar(j, c) of a compact 24-entry block (0 where structurally absent)
  SRBD_DEV static double abar(const double* A, int j, int c) {
    if (c == j) return A[j];
    if (j < 3 && c >= 6 && c < 9) return A[12 + 3 * j + (c - 6)];
    if (j >= 3 && j < 6 && c == j + 6) return A[21 + j - 3];
    return 0.0;
  }
  // everything that depends only on the stage data (not on the iterate); q = row index of the
  // calling lane (0..15).  Writes shared memory of stage buffer hh and this lane's registers.
  SRBD_DEV void stage_derived(int hh, int q) {
    double* s = sb(hh);
    if (q < 2) {                  // mz: plain diagonal of Phi_u
      s[SB_DER + q] = rcp(s[SB_HU + 8 + 3 * q] + IP_BETA);
    } else if (q < 4) {           // mx: eliminated together with its equality row
      const int i = q - 2;
      const double eps = s[SB_HU + 6 + 3 * i] + IP_BETA;
      const double m = s[SB_MX + i];
      const double iden = rcp(m * m + IP_DELTA * eps);
      s[SB_DER + 2 + i] = IP_DELTA * iden;   // psi'
      s[SB_DER + 4 + i] = iden;
      s[SB_DER + 6 + i] = eps;
    }
    const int qq = q < 12 ? q : 0;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int a = 0; a < 4; ++a) { cidx[t][a] = qq; cval[t][a] = 0.0; }
    if (q < 12) {
      {   // LP: row q of C_k = Abar_k diag(phix_{k-1} E_{k-1})
        const double* A = s + SB_AP; const double* ph = s + SB_PHXP; const double* e = s + SB_EP;
        cval[0][0] = A[q] * (ph[q] * e[q]);
        if (q < 3) {
#pragma unroll
          for (int j = 0; j < 3; ++j) { cidx[0][1 + j] = 6 + j; cval[0][1 + j] = A[12 + 3 * q + j] * (ph[6 + j] * e[6 + j]); }
        } else if (q < 6) { cidx[0][1] = q + 6; cval[0][1] = A[21 + q - 3] * (ph[q + 6] * e[q + 6]); }
      }
      {   // LN: row q of C_{k+1}' , C_{k+1} = Abar_{k+1} diag(phix_k E_k)
        const double* A = s + SB_AN;
        const double wr = s[SB_PHX + q] * s[SB_E + q];
        cval[1][0] = A[q] * wr;
        if (q >= 6 && q < 9) {
#pragma unroll
          for (int i = 0; i < 3; ++i) { cidx[1][1 + i] = i; cval[1][1 + i] = A[12 + 3 * i + (q - 6)] * wr; }
        } else if (q >= 9) { cidx[1][1] = q - 6; cval[1][1] = A[21 + q - 9] * wr; }
      }
      // 36 static slots per type: [0..11] diagonal, [12..20] rows 0-2 x cols 6-8, [21..23] rows
      // 3-5 (col r+6), [24..32] rows 6-8 x cols 0-2, [33..35] rows 9-11 (col r-6)
      double* c0 = s + SB_CM; double* c1 = s + SB_CM + 36;
      c0[q] = cval[0][0]; c1[q] = cval[1][0];
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { c0[12 + 3 * q + j] = cval[0][1 + j]; c1[12 + 3 * q + j] = 0.0; }
      } else if (q < 6) { c0[21 + q - 3] = cval[0][1]; c1[21 + q - 3] = 0.0; }
      else if (q < 9) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { c0[24 + 3 * (q - 6) + i] = 0.0; c1[24 + 3 * (q - 6) + i] = cval[1][1 + i]; }
      } else { c0[33 + q - 9] = 0.0; c1[33 + q - 9] = cval[1][1]; }
      // state part of S_kk: delta I + E phix E + Abar_k phix_{k-1} Abar_k'   (row q, lower part)
      const double* A = s + SB_AP; const double* ph = s + SB_PHXP;
      int ai[4]; double av[4];
      ai[0] = q; ai[1] = q; ai[2] = q; ai[3] = q;
      av[0] = A[q] * ph[q]; av[1] = 0.0; av[2] = 0.0; av[3] = 0.0;
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { ai[1 + j] = 6 + j; av[1 + j] = A[12 + 3 * q + j] * ph[6 + j]; }
      } else if (q < 6) { ai[1] = q + 6; av[1] = A[21 + q - 3] * ph[q + 6]; }
      const double dg = s[SB_E + q] * s[SB_E + q] * s[SB_PHX + q] + IP_DELTA;
      for (int j = 0; j <= q; ++j) {
        double acc = (j == q) ? dg : 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) acc += av[a] * abar(A, j, ai[a]);
        s[SB_SX + pk(q, j)] = acc;
      }
      s[SB_SX0 + q] = dg;
    }
  }
  // this lane's coupling rows (registers) from the coefficient slots of stage buffer hh
  SRBD_DEV void restore_coupling(int hh, int q) {
    const double* c0 = sb(hh) + SB_CM; const double* c1 = c0 + 36;
    const int qq = q < 12 ? q : 0;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int a = 0; a < 4; ++a) { cidx[t][a] = qq; cval[t][a] = 0.0; }
    if (q < 12) {
      cval[0][0] = c0[q]; cval[1][0] = c1[q];
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { cidx[0][1 + j] = 6 + j; cval[0][1 + j] = c0[12 + 3 * q + j]; }
      } else if (q < 6) { cidx[0][1] = q + 6; cval[0][1] = c0[21 + q - 3]; }
      else if (q < 9) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { cidx[1][1 + i] = i; cval[1][1 + i] = c1[24 + 3 * (q - 6) + i]; }
      } else { cidx[1][1] = q - 6; cval[1][1] = c1[33 + q - 9]; }
    }
  }
  // general kernel: the ten stages do not fit in shared memory.  Their stage buffers (gathered values +
  // everything derived from them) are prepared ONCE per call and parked in the work buffer; the
  // per-round reload is then a coalesced 488-double copy instead of 318 scattered, switch-dispatched
  // loads plus the derivation
  SRBD_DEV void prepare_stages() {
#pragma unroll 1
    for (int t = 0; t < 5; ++t) {
      const int k = 2 * t + h;
      w.sync();
      load_stage_raw(h, k, r, 16);
      w.sync();
      stage_derived(h, r);
      w.sync();
      const double* s = sb(h);
      double* dst = ws + WK_STAGE + WK_STAGE_STRIDE * k;
      for (int i = r; i < SB_SIZE; i += 16) dst[i] = s[i];
    }
    w.sync();
  }
  // per-round (general) or once (stage-invariant) stage load; all lanes must call
  SRBD_DEV void stage(int k, int lo = 0, int hi = SB_SIZE) {
    if (!kUniform) {
      w.sync();
      double* s = sb(h);
      const double* src = ws + WK_STAGE + WK_STAGE_STRIDE * k;
      for (int i = lo + r; i < hi; i += 16) s[i] = src[i];
      w.sync();
      restore_coupling(h, r);
    }
  }
  SRBD_DEV void stage_once() {
    if (kUniform) {
      load_stage_raw(0, 1, lane, 32);
      w.sync();
      // both halves compute the (identical) derived data: every lane needs its coupling rows
      // in registers; the duplicate shared-memory writes store identical values
      stage_derived(0, r);
      w.sync();
    }
  }
  // x-part of S_kk(rr, j)
  SRBD_DEV double sx_entry(const double* s, int k, int j) const {
    if (kUniform && k == 0) return j == rr ? s[SB_SX0 + rr] : 0.0;
    return s[SB_SX + prow[j]];
  }

  // ------------------------------------------------------------------------------------
  // Psi_k applied to a 12-vector held in shared memory at `v`; returns entry rr
  // ------------------------------------------------------------------------------------
  SRBD_DEV void psi_init() {
    const int c = rr;
    ps_ft = (c < 6) ? c / 3 : (c >= 9 ? 1 : 0);
    if (c < 6) { ps_kind = (c % 3 == 2) ? 1 : 0; ps_i = c % 3; }
    else if (c == 7 || c == 10) { ps_kind = 0; ps_i = 2; }
    else if (c == 8 || c == 11) { ps_kind = 2; ps_i = (c == 11); }
    else { ps_kind = 3; ps_i = (c == 9); }
  }
  SRBD_DEV double psi_apply(const double* ps, const double* der, const double* v) const {
    // ps: [ft*7 + {0..2: 1/a (fx,fy,my), 3..5: t (fx,fy,my), 6: 1/sigma}]
    const double* p = ps + 7 * ps_ft;
    const int fx = 3 * ps_ft;
    const double dot = p[3] * v[fx] + p[4] * v[fx + 1] + p[5] * v[7 + fx] - v[fx + 2];
    double diag, q;
    if (ps_kind == 0) { diag = p[ps_i]; q = p[3 + ps_i]; }
    else if (ps_kind == 1) { diag = 0.0; q = -1.0; }
    else { diag = der[(ps_kind == 2 ? 0 : 2) + ps_i]; q = 0.0; }
    return diag * v[rr] + q * (p[6] * dot);
  }

  // ------------------------------------------------------------------------------------
  // rho_k (into vy) from r1~ (vx) and the equality right-hand sides held in vy:
  // on entry vy[12k+r] = -r4_dyn (`re`, or 0 for the corrector), vy[120+2k+i] = -r4_mx
  // ------------------------------------------------------------------------------------
  SRBD_DEV void build_rho(int k) const {
    const double* s = sb(h);
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    if (r < 12) {
      double v = psi_apply(psi(k), s + SB_DER, vx + 120 + 12 * k);
      if (ps_kind == 3) v += s[SB_MX + ps_i] * s[SB_DER + 4 + ps_i] * (-vy[120 + 2 * k + ps_i]);   // (m/den) r4_mx
      b[r] = v;
    }
    w.sync();
    if (r < 12) {
      double bv[12], Br[12];
      ld12(b, bv);
      ld12(s + SB_B + 14 * r, Br);
      double r0 = s[SB_E + r] * s[SB_PHX + r] * vx[12 * k + r] + vy[12 * k + r], r1 = 0.0;
#pragma unroll
      for (int c = 0; c < 12; c += 2) { r0 += Br[c] * bv[c]; r1 += Br[c + 1] * bv[c + 1]; }
      if (k >= 1) {
        const double* A = s + SB_AP;
        const double* ph = s + SB_PHXP;
        const double* xp = vx + 12 * (k - 1);
        r1 += A[r] * ph[r] * xp[r];
        if (r < 3) { for (int j = 0; j < 3; ++j) r1 += A[12 + 3 * r + j] * ph[6 + j] * xp[6 + j]; }
        else if (r < 6) r1 += A[21 + r - 3] * ph[r + 6] * xp[r + 6];
      }
      vy[12 * k + r] = r0 + r1;
    }
    w.sync();
  }

  // ------------------------------------------------------------------------------------
  // row rr of S_kk into a[12] (uses Psi_k, the stage matrices and the hoisted state part)
  // ------------------------------------------------------------------------------------
  SRBD_DEV void form_row(int k, double a[12]) const {
    const double* s = sb(h);
    const double* ps = psi(k);
    const double* der = s + SB_DER;
    double* b = bc(h);
    double wv[12];
    double pL, pR;
    {
      double Br[12];
      ld12(s + SB_B + 14 * rr, Br);
      pL = ps[3] * Br[0] + ps[4] * Br[1] + ps[5] * Br[7] - Br[2];
      pR = ps[10] * Br[3] + ps[11] * Br[4] + ps[12] * Br[10] - Br[5];
      wv[0] = ps[0] * Br[0]; wv[1] = ps[1] * Br[1]; wv[2] = 0.0;
      wv[3] = ps[7] * Br[3]; wv[4] = ps[8] * Br[4]; wv[5] = 0.0;
      wv[6] = der[2] * Br[6]; wv[7] = ps[2] * Br[7]; wv[8] = der[0] * Br[8];
      wv[9] = der[3] * Br[9]; wv[10] = ps[9] * Br[10]; wv[11] = der[1] * Br[11];
    }
    if (r < 12) { b[12 + r] = pL; b[24 + r] = pR; }
    w.sync();
    const double sL = ps[6] * pL, sR = ps[13] * pR;
    double pl[12], pr[12];
    ld12(b + 12, pl);
    ld12(b + 24, pr);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const double* Bj = s + SB_B + 14 * j;
      double acc = sL * pl[j] + sR * pr[j] + sx_entry(s, k, j);
      if (j < 3 || (j >= 6 && j < 9)) {          // dense rows of Bbar (euler, omega)
        double bj[12];
        ld12(Bj, bj);
        double acc2 = 0.0;
#pragma unroll
        for (int c = 0; c < 12; c += 2) { acc += wv[c] * bj[c]; acc2 += wv[c + 1] * bj[c + 1]; }
        acc += acc2;
      } else {                                    // position / velocity rows: two entries
        const int c1 = (j < 6) ? j - 3 : j - 9, c2 = c1 + 3;
        acc += wv[c1] * Bj[c1] + wv[c2] * Bj[c2];
      }
      a[j] = acc;
    }
    w.sync();   // b (pL/pR) is reused by the caller
  }

  // ------------------------------------------------------------------------------------
  // one elimination round: a = row rr of S_kk - M P_nb M'  (M = coupling of `type`);
  // mode 0: invert (symmetric sweep) and store P_k ; mode 1: store the updated S~ row only;
  // load: the row was prepared by an earlier round (in P_k) instead of being formed here
  // ------------------------------------------------------------------------------------
  SRBD_DEV void factor_round(int k, int nb, int type, int mode, bool load, bool act) {
    double a[12];
    double* Pk = P(k);
    const double* s = sb(h);
    if (load) {
#pragma unroll
      for (int c = 0; c < 12; ++c) a[c] = Pk[prow[c]];
      w.sync();
    } else {
      form_row(k, a);
    }
    if (nb >= 0) {
      const double* c = s + SB_CM + 36 * type;
      const double* Pn = P(nb);
      double U[12];
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = cval[type][0] * Pn[prow[m]];
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const int i = cidx[type][q];
        const double v = cval[type][q];
        const int ib = (i * (i + 1)) >> 1;
#pragma unroll
        for (int m = 0; m < 12; ++m) U[m] += v * Pn[m <= i ? ib + m : ((m * (m + 1)) >> 1) + i];
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
    if (mode == 1 && act && r < 12) {
#pragma unroll
      for (int c = 0; c < 12; ++c) if (c <= r) Pk[prow[c]] = a[c];
    }
    // symmetric sweep on all 12 pivots: every lane keeps one full row, the unscaled pivot row
    // is broadcast through shared memory; the result is -S~^-1 (negated on store).  Unlike a
    // plain Gauss-Jordan inverse this stays symmetric, so keeping only the lower triangle
    // costs no accuracy.  (Executed, with the result discarded, also for mode 1 / inactive
    // halves, which keeps the warp convergent.)
    double* b = bc(h);
#pragma unroll
    for (int p = 0; p < 12; ++p) {
      if (r == p) st12(b, a);
      w.sync();
      double bv[12];
      ld12(b, bv);
      const double di = rcp(bv[p]);
      const double t = b[rr] * di;
      if (r == p) {
#pragma unroll
        for (int c = 0; c < 12; ++c) a[c] = (c == p) ? -di : bv[c] * di;
      } else {
#pragma unroll
        for (int c = 0; c < 12; ++c) a[c] = (c == p) ? t : a[c] - t * bv[c];
      }
      w.sync();
    }
    if (mode == 0 && act && r < 12) {
#pragma unroll
      for (int c = 0; c < 12; ++c) if (c <= r) Pk[prow[c]] = -a[c];
    }
    w.sync();
  }

  SRBD_DEV void factor() {
    // rounds 0..3: top k = t (nb k-1, LP) | bottom k = 9-t (nb k+1, LN)
    // round 4: top factors k=4 ; bottom forms S_55 and applies the k=6 update (store only)
    // round 5: top loads the prepared S_55, applies the k=4 update and inverts ; bottom idle
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      int k, nb, type, mode;
      bool act = true, load = false;
      if (t < 4) { k = h ? 9 - t : t; nb = (t == 0) ? -1 : (h ? k + 1 : k - 1); type = h; mode = 0; }
      else if (t == 4) { k = h ? 5 : 4; nb = h ? 6 : 3; type = h; mode = h ? 1 : 0; }
      else { k = 5; nb = 4; type = 0; mode = 0; act = (h == 0); load = true; }
      stage(k);
      factor_round(k, nb, type, mode, load, act);
    }
  }

  // ------------------------------------------------------------------------------------
  // block-Thomas solve on vy[0..119]  (rho -> t -> dy), twisted
  // ------------------------------------------------------------------------------------
  SRBD_DEV double prow_dot(const double* Pm, const double* v) const {
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
    for (int m = 0; m < 12; m += 3) { q0 += Pm[prow[m]] * v[m]; q1 += Pm[prow[m + 1]] * v[m + 1]; q2 += Pm[prow[m + 2]] * v[m + 2]; }
    return q0 + q1 + q2;
  }
  SRBD_DEV void solve_thomas() {
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    // forward: t_k -= M (P_nb t_nb)
#pragma unroll 1
    for (int t = 1; t < 6; ++t) {
      int k, nb, type;
      bool act = true;
      if (t < 4) { k = h ? 9 - t : t; nb = h ? k + 1 : k - 1; type = h; }
      else if (t == 4) { k = h ? 5 : 4; nb = h ? 6 : 3; type = h; }
      else { k = 5; nb = 4; type = 0; act = (h == 0); }
      stage(k, SB_CM, SB_CM + 72);   // the substitution rounds only use the coupling rows
      double tv[12];
      ld12(vy + 12 * nb, tv);
      const double q = prow_dot(P(nb), tv);
      if (r < 12) b[r] = q;
      w.sync();
      if (act && r < 12) {
        double tk = vy[12 * k + r];
#pragma unroll
        for (int a = 0; a < 4; ++a) tk -= cval[type][a] * b[cidx[type][a]];
        vy[12 * k + r] = tk;
      }
      w.sync();
    }
    // backward: dy_5 ; then 4|6 ; 3|7 ; 2|8 ; 1|9 ; 0|-   (top uses C_{k+1}', bottom uses C_k)
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      int k = (t == 0) ? 5 : (h ? 5 + t : 5 - t);
      const bool act = (t == 0) ? (h == 0) : (k <= 9);
      if (k > 9) k = 9;
      const int nb = (t == 0) ? -1 : (h ? k - 1 : k + 1);
      const int type = h ? 0 : 1;
      stage(k, SB_CM, SB_CM + 72);
      double tk = vy[12 * k + rr];
      if (nb >= 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) tk -= cval[type][a] * vy[12 * nb + cidx[type][a]];
      }
      if (r < 12) b[r] = tk;
      w.sync();
      double bv[12];
      ld12(b, bv);
      const double dy = prow_dot(P(k), bv);
      if (act && r < 12) vy[12 * k + r] = dy;
      w.sync();
    }
  }

  // ------------------------------------------------------------------------------------
  // recover dx (into vx), dy_mx (into vy) and this lane's ds/dz of inequality row 16k + r
  // ------------------------------------------------------------------------------------
  SRBD_DEV void recover(int k, double& ds_out, double& dz_out) const {
    const double* s = sb(h);
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc(h);
    const int j = 16 * k + r;
    // per-row barrier data (L2): issue the loads early
    const double rD = ws[WK_D + j], rDg = ws[WK_DG + j], rR2 = ws[WK_R2 + j], rWi = ws[WK_WI + j];
    double dyk[12];
    ld12(vy + 12 * k, dyk);
    if (r < 12) {
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int i = 0; i < 12; i += 2) { q0 += s[SB_B + 14 * i + r] * dyk[i]; q1 += s[SB_B + 14 * (i + 1) + r] * dyk[i + 1]; }
      b[r] = vx[120 + 12 * k + r] - (q0 + q1);
    }
    w.sync();
    if (r < 12) {
      double du;
      if (ps_kind == 3) {
        const int i = ps_i;
        const double m = s[SB_MX + i], iden = s[SB_DER + 4 + i], eps = s[SB_DER + 6 + i];
        const double r4 = -vy[120 + 2 * k + i];
        du = (IP_DELTA * b[r] + m * r4) * iden;
        vy[120 + 2 * k + i] = (m * b[r] - eps * r4) * iden;     // dy_mx
      } else {
        du = psi_apply(psi(k), s + SB_DER, b);
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
      const int ft = r >> 3, jj = r & 7;
      const double gb = s[SB_G + 12 * ft + 4 + jj];
      double ga = 0.0; int ca = 3 * ft + 2;
      if (jj < 4) { ga = s[SB_G + 12 * ft + jj]; ca = 3 * ft + (jj >> 1); }
      else if (jj < 6) { ga = s[SB_G + 24 + 2 * ft + jj - 4]; ca = 7 + 3 * ft; }
      const double* du = vx + 120 + 12 * k;
      const double gdu = ga * du[ca] + gb * du[3 * ft + 2];
      const double dz = rD * gdu + rDg;
      dz_out = dz;
      ds_out = (rR2 - dz) * rWi;
    }
  }
  // over the rows of input column rr: gdg = sum_j G[j,c] Dg_j (Dg of the stage in dgs),
  // gz = sum_j G[j,c] z_j (when zrow != nullptr)
  SRBD_DEV void gather_col(const double* zrow, double& gdg_out, double& gz_out) const {
    const double* s = sb(h);
    const double* dgs = sm + SM::OFF_DGS + 16 * h;
    const int c = rr;
    int row0 = 0, cnt = 0, gofs = 0;
    if (c < 6) {
      const int ft = c / 3, q = c % 3;
      if (q == 0) { row0 = 8 * ft; cnt = 2; gofs = 12 * ft; }
      else if (q == 1) { row0 = 8 * ft + 2; cnt = 2; gofs = 12 * ft + 2; }
      else { row0 = 8 * ft; cnt = 8; gofs = 12 * ft + 4; }
    } else if (c == 7 || c == 10) { const int ft = (c == 10); row0 = 8 * ft + 4; cnt = 2; gofs = 24 + 2 * ft; }
    double gdg = 0.0, gz = 0.0;
    for (int i = 0; i < cnt; ++i) {
      const double gv = s[SB_G + gofs + i];
      gdg += gv * dgs[row0 + i];
      if (zrow) gz += gv * zrow[row0 + i];
    }
    gdg_out = gdg;
    gz_out = gz;
  }

  SRBD_DEV double warp_sum(double v) const {
    for (int m = 16; m >= 1; m >>= 1) v += w.shfl_xor(v, m);
    return v;
  }
  SRBD_DEV double warp_min(double v) const {
    for (int m = 16; m >= 1; m >>= 1) { const double o = w.shfl_xor(v, m); v = o < v ? o : v; }
    return v;
  }
  SRBD_DEV static double ratio(double v, double dv) { return dv < 0.0 ? -v * rcp(dv) : 1.0; }
  SRBD_DEV static double step_len(double mn) { return fmax(fmin(1.0, 0.99 * mn), 1e-12); }

  // ------------------------------------------------------------------------------------
  // the whole call: num_iters Newton iterations for this warp's environment
  // ------------------------------------------------------------------------------------
  SRBD_DEV int run(const PdipmArgs& A, long long e) {
    lane = w.lane(); h = lane >> 4; r = lane & 15; rr = r < 12 ? r : 11;
#pragma unroll
    for (int m = 0; m < 12; ++m) prow[m] = pk(rr, m);
    psi_init();
    gQ = A.in[IP_Q] + e * NNZ_H; gG = A.in[IP_G] + e * NNZ_G; gA = A.in[IP_A] + e * NNZ_A;
    gF = A.in[IP_F] + e * NZ; gH = A.in[IP_H] + e * NINEQ; gB = A.in[IP_B] + e * NEQ;
    oX = A.out[IPO_X] + e * NZ; oS = A.out[IPO_S] + e * NINEQ; oZ = A.out[IPO_Z] + e * NINEQ;
    oY = A.out[IPO_Y] + e * NEQ; oRes = A.out[IPO_RES] + e * 4; oMu = A.out[IPO_MU] + e;
    ws = A.work + e * (long long)A.work_stride;

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
    if (!kUniform) prepare_stages();
    w.sync();
    stage_once();

    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* dgs = sm + SM::OFF_DGS + 16 * h;
    double* sx = sm + SM::OFF_P;         // iterate scratch (phase A only)
    double* ss = sx + 240; double* sz = sx + 400; double* sy = sx + 560;
    int it = 0;
    double n_rx = 0, n_rs = 0, n_re = 0, mu_new = 0;
#pragma unroll 1
    for (; it < A.num_iters; ++it) {
      // ======================= phase A: residuals, weights, Psi, r1~ =====================
      for (int i = lane; i < NZ; i += 32) sx[i] = oX[i];
      for (int i = lane; i < NINEQ; i += 32) { ss[i] = oS[i]; sz[i] = oZ[i]; }
      for (int i = lane; i < NEQ; i += 32) sy[i] = oY[i];
      w.sync();
      double a_sz = 0, a_rs = 0, a_rx = 0, a_re = 0;
#pragma unroll 1
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        stage(k);
        const double* s = sb(h);
        double* ps = psi(k);
        {   // inequality row j = r
          const int ft = r >> 3, jj = r & 7, j = 16 * k + r;
          const double gb = s[SB_G + 12 * ft + 4 + jj];
          double ga = 0.0; int ca = 3 * ft + 2;
          if (jj < 4) { ga = s[SB_G + 12 * ft + jj]; ca = 3 * ft + (jj >> 1); }
          else if (jj < 6) { ga = s[SB_G + 24 + 2 * ft + jj - 4]; ca = 7 + 3 * ft; }
          const double* u = sx + 120 + 12 * k;
          const double sv = ss[j], zv = sz[j];
          const double rs = ga * u[ca] + gb * u[3 * ft + 2] + sv - ldg(gH + j);
          const double si = rcp(sv);
          const double W = si * zv + IP_DELTA;
          const double wi = rcp(W);
          const double D = rcp(wi + IP_DELTA);
          const double r2 = -(si * (sv * zv));
          const double Dg = D * (r2 * wi + rs);
          ws[WK_D + j] = D; ws[WK_WI + j] = wi; ws[WK_DG + j] = Dg; ws[WK_R2 + j] = r2;
          dgs[r] = Dg;
          a_sz += sv * zv; a_rs += rs * rs;
          // Psi: pairs (0,1)->fx (2,3)->fy (4,5)->my ; (6,7) -> fz only
          const double pD = w.shfl_xor(D, 1), pga = w.shfl_xor(ga, 1), pgb = w.shfl_xor(gb, 1);
          double contrib, ia = 0.0, tq = 0.0;
          if (jj < 6) {
            const double eps = s[SB_HU + ca] + IP_BETA;
            const double aa = eps + ga * ga * D + pga * pga * pD;
            const double bb = ga * gb * D + pga * pgb * pD;
            const double cr = gb * pga - pgb * ga;
            ia = rcp(aa);
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
          if (jj == 6) ps[7 * ft + 6] = rcp(sigma);
        }
        w.sync();
        {   // input column c = rr : rx_u and r1~_u
          const double* u = sx + 120 + 12 * k;
          const double* yk = sy + 12 * k;
          double acc = s[SB_HU + rr] * u[rr] + ldg(gF + 120 + 12 * k + rr);
          double acc2 = 0.0;
#pragma unroll
          for (int i = 0; i < 12; i += 2) { acc += s[SB_B + 14 * i + rr] * yk[i]; acc2 += s[SB_B + 14 * (i + 1) + rr] * yk[i + 1]; }
          acc += acc2;
          if (r == 6) acc += s[SB_MX] * sy[120 + 2 * k];
          if (r == 9) acc += s[SB_MX + 1] * sy[121 + 2 * k];
          double gdg, gz;
          gather_col(sz + 16 * k, gdg, gz);
          acc += gz;
          if (r < 12) { a_rx += acc * acc; vx[120 + 12 * k + rr] = -acc - gdg; }
        }
        if (r < 12) {   // state column c = r and dynamics row r
          const int c = r;
          const double* u = sx + 120 + 12 * k;
          const double* yk = sy + 12 * k;
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
          double Br[12];
          ld12(s + SB_B + 14 * r, Br);
          double re = s[SB_E + r] * sx[12 * k + r] - ldg(gB + 12 * k + r), re2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < 12; cc += 2) { re += Br[cc] * u[cc]; re2 += Br[cc + 1] * u[cc + 1]; }
          re += re2;
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

      // ============ two solves with one factorisation: pass 0 affine, pass 1 corrector ======
      double a_pri_aff = 1.0, a_dual_aff = 1.0, smu = 0.0, mn_s = 1e300, mn_z = 1e300;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {   // corrector: r1x = r4 = 0 (build_rho of stage k also reads stage k - 1)
          for (int i = lane; i < NEQ; i += 32) vy[i] = 0.0;
          for (int i = lane; i < 120; i += 32) vx[i] = 0.0;
          w.sync();
        }
#pragma unroll 1
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t;
          stage(k);
          if (pass == 1) {
            // corrector right-hand side (sparse_pdipm_solver.py:478-479): r1 = r3 = r4 = 0,
            // r2 = -(s z + ds_aff dz_aff - sigma mu) / s
            const int j = 16 * k + r;
            const double sv = oS[j], zv = oZ[j];
            const double r2 = -(rcp(sv) * (sv * zv + ws[WK_DS + j] * ws[WK_DZ + j] - smu));
            const double Dg = ws[WK_D + j] * (r2 * ws[WK_WI + j]);
            ws[WK_R2 + j] = r2; ws[WK_DG + j] = Dg;
            dgs[r] = Dg;
            w.sync();
            double gdg, gz;
            gather_col(nullptr, gdg, gz);
            if (r < 12) vx[120 + 12 * k + rr] = -gdg;
            w.sync();
          }
          build_rho(k);
        }
        if (pass == 0) {
          SRBD_TRACE("rho_aff", vy, 140);
          factor();
          SRBD_TRACE("P", sm + SM::OFF_P, 780);
        }
        solve_thomas();
        if (pass == 0) SRBD_TRACE("dyd_aff", vy, 120);
        // ---- recover directions, ratio tests ---------------------------------------------
        mn_s = 1e300; mn_z = 1e300;   // min over ALL 160 candidates (1.0 where dv >= 0): may exceed 1
#pragma unroll 1
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t;
          stage(k);
          double ds, dz;
          recover(k, ds, dz);
          const int j = 16 * k + r;
          if (pass == 1) { ds += ws[WK_DS + j]; dz += ws[WK_DZ + j]; }
          ws[WK_DS + j] = ds; ws[WK_DZ + j] = dz;      // affine (pass 0) / combined (pass 1)
          mn_s = fmin(mn_s, ratio(oS[j], ds));
          mn_z = fmin(mn_z, ratio(oZ[j], dz));
          w.sync();
        }
        if (pass == 0) {
          for (int i = lane; i < NZ; i += 32) ws[WK_DX + i] = vx[i];
          for (int i = lane; i < NEQ; i += 32) ws[WK_DY + i] = vy[i];
          SRBD_TRACE("dx_aff", vx, 240);
          SRBD_TRACE("dy_aff", vy, 140);
          a_pri_aff = step_len(warp_min(mn_s));
          a_dual_aff = step_len(warp_min(mn_z));
          double a_mu = 0.0;
#pragma unroll 1
          for (int t = 0; t < 5; ++t) {
            const int j = 16 * (h ? 9 - t : t) + r;
            a_mu += (oS[j] + a_pri_aff * ws[WK_DS + j]) * (oZ[j] + a_dual_aff * ws[WK_DZ + j]);
          }
          const double mu_aff = warp_sum(a_mu) / (double)NINEQ;
          const double sg = mu_aff * rcp(mu);
          smu = sg * sg * sg * mu;
          w.sync();
        }
      }
      const double a_pri = step_len(warp_min(mn_s));
      const double a_dual = step_len(warp_min(mn_z));
#ifdef SRBD_HOST_EMUL
      { double dbg[8] = {a_pri_aff, a_dual_aff, smu, mu, a_pri, a_dual, 0.0, 0.0}; SRBD_TRACE("scalars", dbg, 8); }
      SRBD_TRACE("dx_cor", vx, 240);
      SRBD_TRACE("dy_cor", vy, 140);
#endif
      // ======================= update ====================================================
      double a_new = 0.0;
      for (int i = lane; i < NZ; i += 32) oX[i] = oX[i] + a_pri * (ws[WK_DX + i] + vx[i]);
      for (int i = lane; i < NEQ; i += 32) oY[i] = oY[i] + a_dual * (ws[WK_DY + i] + vy[i]);
#pragma unroll 1
      for (int t = 0; t < 5; ++t) {
        const int j = 16 * (h ? 9 - t : t) + r;
        const double sn = fmax(oS[j] + a_pri * ws[WK_DS + j], 1e-8);
        const double zn = fmax(oZ[j] + a_dual * ws[WK_DZ + j], 1e-8);
        oS[j] = sn; oZ[j] = zn;
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
