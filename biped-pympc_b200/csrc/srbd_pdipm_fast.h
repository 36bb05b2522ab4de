// PDIPM Newton iterations for one environment per warp -- the lean fast path for
// stage-invariant problem data with unit coefficients on x_{k+1} and on the Mx rows, i.e.
// for everything the reference's `qp_former` produces (srbd_constraints.py:120-135).
//
// Same mathematics as srbd_pdipm_core.h (exact block elimination of the reference KKT of
// biped_pympc/casadi/sparse_pdipm_solver.py:406-448, twisted block-Thomas over ten 12x12
// Schur blocks); what differs is where the data lives and how many instructions and
// shared-memory wavefronts one Newton step costs:
//
//   * per-lane registers hold the lane's row and column of Bbar, its coupling rows and the
//     row of the pivot-block inverse it produced in the previous elimination round;
//   * pivot blocks are inverted by a symmetric sweep: per pivot one 8-byte store per lane
//     (all-gather of the pivot column), one warp barrier, six 16-byte broadcast loads and
//     eleven FMAs;
//   * the rows of P_k a neighbour round needs are handed over in a 72-double window instead
//     of being re-gathered from the packed storage;
//   * the per-inequality-row data (s, z, the two direction slots, D w^-1) is lane-private and
//     lives in the L2-resident work buffer; only D stays in shared memory;
//   * fp64 path: the lanes' rows of the pivot-block inverses in tensor memory, 1 742 doubles of shared memory per warp
//     -> 16 warps (environments) per SM, persistent; fp32 mode: ten packed blocks in shared memory, 1 908 doubles, 14 warps.
//
// Compiled for the GPU (srbd_pdipm.cu) and, for tests only, for the host under the warp
// emulator of srbd_warp.h.
#pragma once

#include <type_traits>

#include "srbd_pdipm_core.h"
#include "srbd_qp_table.h"

namespace srbd {

// ---- stage data (doubles) ---------------------------------------------------------------
enum : int {
  FB_B = 0,        // Bbar dense [12][14]
  FB_AB = 168,     // Abar compact: diag[12] | R[3][3] (rows 0..2, cols 6..8) | pv[3]
  FB_G = 192,      // G values of a stage [28]
  FB_HX = 220,     // Q_val of a state block [12]
  FB_HU = 232,     // Q_val of an input block [12]
  FB_RAW = 244,    // ---- end of the gathered part; then E[12], Mx[2] are gathered and checked
  FB_PHX = 244,    // 1 / (hx + beta)
  FB_DER = 256,    // psi_mz[2], psi_mx'[2], 1/mx_den[2], eps_mx[2]
  FB_CM = 264,     // 2 x 36 coupling coefficient slots
  FB_SX = 336,     // A phix A' (state part of S_kk without its diagonal term), packed lower (78)
  FB_SIZE = 414,
  FB_NSLOT = FB_RAW + 14
};

// kTm: the fp64 path, whose pivot-block rows live in tensor memory -- only the first 644 doubles of the P region are
// needed (phase A's scratch [0, 500) with the merge block's full rows inside, the full rows of the state part of S_kk)
// and the packed state part at the end of the stage data is not: 1 742 doubles per env, 16 envs per SM.  The fp32 mode
// keeps the ten packed blocks: 1 908 doubles, 14 envs per SM.
template <bool kTm>
struct FastSmemT {
  static constexpr int OFF_P = 0;          // 10 x 78 packed lower; scratch for iterates in phase A
  // fp64 path (the blocks' rows are in tensor memory; phase A's scratch ends at 500, the fused set-up's at 345): the two
  // symmetric blocks that are still exchanged through shared memory are kept as FULL rows [12][12] -- row loads / stores
  // with immediate offsets, no packed-offset registers at all
  // (row stride FS = 14 doubles, like Bbar: the twelve lanes of a half-warp then read their rows with two wavefronts
  //  per 16-byte load instead of three at stride 12 -- profiles/smem_bcast_r1_wavefronts.csv)
  static constexpr int FS = 14;
  static constexpr int OFF_P5F = 332;      // merge block 5: parked S_55 - C_6' P_6 C_6 (round 4 -> 5), then N_5  [12][FS]
  static constexpr int OFF_SXF = 500;      // state part of S_kk including its diagonal term                     [12][FS]
  static constexpr int OFF_VX = kTm ? 668 : 780;   // 240  r1~ -> dx
  static constexpr int OFF_VY = OFF_VX + 240;      // 140  re -> rho -> t -> dy
  static constexpr int OFF_PSI = OFF_VY + 140;     // 10 x 14
  static constexpr int OFF_BC = OFF_PSI + 140;     // 2 x 24 broadcast scratch per half
  static constexpr int OFF_TR = OFF_BC + 48;       // 2 x 32: D*g and z of the stage being processed ...
  static constexpr int HS = kTm ? 14 : 12;         // row stride of the hand-over window (fp64: 14, as for the full rows)
  static constexpr int OFF_PR = OFF_TR;            // ... or (factor only) 2 x 6 handed-over rows of P
  static constexpr int OFF_FB = OFF_PR + 2 * 6 * HS;
  static constexpr int OFF_TM = OFF_FB + (kTm ? FB_SX : FB_SIZE);   // 32-bit tensor-memory address of the warp's slot 0
                                                    // (re-read at every use: a register held across the solve costs spills)
  static constexpr int SIZE = OFF_TM + 2;           // even: every warp's region stays 16-byte aligned
};
static_assert(FastSmemT<true>::SIZE % 2 == 0 && FastSmemT<false>::SIZE % 2 == 0, "16-byte alignment of the warps' regions");
// lane-private per-row scratch in the work buffer
enum : int { FW_DX = 0, FW_DY = 240, FW_S = 380, FW_Z = 540, FW_Q = 700,
             // fused tick only: f of the states, c_d | c_d + A_d x0, F_max * contact
             FW_FX = 860, FW_CD = 980, FW_FC = 1004,
             // working copies of s, z and the bound h of the lane's rows: with ds, dz and D/W above they are addressed
             // as one per-row pointer + immediate (the caller's s / z buffers are read once and written once per call)
             FW_SV = 1024, FW_ZV = 1184, FW_H = 1344,
             FW_D = 1504,    // D = W / (1 + delta W) of the row (written in phase A, read when the step is recovered)
             // fp32 pivot blocks with refinement only: re of the iteration (r4 of the affine solve), the unrefined dx
             FW_RE = 1664, FW_RX = 1792, FW_END = 2032 };
static_assert(FW_END <= PDIPM_WORK_DOUBLES_FAST, "work buffer too small");

// Packed storage of the symmetric 12x12 blocks (P_k, A phix A'): row i keeps its entries (i, 0..i) contiguous, as in the
// triangular layout pk() of srbd_pdipm_core.h, but the rows are laid out in the order 7 0 1 3 6 11 9 4 8 2 5 10.  When
// the twelve lanes of a half-warp read column c of a block -- (c, 0..c) from row c, (i, c) from the rows i > c -- the
// twelve 8-byte words then fall into twelve different bank pairs for every c: 2 wavefronts per 64-bit load instead of
// the 3.5 of the triangular order (sum over c of the worst bank multiplicity: 12 against 21; found by annealing over row
// orders, tools/pack_layout.py).  The address arithmetic keeps the triangular layout's form -- own row: base(rr) + m,
// other rows: constant + rr -- so it costs no registers.
SRBD_DEV int pkf_row(int i) {
  constexpr int kRow[12] = {8, 9, 58, 11, 44, 61, 15, 0, 49, 34, 67, 22};
  return kRow[i];
}
SRBD_DEV int pkf(int i, int j) { return i >= j ? pkf_row(i) + j : pkf_row(j) + i; }

// host: gather table of the fast path from the table of srbd_stage_table.h
inline void build_fast_slots(const StageSlot* sb, StageSlot* fs) {
  for (int i = 0; i < FB_NSLOT; ++i) { fs[i].kind = ST_ZERO; fs[i].rel = 0; }
  for (int i = 0; i < 168; ++i) fs[FB_B + i] = sb[SB_B + i];
  for (int i = 0; i < 24; ++i) fs[FB_AB + i] = sb[SB_AP + i];
  for (int i = 0; i < 28; ++i) fs[FB_G + i] = sb[SB_G + i];
  for (int i = 0; i < 12; ++i) fs[FB_HX + i] = sb[SB_HX + i];
  for (int i = 0; i < 12; ++i) fs[FB_HU + i] = sb[SB_HU + i];
  for (int i = 0; i < 12; ++i) fs[FB_RAW + i] = sb[SB_E + i];
  for (int i = 0; i < 2; ++i) fs[FB_RAW + 12 + i] = sb[SB_MX + i];
}

// entry q*3+i of the (2,6) body-frame wrench of one env: rows [-R' f_L, -R' m_L ; -R' f_R, -R' m_R]
// with m_x zeroed and the cast to fp32 done before the rotation, exactly as
// mpc_controller_cusadi.py:186-203.  R: row-major fp32 state-estimate rotation.
SRBD_DEV float wrench_entry(const double* u0, const float* R, int idx) {
  const int q = idx / 3, i = idx - 3 * q;
  const int base = (q == 0) ? 0 : (q == 1 ? 6 : (q == 2 ? 3 : 9));   // f_L, m_L, f_R, m_R
  float v0 = (float)u0[base], v1 = (float)u0[base + 1], v2 = (float)u0[base + 2];
  if (q == 1 || q == 3) v0 = 0.f;
  return -(R[0 + i] * v0 + R[3 + i] * v1 + R[6 + i] * v2);
}

// kF32 = false: everything in fp64 (the parity path).  kF32 = true, the optional reduced-precision mode
// (SRBD_PRECISION_F32_FACTOR): the ten Schur pivot blocks are formed, coupled, inverted and stored in fp32 -- the part
// of a Newton step with the matrix-matrix work -- while residuals, right-hand sides, the block-Thomas recurrences and
// the iterates stay in fp64; with `refine` the two solves of an iteration are each followed by one step of iterative
// refinement against the fp64 Schur operator.  Accuracy and speed are measured, not promised: DESIGN.md section 8.
template <bool kF32>
struct PdipmFastT {
  using SM = FastSmemT<!kF32>;
  using R = typename std::conditional<kF32, float, double>::type;   // precision of the pivot blocks
  Warp w;
  double* sm;
  const StageSlot* slots;
  // Register diet: 13-16 warps per SM put four on one scheduler (16 K registers), i.e. 128 registers per thread.  Only what
  // the elimination rounds need every time stays in registers; everything else is re-read from
  // shared memory (short latency) instead of being spilled to local memory (L2 latency: the 28 KB
  // of L1 left beside 228 KB of shared memory cannot hold the spills of 448 threads).
  const PdipmArgs* pa;
  long long e;
  int lane, h, r, rr;
  int prow[kF32 ? 12 : 1];      // packed offsets of row rr, in BYTES of the block's precision (one add per access);
                                //   fp32 mode only -- the fp64 path has no packed block left
  int bs0, bs1;                 // sparse rows of column rr of Bbar
  double c0[4], c1[4];          // row rr of C = Abar diag(phix)  /  row rr of C'
  int i0[3], i1[3];             // their off-diagonal column indices
  double phx_rr;                // 1/(hx + beta) of state rr
  int goa, gob, ca, cz;         // inequality row r: G(r, ca) = fb[goa], G(r, cz) = fb[gob]
  double gva, gvb;              // ... and the two coefficients themselves (constant over the call; 20 uses per iteration:
                                //   two registers each beat re-reading them from shared memory since the packed-offset
                                //   registers are gone, -1.2 %)
  int psk, psi_i, psf;          // role of input column rr in Psi: kind, index, foot
  int doff, moff;               // fb offsets of its constant Psi diagonal / 1/mx_den (or of a zero)
  int g_src;                    // lane that ends up with sum_j G[j, rr] v_j in gather_rows (-1: the column has no rows)
  enum : int { FB_ZERO = FB_B + 12 };   // a structurally zero slot (pad column of Bbar)

  // ---- the lane's rows of the pivot-block inverses live in TENSOR MEMORY (fp64 path) ---------------------------------
  // A lane only ever needs ITS OWN row of N_k = -P_k of the blocks its half-warp eliminated (mat-vecs of the block
  // solves, own row in the coupling): lane-private data, which the shared-memory pipe -- the resource that binds this
  // kernel -- serves worst (packed symmetric storage: twelve 8-byte loads with twelve different addresses per half-warp
  // and an address computation each).  Blackwell's tensor memory is 128 lanes x 512 32-bit columns per SM with its own
  // datapath: tcgen05.st / tcgen05.ld move N consecutive columns of every lane of a warp's 32-lane sub-partition in ONE
  // instruction (measured, tools/tmem_probe.cu: a 12-double row per lane out and back in 234 cycles with 14 warps per SM,
  // 1 299 cycles through shared memory).  Warp w owns lanes 32 (w % 4) .. +31 and columns 128 (w / 4) .. +119: five slots
  // of 24 columns; slot t holds the row of the block the lane's half eliminated in round t (top half: block t, bottom
  // half: block 9 - t), so both halves address the same columns in every phase.  Block 5 (the merge block) stays packed
  // in shared memory.  The fp32 mode and the host emulator keep the slots in ordinary memory.
  static constexpr bool kTm = !kF32;
#if defined(__CUDACC__) && !defined(SRBD_HOST_EMUL)
  // tensor-memory address of slot 0 of this warp (lane field = its sub-partition), written by the kernel before run()
  SRBD_DEV unsigned tm_base() const { return *reinterpret_cast<const volatile unsigned*>(sm + SM::OFF_TM); }
  SRBD_DEV void tm_st(int slot, const double v[12]) const {
    const unsigned addr = tm_base() + 24u * (unsigned)slot;
    unsigned q[24];
#pragma unroll
    for (int i = 0; i < 12; ++i) { q[2 * i] = (unsigned)__double2loint(v[i]); q[2 * i + 1] = (unsigned)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]),
                    "r"(q[8]), "r"(q[9]), "r"(q[10]), "r"(q[11]), "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(addr + 16u), "r"(q[16]), "r"(q[17]), "r"(q[18]), "r"(q[19]), "r"(q[20]), "r"(q[21]), "r"(q[22]), "r"(q[23]) : "memory");
  }
  // the load is issued early (its latency then hides behind the exchange of the right-hand side) and completed by
  // tm_ld_finish right before the first use; stores are only waited for when the next load is issued
  SRBD_DEV void tm_ld_issue(int slot, unsigned q[24]) const {
    const unsigned addr = tm_base() + 24u * (unsigned)slot;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                   "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]) : "r"(addr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]) : "r"(addr + 16u) : "memory");
  }
  SRBD_DEV void tm_ld_finish(unsigned q[24], double v[12]) const {
    // the registers are operands of the wait so that the compiler orders every use after it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]),
                   "+r"(q[8]), "+r"(q[9]), "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15]),
                   "+r"(q[16]), "+r"(q[17]), "+r"(q[18]), "+r"(q[19]), "+r"(q[20]), "+r"(q[21]), "+r"(q[22]), "+r"(q[23]) :: "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = __hiloint2double((int)q[2 * i + 1], (int)q[2 * i]);
  }
  SRBD_DEV void tm_ld(int slot, double v[12]) const {
    unsigned q[24];
    tm_ld_issue(slot, q);
    tm_ld_finish(q, v);
  }
#else
  double* tme;                  // host emulation: 32 lanes x 5 slots x 12 doubles of lane-private memory
  SRBD_DEV void tm_st(int slot, const double v[12]) const { for (int i = 0; i < 12; ++i) tme[60 * lane + 12 * slot + i] = v[i]; }
  SRBD_DEV void tm_ld(int slot, double v[12]) const { for (int i = 0; i < 12; ++i) v[i] = tme[60 * lane + 12 * slot + i]; }
  SRBD_DEV void tm_ld_issue(int slot, unsigned q[24]) const { q[0] = (unsigned)slot; }
  SRBD_DEV void tm_ld_finish(unsigned q[24], double v[12]) const { tm_ld((int)q[0], v); }
#endif
  // -(row rr of the block whose load was issued into q) . v
  SRBD_DEV double trow_dot(unsigned q[24], const double v[12]) const {
    double nr[12];
    tm_ld_finish(q, nr);
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
    for (int m = 0; m < 12; m += 3) { q0 = fma(-nr[m], v[m], q0); q1 = fma(-nr[m + 1], v[m + 1], q1); q2 = fma(-nr[m + 2], v[m + 2], q2); }
    return q0 + q1 + q2;
  }

  SRBD_DEV const double* gin(int i, int nnz) const { return pa->in[i] + e * nnz; }
  SRBD_DEV double* gout(int i, int nnz) const { return pa->out[i] + e * nnz; }
  SRBD_DEV double* gws() const { return pa->work + e * (long long)pa->work_stride; }

  SRBD_DEV R* P(int k) const { return reinterpret_cast<R*>(sm + SM::OFF_P) + 78 * k; }
  // entry (rr, m) of a packed block: base + byte offset (written as `Pm[prow[m]]` with element offsets the compiler forms
  // (78 k + prow[m]) * 8 + base per access: two multiply-adds per element instead of one add)
  template <class T> SRBD_DEV T& pe(T* Pm, int m) const {
    using C = typename std::conditional<std::is_const<T>::value, const char, char>::type;
    return *reinterpret_cast<T*>(reinterpret_cast<C*>(Pm) + prow[m]);
  }
  SRBD_DEV double* fb() const { return sm + SM::OFF_FB; }
  SRBD_DEV double* bc() const { return sm + SM::OFF_BC + 24 * h; }
  SRBD_DEV double* tr() const { return sm + SM::OFF_TR + 32 * h; }
  SRBD_DEV R* pr(int hh) const { return reinterpret_cast<R*>(sm + SM::OFF_PR) + 6 * SM::HS * hh; }
  SRBD_DEV R* bcr() const { return reinterpret_cast<R*>(sm + SM::OFF_BC + 24 * h); }   // broadcast scratch in block precision
  SRBD_DEV double* psi(int k) const { return sm + SM::OFF_PSI + 14 * k; }

  // ------------------------------------------------------------------------------------
  // one-time set-up: gather the stage data, derive the iteration-invariant quantities
  // returns nonzero when E or the Mx coefficients are not exactly 1
  // ------------------------------------------------------------------------------------
  SRBD_DEV int setup() {
    double* f = fb();
    int bad = 0;
    if (pa->fused) gather_fused();
    else {
    const double* gQ = gin(IP_Q, NNZ_H); const double* gG = gin(IP_G, NNZ_G); const double* gA = gin(IP_A, NNZ_A);
    for (int i = lane; i < FB_NSLOT; i += 32) {
      const StageSlot sl = slots[i];
      double v = 0.0;
      switch (sl.kind) {   // stage k = 1 stands for every stage
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
    if (pa->inv_flag) {   // the probe of srbd_misc.cu folded into the call: every stage carries the values of stage 1
      int diff = 0;
      for (int i = lane; i < 36 * 8; i += 32) diff |= (ldg(gA + 36 + i) != ldg(gA + i % 36));
      for (int c = lane; c < 12; c += 32) diff |= (ldg(gA + 324 + c) != 1.0) | (ldg(gA + a_xcol_off(c)) != 1.0);
      for (int i = lane; i < 86 * 9; i += 32) diff |= (ldg(gA + 336 + 86 + i) != ldg(gA + 336 + i % 86));
      if (lane < 2) diff |= (ldg(gA + 336 + a_ucol_off(6 + 3 * lane) + A_UCOL_MX_POS) != 1.0);
      for (int i = lane; i < 28 * 9; i += 32) diff |= (ldg(gG + 28 + i) != ldg(gG + i % 28));
      for (int i = lane; i < 108; i += 32) diff |= (ldg(gQ + 12 + i) != ldg(gQ + i % 12)) | (ldg(gQ + 132 + i) != ldg(gQ + 120 + i % 12));
      if (diff) bad = 1;
    }
    for (int m = 16; m >= 1; m >>= 1) bad |= w.shfl_i(bad, lane ^ m);
    if (bad && pa->inv_flag && lane == 0) *pa->inv_flag = 1;
    w.sync();
    }
    const int q = rr;
    f[FB_PHX + q] = rcp(f[FB_HX + q] + IP_BETA);
    if (r < 2) {
      f[FB_DER + r] = rcp(f[FB_HU + 8 + 3 * r] + IP_BETA);             // mz: plain diagonal
    } else if (r < 4) {                                                 // mx: with its equality row
      const int i = r - 2;
      const double eps = f[FB_HU + 6 + 3 * i] + IP_BETA;
      const double iden = rcp(1.0 + IP_DELTA * eps);
      f[FB_DER + 2 + i] = IP_DELTA * iden;
      f[FB_DER + 4 + i] = iden;
      f[FB_DER + 6 + i] = eps;
    }
    w.sync();
    // ---- per-lane registers ----------------------------------------------------------
    // (the lane's own row start is looked up once: kRow[] with a run-time index is a chain of selects)
    const int myrow = pkf_row(rr);
    if constexpr (kF32) {
#pragma unroll
      for (int m = 0; m < 12; ++m) prow[m] = (m <= rr ? myrow + m : pkf_row(m) + rr) * (int)sizeof(R);
    }
    bs0 = 3 + rr % 3; bs1 = 9 + rr % 3;
    phx_rr = f[FB_PHX + rr];
    {
      const double* A = f + FB_AB; const double* ph = f + FB_PHX;
#pragma unroll
      for (int a = 0; a < 4; ++a) { c0[a] = 0.0; c1[a] = 0.0; }
#pragma unroll
      // slots with a zero coefficient point where the lanes with a real entry point (rows 6 + a / a):
      // every distinct address of a lane-varying shared-memory load costs wavefronts
      // (profiles/smem_bcast_r1_wavefronts.csv)
      for (int a = 0; a < 3; ++a) { i0[a] = 6 + a; i1[a] = a; }
      c0[0] = A[q] * ph[q];
      c1[0] = A[q] * ph[q];
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { i0[j] = 6 + j; c0[1 + j] = A[12 + 3 * q + j] * ph[6 + j]; }
      } else if (q < 6) { i0[0] = q + 6; c0[1] = A[21 + q - 3] * ph[q + 6]; }
      else if (q < 9) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { i1[i] = i; c1[1 + i] = A[12 + 3 * i + (q - 6)] * ph[q]; }
      } else { i1[0] = q - 6; c1[1] = A[21 + q - 9] * ph[q]; }
      // 36 static slots per type: [0..11] diagonal, [12..20] rows 0-2 x cols 6-8, [21..23] rows
      // 3-5 (col r+6), [24..32] rows 6-8 x cols 0-2, [33..35] rows 9-11 (col r-6)
      double* m0 = f + FB_CM; double* m1 = f + FB_CM + 36;
      m0[q] = c0[0]; m1[q] = c1[0];
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { m0[12 + 3 * q + j] = c0[1 + j]; m1[12 + 3 * q + j] = 0.0; }
      } else if (q < 6) { m0[21 + q - 3] = c0[1]; m1[21 + q - 3] = 0.0; }
      else if (q < 9) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { m0[24 + 3 * (q - 6) + i] = 0.0; m1[24 + 3 * (q - 6) + i] = c1[1 + i]; }
      } else { m0[33 + q - 9] = 0.0; m1[33 + q - 9] = c1[1]; }
      // A phix A' (row q, lower part)
      for (int j = 0; j <= q; ++j) {
        double acc = c0[0] * PdipmWarp<true>::abar(A, j, q);
#pragma unroll
        for (int a = 0; a < 3; ++a) acc += c0[1 + a] * PdipmWarp<true>::abar(A, j, i0[a]);
        if constexpr (kTm) {            // the same value at both mirror positions: exactly symmetric, as the packed form is
          double* sxf = sm + SM::OFF_SXF;
          const double v = (j == q) ? acc + (phx_rr + IP_DELTA) : acc;
          sxf[SM::FS * q + j] = v;
          sxf[SM::FS * j + q] = v;
        } else {
          f[FB_SX + myrow + j] = acc;   // pkf(q, j) for j <= q: the lane's own row is contiguous
        }
      }
    }
    {   // inequality row r
      const int ft = r >> 3, jj = r & 7;
      cz = 3 * ft + 2;
      gob = FB_G + 12 * ft + 4 + jj;
      goa = FB_ZERO; ca = cz;
      if (jj < 4) { goa = FB_G + 12 * ft + jj; ca = 3 * ft + (jj >> 1); }
      else if (jj < 6) { goa = FB_G + 24 + 2 * ft + jj - 4; ca = 7 + 3 * ft; }
    }
    {   // input column rr
      const int c = rr;
      psf = (c < 6) ? c / 3 : (c >= 9 ? 1 : 0);
      doff = FB_ZERO; moff = FB_ZERO; psi_i = 0;
      if (c < 6) {
        psi_i = c % 3;
        if (psi_i == 2) { psk = 1; psi_i = 0; } else psk = 0;
      } else if (c == 7 || c == 10) { psk = 0; psi_i = 2; }
      else if (c == 8 || c == 11) { psk = 2; psi_i = (c == 11); doff = FB_DER + psi_i; }
      else { psk = 3; psi_i = (c == 9); doff = FB_DER + 2 + psi_i; moff = FB_DER + 4 + psi_i; }
      g_src = -1;
      if (c < 6) g_src = (lane & 16) + 8 * (c / 3) + (c % 3 == 2 ? 6 : 2 * (c % 3));
      else if (c == 7 || c == 10) g_src = (lane & 16) + 8 * (c == 10) + 4;
    }
    w.sync();
    gva = f[goa]; gvb = f[gob];
    w.sync();
    return bad;
  }

  // fused tick: the gathered part of the stage data straight from the 17 qp_former inputs (the
  // value table of srbd_qp_table.h is built in the scratch that later holds P), plus the three
  // stage-dependent vectors the residual phase needs, parked in the work buffer
  SRBD_DEV void gather_fused() {
    double* f = fb();
    double* V = sm + SM::OFF_P;           // V_COUNT
    double* Wr = V + 232;                 // RW_COUNT
    // x_ref (and the linearisation point, if given) are fetched before the value table is built: with the inputs in
    // mapped host memory (host entry) their PCIe round trip then overlaps the table's
    double xrv[4], xlv[4];
    {
      const double* xr = pa->qp_in[IN_XREF] + e * 120;
      const double* xl = pa->qp_in[IN_X] ? pa->qp_in[IN_X] + e * 120 : nullptr;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int i = lane + 32 * m;
        xrv[m] = i < 120 ? xr[i] : 0.0;
        xlv[m] = (xl && i < 120) ? xl[i] : 1.0;
      }
    }
    qp_value_table(pa->qp_in, e, lane, 32, w, V, Wr);
    for (int i = lane; i < 168; i += 32) { const int rw = i / 14, c = i - 14 * rw; f[FB_B + i] = c < 12 ? V[V_NBD + 12 * rw + c] : 0.0; }
    if (lane < 12) { f[FB_AB + lane] = V[V_NEG1]; f[FB_HX + lane] = V[V_Q + lane]; f[FB_HU + lane] = V[V_R + lane]; }
    if (lane < 9) f[FB_AB + 12 + lane] = V[V_NADTR + lane];
    if (lane < 3) f[FB_AB + 21 + lane] = V[V_NADT];
    if (lane < 28) {   // G values of a stage: two feet x 12, then the two my columns (srbd_pattern.h)
      const int q = lane < 24 ? lane % 12 : lane - 24;
      int v;
      if (lane < 24) v = (q == 0 || q == 2 || q == 10) ? V_NEG1 : (q == 1 || q == 3 || q == 11) ? V_ONE : q < 8 ? V_NEGMU : (q == 8 ? V_NEGLT : V_NEGLH);
      else v = (q & 1) ? V_ONE : V_NEG1;
      f[FB_G + lane] = V[v];
    }
    double* ws = gws();
    {   // f_x = grad - H z evaluated as the assembly kernel does (srbd_constraints.py:67-69)
      int m12 = lane % 12;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int i = lane + 32 * m;
        const double q = V[V_Q + m12];
        if (i < 120) ws[FW_FX + i] = q * (xlv[m] - xrv[m]) - q * xlv[m];
        m12 += 8; if (m12 >= 12) m12 -= 12;
      }
    }
    if (lane < 24) ws[FW_CD + lane] = V[V_CD + lane];      // c_d (12) | c_d + A_d x0 (12)
    if (lane < 20) ws[FW_FC + lane] = V[V_FC + lane];
    w.sync();
  }
  // f, h, b of inequality row j / column rr of stage k, whichever mode
  SRBD_DEV double h_of(int k, int j) const {
    if (!pa->fused) return ldg(gin(IP_H, NINEQ) + j);
    const int q = j & 7;
    return q == 7 ? gws()[FW_FC + 2 * k + ((j >> 3) & 1)] : 0.0;
  }

  // ------------------------------------------------------------------------------------
  // Psi_k applied to the 12-vector at v (shared memory); returns entry rr
  // ------------------------------------------------------------------------------------
  SRBD_DEV double psi_apply(const double* ps, const double* v) const {
    const double* p = ps + 7 * psf;
    const int fx = 3 * psf;
    const double dot = p[3] * v[fx] + p[4] * v[fx + 1] + p[5] * v[7 + fx] - v[fx + 2];
    const double dg = psk == 0 ? p[psi_i] : fb()[doff];
    const double q = psk == 0 ? p[3 + psi_i] : (psk == 1 ? -1.0 : 0.0);
    return dg * v[rr] + q * (p[6] * dot);
  }
  // sum_j G[j, rr] v_j over the sixteen inequality rows of a stage, lane r holding v_r and its row's two coefficients
  // (ga on column ca, gb on the foot's fz).  The rows of a foot are eight neighbouring lanes: the pairs (0,1) (2,3) (4,5)
  // carry the fx, fy, my columns, all eight the fz column, so two butterflies and one delivery shuffle replace a
  // shared-memory staging of v and eight strided re-reads per lane
  SRBD_DEV double gather_rows(double v, double ga, double gb) const {
    double pa = ga * v, pz = gb * v;
    pa += w.shfl_xor(pa, 1);
    pz += w.shfl_xor(pz, 1);
    pz += w.shfl_xor(pz, 2);
    pz += w.shfl_xor(pz, 4);
    const double got = w.shfl(((r & 7) == 6) ? pz : pa, g_src < 0 ? lane : g_src);
    return g_src < 0 ? 0.0 : got;
  }
  SRBD_DEV double dot12(const double a[12], const double b[12]) const {
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int c = 0; c < 12; c += 2) { q0 = fma(a[c], b[c], q0); q1 = fma(a[c + 1], b[c + 1], q1); }
    return q0 + q1;
  }
  // column rr of Bbar times the 12-vector at v (shared memory)
  SRBD_DEV double bcol_dot(const double* v) const {
    double bc8[8];
    ld_bcol(bc8);
    return bcol_dot(bc8, v);
  }
  // the eight structural non-zeros of column rr of Bbar (kept in registers over a stage loop)
  SRBD_DEV void ld_bcol(double bc8[8]) const {
    const double* B = fb() + FB_B + rr;
    bc8[0] = B[0]; bc8[1] = B[14]; bc8[2] = B[28]; bc8[3] = B[14 * bs0];
    bc8[4] = B[84]; bc8[5] = B[98]; bc8[6] = B[112]; bc8[7] = B[14 * bs1];
  }
  SRBD_DEV double bcol_dot(const double bc8[8], const double* v) const {
    const double q0 = bc8[0] * v[0] + bc8[1] * v[1] + bc8[2] * v[2] + bc8[3] * v[bs0];
    const double q1 = bc8[4] * v[6] + bc8[5] * v[7] + bc8[6] * v[8] + bc8[7] * v[bs1];
    return q0 + q1;
  }
  // row rr of Bbar times the 12-vector v
  SRBD_DEV double brow_dot(const double v[12]) const {
    double Brow[12];
    ld12(fb() + FB_B + 14 * rr, Brow);
    return dot12(Brow, v);
  }

  // ------------------------------------------------------------------------------------
  // rho_k (into vy) from r1~ (vx) and the equality right-hand sides held in vy;
  // cor: corrector right-hand side (r1x = r4 = 0: only the input part of r1~ is non-zero)
  // ------------------------------------------------------------------------------------
  SRBD_DEV void build_rho(int k, bool cor, const double Brow[12]) const {
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* b = bc();
    {
      double v = psi_apply(psi(k), vx + 120 + 12 * k);
      if (!cor) v = fma(fb()[moff], -vy[120 + 2 * k + psi_i], v);    // (1/den) r4_mx
      b[rr] = v;
    }
    w.sync();
    {
      double bv[12];
      ld12(b, bv);
      double rho = dot12(Brow, bv);
      if (!cor) {
        rho += phx_rr * vx[12 * k + rr] + vy[12 * k + rr];
        if (k >= 1) {
          const double* xp = vx + 12 * (k - 1);
          rho += c0[0] * xp[rr] + c0[1] * xp[i0[0]] + c0[2] * xp[i0[1]] + c0[3] * xp[i0[2]];
        }
      }
      if (r < 12) vy[12 * k + rr] = rho;
    }
    w.sync();
  }

  // ------------------------------------------------------------------------------------
  // row rr of S_kk into a[12]
  // ------------------------------------------------------------------------------------
  SRBD_DEV void form_row(int k, R a[12]) const {
    const double* f = fb();
    const double* ps = psi(k);
    const double* der = f + FB_DER;
    R* b = bcr();
    R wv[12];
    R sL, sR;
    {
      double Bd[12];
      ld12(f + FB_B + 14 * rr, Bd);
      R Brow[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) Brow[i] = (R)Bd[i];
      R psv[14];
      const dbl2* q = reinterpret_cast<const dbl2*>(ps);
#pragma unroll
      for (int i = 0; i < 7; ++i) { const dbl2 t = q[i]; psv[2 * i] = (R)t.x; psv[2 * i + 1] = (R)t.y; }
      const R pL = psv[3] * Brow[0] + psv[4] * Brow[1] + psv[5] * Brow[7] - Brow[2];
      const R pR = psv[10] * Brow[3] + psv[11] * Brow[4] + psv[12] * Brow[10] - Brow[5];
      b[rr] = pL; b[12 + rr] = pR;
      wv[0] = psv[0] * Brow[0]; wv[1] = psv[1] * Brow[1]; wv[2] = (R)0;
      wv[3] = psv[7] * Brow[3]; wv[4] = psv[8] * Brow[4]; wv[5] = (R)0;
      wv[6] = (R)der[2] * Brow[6]; wv[7] = psv[2] * Brow[7]; wv[8] = (R)der[0] * Brow[8];
      wv[9] = (R)der[3] * Brow[9]; wv[10] = psv[9] * Brow[10]; wv[11] = (R)der[1] * Brow[11];
      sL = psv[6] * pL; sR = psv[13] * pR;
    }
    // state part: delta + phix on the diagonal, A phix A' (absent for k = 0: x_0 is a parameter)
    const R fk = (k == 0) ? (R)0 : (R)1;
    const R dg = (R)(phx_rr + IP_DELTA);
    if constexpr (kTm) {
      ld12(sm + SM::OFF_SXF + SM::FS * rr, a);
      if (k == 0) {      // x_0 is a parameter: only the diagonal term (round 0 of the top half; no other round enters)
#pragma unroll
        for (int j = 0; j < 12; ++j) a[j] = (rr == j) ? dg : (R)0;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) a[j] = fmaR(fk, (R)*reinterpret_cast<const double*>(reinterpret_cast<const char*>(f + FB_SX) + 2 * prow[j]), (rr == j) ? dg : (R)0);
    }
    w.sync();
    {
      R pl[12];
      ld12(b, pl);
#pragma unroll
      for (int j = 0; j < 12; ++j) a[j] = fmaR(sL, pl[j], a[j]);
      ld12(b + 12, pl);
#pragma unroll
      for (int j = 0; j < 12; ++j) a[j] = fmaR(sR, pl[j], a[j]);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const double* Bj = f + FB_B + 14 * j;
      if (j < 3 || (j >= 6 && j < 9)) {          // dense rows of Bbar (euler, omega)
        double bj[12];
        ld12(Bj, bj);
        R q0 = (R)0, q1 = (R)0;
#pragma unroll
        for (int c = 0; c < 12; c += 2) {
          if (c != 2) q0 = fmaR(wv[c], (R)bj[c], q0);
          if (c + 1 != 5) q1 = fmaR(wv[c + 1], (R)bj[c + 1], q1);
        }
        a[j] += q0 + q1;
      } else {                                    // position / velocity rows: two entries
        const int c1_ = (j < 6) ? j - 3 : j - 9, c2_ = c1_ + 3;
        a[j] = fmaR(wv[c1_], (R)Bj[c1_], fmaR(wv[c2_], (R)Bj[c2_], a[j]));
      }
    }
  }

  // n -= M P M' = n + M N M' with M the coupling of `type`; own = row rr of N = -P, hand = handed-over rows of N
  SRBD_DEV void couple(R n[12], const R own[12], const R* hand, int type) const {
    const R cv0 = (R)(type ? c1[0] : c0[0]), cv1 = (R)(type ? c1[1] : c0[1]);
    const R cv2 = (R)(type ? c1[2] : c0[2]), cv3 = (R)(type ? c1[3] : c0[3]);
    const int base = type ? 0 : 6;
    int s1 = (type ? i1[0] : i0[0]) - base, s2 = (type ? i1[1] : i0[1]) - base, s3 = (type ? i1[2] : i0[2]) - base;
    s1 = s1 < 0 ? 0 : (s1 > 5 ? 5 : s1); s2 = s2 < 0 ? 0 : (s2 > 5 ? 5 : s2); s3 = s3 < 0 ? 0 : (s3 > 5 ? 5 : s3);
    R U[12];
    {
      R t1[12];
      ld12(hand + SM::HS * s1, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fmaR(cv1, t1[m], cv0 * own[m]);
      ld12(hand + SM::HS * s2, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fmaR(cv2, t1[m], U[m]);
      ld12(hand + SM::HS * s3, t1);
#pragma unroll
      for (int m = 0; m < 12; ++m) U[m] = fmaR(cv3, t1[m], U[m]);
    }
    const double* cm = fb() + FB_CM + 36 * type;
#pragma unroll
    for (int cc = 0; cc < 12; ++cc) {
      R o = U[cc] * (R)cm[cc];
      if (cc < 3) o += U[6] * (R)cm[12 + 3 * cc] + U[7] * (R)cm[12 + 3 * cc + 1] + U[8] * (R)cm[12 + 3 * cc + 2];
      else if (cc < 6) o += U[cc + 6] * (R)cm[21 + cc - 3];
      else if (cc < 9) o += U[0] * (R)cm[24 + 3 * (cc - 6)] + U[1] * (R)cm[24 + 3 * (cc - 6) + 1] + U[2] * (R)cm[24 + 3 * (cc - 6) + 2];
      else o += U[cc - 6] * (R)cm[33 + cc - 9];
      n[cc] += o;
    }
  }

  // symmetric sweep on all 12 pivots of the SPD block whose row rr is a[]; returns row rr of the NEGATED
  // inverse, which is what sweeping all pivots yields: every consumer (the packed blocks P_k, the hand-over window, the
  // mat-vecs, the couplings) works with N_k = -P_k and flips the sign in an FMA operand for free instead of spending
  // twelve negations per block here.  Lane c publishes its entry of pivot column p, which by symmetry is entry c of
  // pivot row p (one 8-byte store per lane instead of a 96-byte row).  The pivot row is scaled
  // immediately: keeping it unscaled with a deferred factor costs 100x in the residual
  // |P S - I| (measured), although the entrywise error is the same.
  // one pivot step on the lane's row: the pivot lane takes the scaled pivot row from the published column (by symmetry
  // entry c of the pivot row; bitwise what lane c computes as its t -- using the lane's own copy of the row instead costs
  // a factor of 200 in the accuracy of dx, measured, and so does keeping the pivot row unscaled with a deferred factor),
  // every other lane the rank-1 update.  A divergent branch: written as predicated PTX (@q mul / @!q fma) ptxas turns
  // the pair into unconditional instructions plus four FSELs per entry (84 instead of 47 instructions per pivot step).
  // (a[p] is set outside the branch with a select: written as `if (rr == p) a[p] = -di` in all twelve unrolled steps the
  //  compiler recognises a[rr], indexes the row dynamically and moves it to local memory: 2.4x slower, measured)
  SRBD_DEV void sweep_update(R a[12], const R bv[12], R di, R t, bool piv, int p) const {
    if (piv) {
#pragma unroll
      for (int c = 0; c < 12; ++c) if (c != p) a[c] = bv[c] * di;
    } else {
#pragma unroll
      for (int c = 0; c < 12; ++c) if (c != p) a[c] = fmaR(-t, bv[c], a[c]);
    }
  }

  SRBD_DEV void sweep_invert(R a[12]) const {
    R* b = bcr();
    w.sync();   // the broadcast scratch still holds form_row's exchange
#pragma unroll
    for (int p = 0; p < 12; ++p) {
      R* bp = b + 12 * (p & 1);
      bp[rr] = a[p];
      w.sync();
      R bv[12];
      ld12(bp, bv);
      const R di = rcp(bv[p]);
      const bool piv = (rr == p);
      const R t = a[p] * di;
      sweep_update(a, bv, di, t, piv, p);
      a[p] = piv ? -di : t;
    }
  }

  // fwd: the forward substitution of the affine solve (t_next -= M (P_k t_k), solve_thomas) is done here, at the end of
  // each round, while the lane's row of P_k is still in registers (saves the packed re-read of five mat-vecs)
  SRBD_DEV void factor(bool fwd) {
    // rounds 0..3: top k = t (coupling from k-1) | bottom k = 9-t (coupling from k+1)
    // round 4: top k = 4 | bottom: S_55 - C_6' P_6 C_6 (block 5 with its coupling from 6), parked packed in P(5)
    // round 5: top k = 5: the parked rows, coupling from 4, sweep ; bottom idle
    // The own row of the neighbour's inverse is re-read from tensor memory (fp32 mode: from the packed storage), the
    // three other rows a lane needs come from the 72-entry hand-over window.
    double* vy = sm + SM::OFF_VY;
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      const int k = (t == 5) ? 5 : (h ? 9 - t : t);
      const bool act = (t < 4) || (h == 0);
      R n[12];
      if (t < 5) form_row(k, n);
      else {
        if constexpr (kTm) ld12(sm + SM::OFF_P5F + SM::FS * rr, n);
        else {
          const R* P5 = P(5);
#pragma unroll
          for (int c = 0; c < 12; ++c) n[c] = pe(P5, c);
        }
      }
      if (t > 0) {
        R o[12];
        if constexpr (kTm) {
          tm_ld(t - 1, o);    // own row of the block this half eliminated in the previous round (round 5: block 4)
        } else {
          const R* Pn = P((t == 5) ? 4 : (h ? k + 1 : k - 1));
#pragma unroll
          for (int c = 0; c < 12; ++c) o[c] = pe(Pn, c);
        }
        couple(n, o, pr(h), h);
      }
      if (t == 4 && h == 1) {
        if constexpr (kTm) { if (r < 12) st12(sm + SM::OFF_P5F + SM::FS * rr, n); }
        else {
          R* P5 = P(5);
#pragma unroll
          for (int c = 0; c < 12; ++c) if (c <= rr) pe(P5, c) = n[c];
        }
      }
      sweep_invert(n);
      if (fwd && t < 5) {
        double tv[12];
        ld12(vy + 12 * k, tv);
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
        for (int m = 0; m < 12; m += 3) {
          q0 = fma(-(double)n[m], tv[m], q0); q1 = fma(-(double)n[m + 1], tv[m + 1], q1); q2 = fma(-(double)n[m + 2], tv[m + 2], q2);
        }
        double* b = bc();
        // fp64: the sweep's last pivot row sits in the second half of the scratch, slots 0..11 are free; the fp32 sweep's
        // two 12-float buffers overlap them
        if (kF32) w.sync();
        b[rr] = q0 + q1 + q2;
        w.sync();
        const double cv0 = h ? c1[0] : c0[0], cv1 = h ? c1[1] : c0[1], cv2 = h ? c1[2] : c0[2], cv3 = h ? c1[3] : c0[3];
        const int j1 = h ? i1[0] : i0[0], j2 = h ? i1[1] : i0[1], j3 = h ? i1[2] : i0[2];
        const double d = cv0 * b[rr] + cv1 * b[j1] + cv2 * b[j2] + cv3 * b[j3];
        const int kn = (t < 3) ? (h ? k - 1 : k + 1) : (t == 3 ? (h ? 5 : 4) : 5);
        if (act && r < 12) vy[12 * kn + rr] -= d;
      }
      if constexpr (kTm) {
        // warp-collective: every lane stores (the idle bottom half of round 4 writes a row nobody reads)
        if (t < 5) tm_st(t, n);
      }
      if (act) {
        if constexpr (kTm) {      // fp64: only the merge block goes through shared memory (full rows)
          if (t == 5 && r < 12) st12(sm + SM::OFF_P5F + SM::FS * rr, n);
        } else {
          R* Pk = P(k);
#pragma unroll
          for (int c = 0; c < 12; ++c) if (c <= rr) pe(Pk, c) = n[c];
        }
        const int slot = h ? rr : rr - 6;
        if (slot >= 0 && slot < 6) st12(pr(h) + SM::HS * slot, n);
      }
      w.sync();
    }
  }

  // ------------------------------------------------------------------------------------
  // block-Thomas solve on vy[0..119]  (rho -> t -> dy), twisted
  // ------------------------------------------------------------------------------------
  // row rr of a pivot-block inverse (stored negated, packed) times the 12-vector v (fp32 blocks: the stored entries are widened, the
  // products and sums are fp64 -- with fp32 products the refined mode stops at 2e-4 instead of 1e-5)
  SRBD_DEV double prow_dot(const R* Pm, const double v[12]) const {
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
    for (int m = 0; m < 12; m += 3) {
      q0 = fma(-(double)pe(Pm, m), v[m], q0); q1 = fma(-(double)pe(Pm, m + 1), v[m + 1], q1); q2 = fma(-(double)pe(Pm, m + 2), v[m + 2], q2);
    }
    return q0 + q1 + q2;
  }
  SRBD_DEV void solve_thomas(bool fwd_done = false) {
    double* vy = sm + SM::OFF_VY;
    double* b = bc();
    // The blocks are walked with per-lane strides (top half: towards stage 5 from 0, bottom half: from 9) instead of
    // recomputing k, its neighbour and the activity flags from t in every trip; the coupling rows of the direction are
    // selected once per loop, not once per use.
    const int dk = h ? -1 : 1;
    // forward: t_k -= M (P_nb t_nb), nb -> k = nb + dk: top 0->1 ... 3->4, 4->5 ; bottom 9->8 ... 6->5, then idle (its
    // trip 5 recomputes 5->4 without storing).  fwd_done: factor(true) has already done it
    if (!fwd_done) {
      const double cv0 = h ? c1[0] : c0[0], cv1 = h ? c1[1] : c0[1], cv2 = h ? c1[2] : c0[2], cv3 = h ? c1[3] : c0[3];
      const int j1 = h ? i1[0] : i0[0], j2 = h ? i1[1] : i0[1], j3 = h ? i1[2] : i0[2];
      double* vnb = vy + (h ? 108 : 0);
      const R* Pnb = P(h ? 9 : 0);
#pragma unroll    // unrolled (slot numbers and trip flags become immediates): -0.8 % each
      for (int t = 1; t < 6; ++t) {
        const bool act = (t < 5) || (h == 0);
        unsigned tq[24];
        if constexpr (kTm) tm_ld_issue(t - 1, tq);
        double tv[12];
        ld12(vnb, tv);
        if constexpr (kTm) b[rr] = trow_dot(tq, tv);
        else b[rr] = prow_dot(Pnb, tv);
        w.sync();
        const double d = cv0 * b[rr] + cv1 * b[j1] + cv2 * b[j2] + cv3 * b[j3];
        vnb += 12 * dk; Pnb += 78 * dk;
        if (act && r < 12) vnb[rr] -= d;
        w.sync();
      }
    }
    // backward: dy_5 ; then 4|6 ; 3|7 ; 2|8 ; 1|9 ; 0|-   (top uses C_{k+1}' (type 1), bottom uses C_k (type 0))
    {
      const double cv0 = h ? c0[0] : c1[0], cv1 = h ? c0[1] : c1[1], cv2 = h ? c0[2] : c1[2], cv3 = h ? c0[3] : c1[3];
      const int j1 = h ? i0[0] : i1[0], j2 = h ? i0[1] : i1[1], j3 = h ? i0[2] : i1[2];
      double* vk = vy + 60;
      const R* Pk = P(5);
      {
        b[rr] = vk[rr];
        w.sync();
        double bv[12];
        ld12(b, bv);
        double dy;
        if constexpr (kTm) {
          double nr[12];
          ld12(sm + SM::OFF_P5F + SM::FS * rr, nr);
          double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
          for (int m = 0; m < 12; m += 3) { q0 = fma(-nr[m], bv[m], q0); q1 = fma(-nr[m + 1], bv[m + 1], q1); q2 = fma(-nr[m + 2], bv[m + 2], q2); }
          dy = q0 + q1 + q2;
        } else dy = prow_dot(Pk, bv);
        if (h == 0 && r < 12) vk[rr] = dy;
        w.sync();
      }
#pragma unroll    // unrolled (slot numbers and trip flags become immediates): -0.8 % each
      for (int t = 1; t < 6; ++t) {
        const double* yn = vk;                  // the block solved in the previous trip
        // the bottom half has four blocks for the top half's five.  Packed storage: it idles in the last trip.  Tensor
        // memory: it idles in the FIRST trip (6 7 8 9 in trips 2..5 next to 3 2 1 0), so that both halves read the slot
        // 5 - t -- the rows were stored in elimination order, top block t and bottom block 9 - t in slot t
        const bool act = kTm ? ((t >= 2) || (h == 0)) : ((t < 5) || (h == 0));
        unsigned tq[24];
        if constexpr (kTm) tm_ld_issue(5 - t, tq);
        if (act) { vk -= 12 * dk; Pk -= 78 * dk; }
        const double tk = vk[rr] - (cv0 * yn[rr] + cv1 * yn[j1] + cv2 * yn[j2] + cv3 * yn[j3]);
        b[rr] = tk;
        w.sync();
        double bv[12];
        ld12(b, bv);
        double dy;
        if constexpr (kTm) dy = trow_dot(tq, bv);
        else dy = prow_dot(Pk, bv);
        if (act && r < 12) vk[rr] = dy;
        w.sync();
      }
    }
  }

  SRBD_DEV double warp_sum(double v) const {
    for (int m = 16; m >= 1; m >>= 1) v += w.shfl_xor(v, m);
    return v;
  }
  SRBD_DEV double warp_min(double v) const {
    for (int m = 16; m >= 1; m >>= 1) { const double o = w.shfl_xor(v, m); v = o < v ? o : v; }
    return v;
  }
  // stage-shifted warm start (warm_xy == 2): floor of the re-initialised slacks, floor of the carried multipliers and
  // the complementarity product rows are re-centred to where the contact schedule changed their bound -- tuned on the
  // closed-loop rollouts of tests/closed_loop.py with the oracle
  static constexpr double WARM_S_MIN = 1e-2, WARM_Z_MIN = 1e-3, WARM_MU0 = 0.1;
  SRBD_DEV static double ratio(double v, double dv) { return dv < 0.0 ? -v * rcp(dv) : 1.0; }
  SRBD_DEV static double step_len(double mn) { return fmax(fmin(1.0, 0.99 * mn), 1e-12); }

  // ------------------------------------------------------------------------------------
  // the whole call
  // ------------------------------------------------------------------------------------
  SRBD_DEV int run(const PdipmArgs& A, long long env) {
    pa = &A; e = env;
    lane = w.lane();
#if defined(__CUDACC__) && !defined(SRBD_HOST_EMUL)
    asm volatile("" : "+r"(lane));   // keep it in a register: the compiler otherwise re-reads the special register all over
#endif
    h = lane >> 4; r = lane & 15; rr = r < 12 ? r : 11;
    {
      double* oX = gout(IPO_X, NZ); double* oS = gout(IPO_S, NINEQ); double* oZ = gout(IPO_Z, NINEQ); double* oY = gout(IPO_Y, NEQ);
      const bool warm = A.cold_start && A.warm_xy && !(A.reset && A.reset[e] != 0);
      const bool keep_xy = warm;
      // warm_xy == 2 (fused tick only): (x, y, z) of the previous tick are kept -- shifted by one stage when the contact
      // schedule has rolled on -- and s, z are re-centred where the schedule changed (below).  The previous tick's
      // F_max * contact row is still in the work buffer: read it before the set-up overwrites it.
      const bool shifted = warm && A.warm_xy == 2 && A.fused;
      double fc_old = 0.0;
      if (shifted && lane < 20) fc_old = gws()[FW_FC + lane];
      if (A.cold_start) {
        if (!keep_xy) {
          for (int i = lane; i < NZ; i += 32) oX[i] = 0.0;
          for (int i = lane; i < NEQ; i += 32) oY[i] = 1.0;
        }
        if (!shifted) for (int i = lane; i < NINEQ; i += 32) oZ[i] = 1.0;
      } else {
        const double* ix = gin(IP_X, NZ); const double* is = gin(IP_S, NINEQ);
        const double* iz = gin(IP_Z, NINEQ); const double* iy = gin(IP_Y, NEQ);
        for (int i = lane; i < NZ; i += 32) oX[i] = ix[i];
        for (int i = lane; i < NINEQ; i += 32) { oS[i] = is[i]; oZ[i] = iz[i]; }
        for (int i = lane; i < NEQ; i += 32) oY[i] = iy[i];
      }
      if (setup()) {   // not qp_former-shaped data: fail loudly (NaN iterates, iteration count -1)
        const double nan = __builtin_nan("");
        for (int i = lane; i < NZ; i += 32) oX[i] = nan;
        for (int i = lane; i < NINEQ; i += 32) { oS[i] = nan; oZ[i] = nan; }
        for (int i = lane; i < NEQ; i += 32) oY[i] = nan;
        if (lane == 0) {
          double* oRes = gout(IPO_RES, 4);
          oRes[0] = nan; oRes[1] = nan; oRes[2] = nan; oRes[3] = nan; gout(IPO_MU, 1)[0] = nan;
          if (A.iters_done) A.iters_done[e] = -1;
        }
        return -1;
      }
      if (shifted) {
        // ---- one-stage shift when the contact schedule rolled (stage k now is what stage k + 1 was) ---------------
        const double fc_new = lane < 20 ? gws()[FW_FC + lane] : 0.0;
        const bool rolled = w.any(lane < 20 && fc_new != fc_old);
        if (rolled) {
          double xv[8], yv[5], zv[5];
#pragma unroll
          for (int m = 0; m < 8; ++m) {   // x: states 0..119 and inputs 120..239, twelve per stage
            const int i = lane + 32 * m;
            const int src = (i < NZ && (i % 120) < 108) ? i + 12 : i;
            xv[m] = i < NZ ? oX[src] : 0.0;
          }
#pragma unroll
          for (int m = 0; m < 5; ++m) {   // y: dynamics rows 0..119 (twelve per stage), Mx rows 120..139 (two per stage); z: sixteen
            const int i = lane + 32 * m;
            const int ys = i < 108 ? i + 12 : ((i >= 120 && i < 138) ? i + 2 : i);
            yv[m] = i < NEQ ? oY[ys] : 0.0;
            const int zs = i < 144 ? i + 16 : i;
            zv[m] = oZ[zs];
          }
          w.sync();
#pragma unroll
          for (int m = 0; m < 8; ++m) { const int i = lane + 32 * m; if (i < NZ) oX[i] = xv[m]; }
#pragma unroll
          for (int m = 0; m < 5; ++m) { const int i = lane + 32 * m; if (i < NEQ) oY[i] = yv[m]; oZ[i] = zv[m]; }
          const double nxt = w.shfl(fc_old, lane + 2 < 20 ? lane + 2 : lane);
          fc_old = nxt;
          w.sync();
        }
        // ---- s = max(h - G x, s_min); z kept (floored) where the row's bound is unchanged, re-centred (z = mu0 / s)
        //      where the schedule changed it (rows 7 / 15 of a stage: fz <= F_max * contact) -----------------------
        const double* f = fb();
        const double ga = gva, gb = gvb;
        const int q7 = (r & 7) == 7;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t, j = 16 * k + r;
          const double gx = ga * oX[120 + 12 * k + ca] + gb * oX[120 + 12 * k + cz];
          const double sv = fmax(h_of(k, j) - gx, WARM_S_MIN);
          const double fo = w.shfl(fc_old, 2 * k + ((r >> 3) & 1));
          const bool changed = q7 && fo != gws()[FW_FC + 2 * k + ((r >> 3) & 1)];
          oS[j] = sv;
          oZ[j] = changed ? WARM_MU0 * rcp(sv) : fmax(oZ[j], WARM_Z_MIN);
        }
      } else if (A.cold_start) {   // s = max(h - G x, 1) (x = 0 unless warm); in fused mode h only exists after the set-up
        const double* f = fb();
        const double ga = gva, gb = gvb;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t, j = 16 * k + r;
          const double gx = keep_xy ? ga * oX[120 + 12 * k + ca] + gb * oX[120 + 12 * k + cz] : 0.0;
          oS[j] = fmax(h_of(k, j) - gx, 1.0);
        }
      }
      w.sync();
      {
        double* ws = gws();
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const int k = h ? 9 - t : t, j = 16 * k + r;
          double* row = ws + j;
          row[FW_SV] = oS[j]; row[FW_ZV] = oZ[j]; row[FW_H] = h_of(k, j);
        }
      }
    }

    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    int it = 0;
    double n_rx = 0, n_rs = 0, n_re = 0, mu_new = 0;
#pragma unroll 1
    for (; it < A.num_iters; ++it) {
      double mu;
      phase_a(mu, n_rx, n_rs, n_re);

      // ============ two solves with one factorisation: pass 0 affine, pass 1 corrector ======
      double smu = 0.0, mn_s = 1e300, mn_z = 1e300;
#ifdef SRBD_HOST_EMUL
      double a_pri_aff = 0, a_dual_aff = 0;
#endif
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        {
          // the lane's row of Bbar stays in registers for the five stages: a lane-varying row load costs
          // twice the wavefronts of a broadcast (profiles/smem_bcast_r1_wavefronts.csv)
          double Brow[12];
          ld12(fb() + FB_B + 14 * rr, Brow);
#pragma unroll 1
          for (int t = 0; t < 5; ++t) {
            const int k = h ? 9 - t : t;
            if (pass == 1) corrector_rhs(k, smu);
            build_rho(k, pass == 1, Brow);
          }
        }
        if (pass == 0) {
          SRBD_TRACE("rho_aff", vy, 140);
          factor(true);
          SRBD_TRACE("P", sm + SM::OFF_P, 780);
        }
        solve_thomas(pass == 0);
        if (pass == 0) SRBD_TRACE("dyd_aff", vy, 120);
        mn_s = 1e300; mn_z = 1e300;   // min over ALL 160 candidates (1.0 where dv >= 0): may exceed 1
        {
          double bc8[8];
          ld_bcol(bc8);
#pragma unroll 1
          for (int t = 0; t < 5; ++t) recover(h ? 9 - t : t, pass, bc8, mn_s, mn_z);
        }
        if (kF32) for (int q = 2; q <= A.precision; ++q) refine(pass, mn_s, mn_z);
        if (pass == 0) {
          w.sync();
          double* ws = gws();
          for (int i = lane; i < NZ; i += 32) ws[FW_DX + i] = vx[i];
          for (int i = lane; i < NEQ; i += 32) ws[FW_DY + i] = vy[i];
          SRBD_TRACE("dx_aff", vx, 240);
          SRBD_TRACE("dy_aff", vy, 140);
          const double ap = step_len(warp_min(mn_s));
          const double ad = step_len(warp_min(mn_z));
#ifdef SRBD_HOST_EMUL
          a_pri_aff = ap; a_dual_aff = ad;
#endif
          double a_mu = 0.0;
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const double* row = ws + 16 * (h ? 9 - t : t) + r;
            a_mu = fma(fma(ap, row[FW_S], row[FW_SV]), fma(ad, row[FW_Z], row[FW_ZV]), a_mu);
          }
          const double mu_aff = warp_sum(a_mu) / (double)NINEQ;
          const double sg = mu_aff * rcp(mu);
          smu = sg * sg * sg * mu;
          w.sync();
        }
      }
      const double a_pri = step_len(warp_min(mn_s));
      const double a_dual = step_len(warp_min(mn_z));
      w.sync();
#ifdef SRBD_HOST_EMUL
      { double dbg[8] = {a_pri_aff, a_dual_aff, smu, mu, a_pri, a_dual, 0.0, 0.0}; SRBD_TRACE("scalars", dbg, 8); }
      SRBD_TRACE("dx_cor", vx, 240);
      SRBD_TRACE("dy_cor", vy, 140);
#endif
      // ======================= update ====================================================
      {
        double* ws = gws();
        double* oX = gout(IPO_X, NZ); double* oY = gout(IPO_Y, NEQ);
        double a_new = 0.0;
        {   // all global loads of a lane are issued before the first dependent use (L2 latency once)
          double xa[8], xo[8], ya[5], yo[5];
#pragma unroll
          for (int m = 0; m < 8; ++m) { const int i = lane + 32 * m; const bool on = i < NZ; xa[m] = on ? ws[FW_DX + i] : 0.0; xo[m] = on ? oX[i] : 0.0; }
#pragma unroll
          for (int m = 0; m < 5; ++m) { const int i = lane + 32 * m; const bool on = i < NEQ; ya[m] = on ? ws[FW_DY + i] : 0.0; yo[m] = on ? oY[i] : 0.0; }
#pragma unroll
          for (int m = 0; m < 8; ++m) { const int i = lane + 32 * m; if (i < NZ) oX[i] = fma(a_pri, xa[m] + vx[i], xo[m]); }
#pragma unroll
          for (int m = 0; m < 5; ++m) { const int i = lane + 32 * m; if (i < NEQ) oY[i] = fma(a_dual, ya[m] + vy[i], yo[m]); }
        }
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          double* row = ws + 16 * (h ? 9 - t : t) + r;
          const double sn = fmax(fma(a_pri, row[FW_S], row[FW_SV]), 1e-8);
          const double zn = fmax(fma(a_dual, row[FW_Z], row[FW_ZV]), 1e-8);
          row[FW_SV] = sn; row[FW_ZV] = zn;
          a_new = fma(sn, zn, a_new);
        }
        mu_new = warp_sum(a_new) / (double)NINEQ;
      }
      w.sync();
      if (A.tol > 0.0 && n_rx < A.tol && n_rs < A.tol && n_re < A.tol && mu_new < A.tol) { ++it; break; }
    }
    {   // s, z of the lane's rows back to the caller's buffers
      const double* ws = gws();
      double* oS = gout(IPO_S, NINEQ); double* oZ = gout(IPO_Z, NINEQ);
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const int j = 16 * (h ? 9 - t : t) + r;
        oS[j] = ws[FW_SV + j]; oZ[j] = ws[FW_ZV + j];
      }
    }
    if (lane == 0) {
      double* oRes = gout(IPO_RES, 4);
      oRes[0] = n_rx; oRes[1] = n_rs; oRes[2] = n_re; oRes[3] = mu_new;
      gout(IPO_MU, 1)[0] = mu_new;
      if (A.iters_done) A.iters_done[e] = it;
    }
    if (A.fused && A.wrench && lane < 12)   // u0 -> body-frame wrench (the x updates above are visible: w.sync())
      A.wrench[e * 12 + lane] = wrench_entry(gout(IPO_X, NZ) + NX * N, A.rot + e * 9, lane);
    return it;
  }

  // ======================= phase A: residuals, weights, Psi, r1~ ===========================
  SRBD_DEV void phase_a(double& mu, double& n_rx, double& n_rs, double& n_re) const {
    const double* f = fb();
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* sx = sm + SM::OFF_P;         // iterate scratch (phase A only): x | y | (hx + beta) x
    double* sy = sx + 240; double* xs = sx + 380;
    double* ws = gws();
    const bool fused = pa->fused != 0;
    const double* gF = gin(IP_F, NZ); const double* gB = gin(IP_B, NEQ);
    {
      const double* oX = gout(IPO_X, NZ); const double* oY = gout(IPO_Y, NEQ);
      double xv[8], yv[5];
#pragma unroll
      for (int m = 0; m < 8; ++m) { const int i = lane + 32 * m; xv[m] = i < NZ ? oX[i] : 0.0; }
#pragma unroll
      for (int m = 0; m < 5; ++m) { const int i = lane + 32 * m; yv[m] = i < NEQ ? oY[i] : 0.0; }
      int m12 = lane % 12;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int i = lane + 32 * m;
        if (i < NZ) sx[i] = xv[m];
        if (i < 120) xs[i] = xv[m] * (f[FB_HX + m12] + IP_BETA);
        m12 += 8; if (m12 >= 12) m12 -= 12;
      }
#pragma unroll
      for (int m = 0; m < 5; ++m) { const int i = lane + 32 * m; if (i < NEQ) sy[i] = yv[m]; }
    }
    w.sync();
    const int jj = r & 7, ft = r >> 3;
    const double valid = r < 12 ? 1.0 : 0.0;
    double a_sz = 0, a_rs = 0, a_rx = 0, a_re = 0;
    double Brow[12];
    ld12(f + FB_B + 14 * rr, Brow);
#pragma unroll 1
    for (int t = 0; t < 5; ++t) {
      const int k = h ? 9 - t : t;
      const int j = 16 * k + r;
      double* row = ws + j;
      const double sv = row[FW_SV], zv = row[FW_ZV], hv = row[FW_H];
      double fu = 0.0, fxv, bd, bm = 0.0;           // fused: f_u = 0, b_mx = 0 by construction
      if (fused) { fxv = ws[FW_FX + 12 * k + rr]; bd = ws[FW_CD + (k == 0 ? 12 : 0) + rr]; }
      else {
        fu = ldg(gF + 120 + 12 * k + rr); fxv = ldg(gF + 12 * k + rr); bd = ldg(gB + 12 * k + rr);
        bm = ldg(gB + 120 + 2 * k + (r & 1));
      }
      const double* u = sx + 120 + 12 * k;
      double* ps = psi(k);
      double gdg, gz;
      {   // inequality row j
        const double ga = gva, gb = gvb;
        const double pga = w.shfl_xor(ga, 1), pgb = w.shfl_xor(gb, 1);
        const double rs = ga * u[ca] + gb * u[cz] + sv - hv;
        const double si = rcp(sv);
        const double W = fma(si, zv, IP_DELTA);
        const double q = rcp(fma(IP_DELTA, W, 1.0));     // D / W
        const double D = W * q;
        const double Dg = fma(D, rs, -(q * zv));         // D (r2 / W - r3), r2 = -z, r3 = -rs
        row[FW_S] = fma(IP_DELTA, Dg, -rs); row[FW_Z] = Dg; row[FW_Q] = q;
        row[FW_D] = D;
        gdg = gather_rows(Dg, ga, gb); gz = gather_rows(zv, ga, gb);
        a_sz = fma(sv, zv, a_sz); a_rs = fma(rs, rs, a_rs);
        // Psi: pairs (0,1)->fx (2,3)->fy (4,5)->my ; (6,7) -> fz only
        const double pD = w.shfl_xor(D, 1);
        const double gg = gb * gb * D + pgb * pgb * pD;
        double contrib = gg, ia = 0.0, tq = 0.0;
        if (jj < 6) {
          const double eps_ca = f[FB_HU + ca] + IP_BETA;
          const double cr = gb * pga - pgb * ga;
          const double aa = eps_ca + ga * ga * D + pga * pga * pD;
          const double bb = ga * gb * D + pga * pgb * pD;
          ia = rcp(aa);
          tq = bb * ia;
          contrib = (eps_ca * gg + D * pD * cr * cr) * ia;
        }
        if (jj & 1) contrib = 0.0;
        contrib += w.shfl_xor(contrib, 1);
        contrib += w.shfl_xor(contrib, 2);
        contrib += w.shfl_xor(contrib, 4);
        if (jj < 6 && !(jj & 1)) { ps[7 * ft + (jj >> 1)] = ia; ps[7 * ft + 3 + (jj >> 1)] = tq; }
        if (jj == 6) ps[7 * ft + 6] = rcp(f[FB_HU + cz] + IP_BETA + contrib);
      }
      {   // input column rr : rx_u and r1~_u
        const double* yk = sy + 12 * k;
        double acc = fma(f[FB_HU + rr], u[rr], fu) + bcol_dot(yk);
        if (psk == 3) acc += sy[120 + 2 * k + psi_i];
        acc += gz;
        a_rx = fma(valid * acc, acc, a_rx);
        vx[120 + 12 * k + rr] = -acc - gdg;
        // state column rr
        const double hx_rr = f[FB_HX + rr];
        double ax = fma(hx_rr, sx[12 * k + rr], fxv) + yk[rr];
        if (k <= 8) {
          const double* yn = sy + 12 * (k + 1);
          ax = fma(hx_rr + IP_BETA, c1[0] * yn[rr] + c1[1] * yn[i1[0]] + c1[2] * yn[i1[1]] + c1[3] * yn[i1[2]], ax);
        }
        a_rx = fma(valid * ax, ax, a_rx);
        vx[12 * k + rr] = -ax;
        // dynamics row rr
        double uv[12];
        ld12(u, uv);
        double re = sx[12 * k + rr] - bd + dot12(Brow, uv);
        if (k >= 1) {
          const double* xp = xs + 12 * (k - 1);
          re += c0[0] * xp[rr] + c0[1] * xp[i0[0]] + c0[2] * xp[i0[1]] + c0[3] * xp[i0[2]];
        }
        a_re = fma(valid * re, re, a_re);
        vy[12 * k + rr] = re;
        if (kF32 && pa->precision >= 2) ws[FW_RE + 12 * k + rr] = re;
        if (r >= 12 && r < 14) {   // Mx rows
          const double rm = u[6 + 3 * (r - 12)] - bm;
          a_re = fma(rm, rm, a_re);
          vy[120 + 2 * k + (r - 12)] = rm;
        }
      }
      w.sync();
    }
    a_sz = warp_sum(a_sz); a_rs = warp_sum(a_rs); a_rx = warp_sum(a_rx); a_re = warp_sum(a_re);
    mu = a_sz / (double)NINEQ;
    n_rx = sqrt(a_rx); n_rs = sqrt(a_rs); n_re = sqrt(a_re);
  }

  // corrector right-hand side of stage k (sparse_pdipm_solver.py:478-479): r1 = r3 = r4 = 0,
  // r2 = -(s z + ds_aff dz_aff - sigma mu) / s ; leaves -G' D g in the input part of vx
  SRBD_DEV void corrector_rhs(int k, double smu) const {
    double* vx = sm + SM::OFF_VX;
    double* row = gws() + 16 * k + r;
    const double sv = row[FW_SV], zv = row[FW_ZV], dsa = row[FW_S], dza = row[FW_Z], q = row[FW_Q];
    const double r2 = -(rcp(sv) * (sv * zv + dsa * dza - smu));
    const double Dg = q * r2;
    row[FW_S] = fma(IP_DELTA, Dg, dsa); row[FW_Z] = dza + Dg;
    vx[120 + 12 * k + rr] = -gather_rows(Dg, gva, gvb);
    w.sync();
  }

  // fp32 pivot blocks only: one step of iterative refinement of the solve just recovered.  With dx = Phi^-1 (r1~ - A' dy)
  // the residual of the Schur system is e = rho - S dy = A dx - delta dy - r4, evaluated here in fp64 from the recovered
  // step (the Mx rows and their inputs were eliminated in closed form and carry no error); S ddy = e is solved with the
  // same fp32 blocks and the correction is recovered like a corrector solve with a zero right-hand side (r1~ = 0,
  // r4 = 0), whose ds / dz parts accumulate onto the lane-private slots.  Measured: DESIGN.md section 8.
  SRBD_DEV void refine(int pass, double& mn_s, double& mn_z) {
    const double* f = fb();
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    double* keep = sm + SM::OFF_PR;   // the hand-over window is only live inside factor(): 144 >= 140 doubles
    double* ws = gws();
    w.sync();
    for (int i = lane; i < NEQ; i += 32) keep[i] = vy[i];
    for (int i = lane; i < NZ; i += 32) ws[FW_RX + i] = vx[i];
    double ev[5];
    {
      double Brow[12];
      ld12(f + FB_B + 14 * rr, Brow);
      const double h0 = f[FB_HX + rr] + IP_BETA, h1 = f[FB_HX + i0[0]] + IP_BETA, h2 = f[FB_HX + i0[1]] + IP_BETA, h3 = f[FB_HX + i0[2]] + IP_BETA;
#pragma unroll 1
      for (int t = 0; t < 5; ++t) {
        const int k = h ? 9 - t : t;
        double uv[12];
        ld12(vx + 120 + 12 * k, uv);
        double e = vx[12 * k + rr] + dot12(Brow, uv) - IP_DELTA * vy[12 * k + rr];
        if (k >= 1) {
          const double* xp = vx + 12 * (k - 1);
          e += c0[0] * h0 * xp[rr] + c0[1] * h1 * xp[i0[0]] + c0[2] * h2 * xp[i0[1]] + c0[3] * h3 * xp[i0[2]];
        }
        if (pass == 0) e += ws[FW_RE + 12 * k + rr];   // - r4 = + re
        ev[t] = e;
      }
    }
    w.sync();
#pragma unroll
    for (int t = 0; t < 5; ++t) if (r < 12) vy[12 * (h ? 9 - t : t) + rr] = ev[t];
    for (int i = lane; i < 120; i += 32) vx[120 + i] = 0.0;
    w.sync();
    solve_thomas();
    mn_s = 1e300; mn_z = 1e300;
    {
      double bc8[8];
      ld_bcol(bc8);
#pragma unroll 1
      for (int t = 0; t < 5; ++t) recover(h ? 9 - t : t, 1, bc8, mn_s, mn_z);
    }
    w.sync();
    for (int i = lane; i < NZ; i += 32) vx[i] += ws[FW_RX + i];
    for (int i = lane; i < NEQ; i += 32) vy[i] += keep[i];
    w.sync();
  }

  // recover dx, du (into vx), dy_mx (into vy) and this lane's ds / dz of inequality row 16k + r
  SRBD_DEV void recover(int k, int pass, const double bc8[8], double& mn_s, double& mn_z) const {
    const double* f = fb();
    double* vx = sm + SM::OFF_VX;
    double* vy = sm + SM::OFF_VY;
    const int j = 16 * k + r;
    double* row = gws() + j;
    const double sv = row[FW_SV], zv = row[FW_ZV], sS = row[FW_S], sZ = row[FW_Z], Dj = row[FW_D];
    double* b = bc();
    double* vu = vx + 120 + 12 * k;
    const double* dyk = vy + 12 * k;
    b[rr] = vu[rr] - bcol_dot(bc8, dyk);
    w.sync();
    {
      double du = psi_apply(psi(k), b);
      if (psk == 3) {     // Mx: input and its equality multiplier in closed form
        const double r4 = pass ? 0.0 : -vy[120 + 2 * k + psi_i];
        const double eps = f[FB_DER + 6 + psi_i], iden = f[moff];
        du = fma(IP_DELTA, b[rr], r4) * iden;
        if (r < 12) vy[120 + 2 * k + psi_i] = (b[rr] - eps * r4) * iden;
      }
      double wx = (pass ? 0.0 : vx[12 * k + rr]) - dyk[rr];
      if (k <= 8) {
        const double* dyn = vy + 12 * (k + 1);
        wx = fma(-(f[FB_HX + rr] + IP_BETA), c1[0] * dyn[rr] + c1[1] * dyn[i1[0]] + c1[2] * dyn[i1[1]] + c1[3] * dyn[i1[2]], wx);
      }
      if (r < 12) { vx[12 * k + rr] = phx_rr * wx; vu[rr] = du; }
    }
    w.sync();
    {
      const double gdu = gva * vu[ca] + gvb * vu[cz];
      const double Dg = Dj * gdu;
      const double dz = sZ + Dg;
      const double ds = fma(IP_DELTA, Dg, sS - gdu);
      row[FW_S] = ds; row[FW_Z] = dz;       // affine (pass 0) / combined (pass 1)
      mn_s = fmin(mn_s, ratio(sv, ds));
      mn_z = fmin(mn_z, ratio(zv, dz));
    }
  }
};
using PdipmFast = PdipmFastT<false>;

}  // namespace srbd
