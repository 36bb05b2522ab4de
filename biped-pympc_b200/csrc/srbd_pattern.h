// Fixed problem dimensions and CCS sparsity patterns of the SRBD-MPC QP.
//
// The patterns are the structural sparsity of the reference's CasADi `qp_former`
// outputs (biped_pympc/casadi/srbd_constraints.py:139,225; captured at codegen time by
// biped_pympc/casadi/generate_solver_function.py:69-76) and are therefore the *type* of
// the matrix arguments of `sparse_pdipm_multiple_iterations`
// (biped_pympc/casadi/sparse_pdipm_solver.py:357-534).  Everything here is integer
// index arithmetic shared by the host (table construction) and the kernels.
//
// Variable order  z = [x_1 .. x_10, u_0 .. u_9]         (srbd_constraints.py:26,105)
//   x_{k+1}[c] -> 12k + c          u_k[c] -> 120 + 12k + c
// Equality rows   dyn_k[r] -> 12k + r  ;  Mx_left_k -> 120 + 2k ; Mx_right_k -> 121 + 2k
// Inequality rows 16k + 8*foot + j    (srbd_constraints.py:186-222)
#pragma once

#if defined(__CUDACC__)
#define SRBD_HD __host__ __device__ __forceinline__
#else
#define SRBD_HD inline
#endif

namespace srbd {

constexpr int N = 10;        // horizon
constexpr int NX = 12;       // states  [roll,pitch,yaw, p(3), omega(3), v(3)]
constexpr int NU = 12;       // inputs  [f_L(3), f_R(3), m_L(3), m_R(3)]
constexpr int NZ = 240;
constexpr int NEQ = 140;
constexpr int NINEQ = 160;
constexpr int NNZ_H = 240;
constexpr int NNZ_A = 1196;
constexpr int NNZ_G = 280;

// ---- qp_former argument list (srbd_constraints.py:77-78) -------------------------------
constexpr int QP_N_IN = 17;
constexpr int QP_N_OUT = 6;
enum QpIn { IN_X0 = 0, IN_X, IN_U, IN_XREF, IN_DT, IN_M, IN_MU, IN_RBODY, IN_IWORLD, IN_BODY,
            IN_LFOOT, IN_RFOOT, IN_CONTACT, IN_Q, IN_R, IN_RESLIN, IN_RESANG };
enum QpOut { OUT_H = 0, OUT_F, OUT_A, OUT_B, OUT_G, OUT_D };
// ---- sparse_pdipm_multiple_iterations argument list (sparse_pdipm_solver.py:532-534) ---
constexpr int IP_N_IN = 10;
constexpr int IP_N_OUT = 6;
enum IpIn { IP_Q = 0, IP_G, IP_A, IP_F, IP_H, IP_B, IP_X, IP_S, IP_Z, IP_Y };
enum IpOut { IPO_X = 0, IPO_S, IPO_Z, IPO_Y, IPO_RES, IPO_MU };

// ---- A (140 x 240, 1196 nnz), column-major, rows ascending inside a column -------------
// x-column (k, c) = column 12k+c, k = 0..9:
//   [0]            E_k[c]      (row 12k+c, coefficient of x_{k+1}[c] in dyn_k; +1)
//   k <= 8 only, rows of block k+1 (coefficients of x_{k+1} in dyn_{k+1}; -A_d[:,c]):
//     c in 0..5 : [1] row c
//     c in 6..8 : [1..3] rows 0,1,2   [4] row c
//     c in 9..11: [1] row c-6         [2] row c
SRBD_HD int a_xcol_off(int c) { return c < 6 ? 2 * c : (c < 9 ? 12 + 5 * (c - 6) : 27 + 3 * (c - 9)); }
SRBD_HD int a_xcol_start(int k, int c) { return k < 9 ? 36 * k + a_xcol_off(c) : 324 + c; }
// position inside x-column c (k<=8) of block-(k+1) row r, or -1 when structurally zero
SRBD_HD int a_xcol_pos(int r, int c) {
  if (r == c) return c < 6 ? 1 : (c < 9 ? 4 : 2);
  if (c >= 6 && c < 9 && r < 3) return 1 + r;
  if (c >= 9 && r == c - 6) return 1;
  return -1;
}
// u-column (k, c) = column 120+12k+c:  rows of block k (-B_d[:,c]) then the Mx row
//   c in 0..5 : rows 0,1,2, 3+c%3, 6,7,8, 9+c%3                   (8)
//   c in 6..11: rows 0,1,2, 6,7,8  (+ Mx row for c == 6 / c == 9)  (6 or 7)
SRBD_HD int a_ucol_off(int c) {
  return c < 6 ? 8 * c : (c == 6 ? 48 : c == 7 ? 55 : c == 8 ? 61 : c == 9 ? 67 : c == 10 ? 74 : 80);
}
SRBD_HD int a_ucol_start(int k, int c) { return 336 + 86 * k + a_ucol_off(c); }
SRBD_HD int a_ucol_pos(int r, int c) {
  if (r < 3) return r;
  if (r >= 6 && r < 9) return c < 6 ? r - 2 : r - 3;
  if (c >= 6) return -1;
  if (r == 3 + c % 3) return 3;
  if (r == 9 + c % 3) return 7;
  return -1;
}
constexpr int A_UCOL_MX_POS = 6;   // position of the Mx row entry in u-columns 6 and 9

// ---- G (160 x 240, 280 nnz): 28 per stage, u-columns only ------------------------------
// stage base 28k; foot ft in {0,1}: gb = 28k + 12 ft ; my column base mb = 28k + 24 + 2 ft
//   row j of the foot:  j=0,1: (fx @ gb+j , fz @ gb+4+j)      j=2,3: (fy @ gb+j , fz @ gb+4+j)
//                       j=4,5: (fz @ gb+4+j, my @ mb+j-4)     j=6,7: (fz @ gb+4+j)
SRBD_HD int g_foot_base(int k, int ft) { return 28 * k + 12 * ft; }
SRBD_HD int g_my_base(int k, int ft) { return 28 * k + 24 + 2 * ft; }
// input index (0..11) of a foot's fx, fy, fz, my
SRBD_HD int u_fx(int ft) { return 3 * ft; }
SRBD_HD int u_fy(int ft) { return 3 * ft + 1; }
SRBD_HD int u_fz(int ft) { return 3 * ft + 2; }
SRBD_HD int u_my(int ft) { return 7 + 3 * ft; }

// constants baked into the reference functions
constexpr double F_MAX = 500.0;   // srbd_constraints.py:31
constexpr double LT = 0.07;       // srbd_constraints.py:161
constexpr double LH = 0.04;       // srbd_constraints.py:162
constexpr double GRAV_Z = -9.81;  // srbd_centroidal_model.py:136
constexpr double IP_BETA = 1e-8;  // generate_solver_function.py:112
constexpr double IP_DELTA = 1e-8; // sparse_pdipm_solver.py:416

}  // namespace srbd
