/* CPU oracle (plain C) for the batched SRBD-MPC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call this file.  The product (biped-pympc_b200/) never does.
 *
 * PARITY PINNED TO THE REFERENCE'S SOURCE, NOT TO CASADI: CasADi is not available offline and the
 * reference ships no golden vectors for this path.  The committed fixture
 * tests/golden/srbd_ref_shim_v1.npz is what the reference's own unmodified Python files compute when
 * executed over a numeric stand-in for the casadi module (tests/golden/make_golden_ref.py; full
 * statement in oracle/srbd_oracle.py); tests/test_reference_pin.py checks this file against it
 * (assembly <= 1e-12, iterates <= 1e-9 after the reference's 5 iterations), tests/test_oracle_c.py
 * against the numpy restatement.  This is a C restatement of the same two reference functions:
 *
 *   srbd_oracle_qp_former  <- biped_pympc/casadi/srbd_constraints.py:20-227 with the dynamics
 *                             of biped_pympc/casadi/srbd_centroidal_model.py:101-166 (literal
 *                             RK4 of the affine map; b = A z - eq(z), d = G z - ineq(z))
 *   srbd_oracle_pdipm      <- biped_pympc/casadi/sparse_pdipm_solver.py:357-534: per Newton
 *                             step the full 700x700 regularised KKT in variable order
 *                             [dx, ds, dz, dy], a sparse LDL^T WITHOUT numerical pivoting
 *                             (ca.ldl semantics, :451), two solves with the same factors,
 *                             ratio tests, clamps.  CasADi orders the KKT with AMD; its exact
 *                             permutation cannot be reproduced here, so a fixed fill-reducing
 *                             stage-interleaved ordering is used instead (any symmetric
 *                             permutation is admissible: the KKT is quasi-definite).
 *
 * REAL = double by default; compile with -DSRBD_REAL_LONG_DOUBLE for the 80-bit tie-breaker
 * build (inputs/outputs stay double).  OpenMP `parallel for` over environments, the
 * reference's own CPU convention (cusadi/src/benchmark_functions/evaluate_parallel_cpu.cpp:79-82).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef SRBD_REAL_LONG_DOUBLE
typedef long double REAL;
#define RSQRT sqrtl
#define RFMAX fmaxl
#define RFMIN fminl
#define FN(name) name##_ld
#else
typedef double REAL;
#define RSQRT sqrt
#define RFMAX fmax
#define RFMIN fmin
#define FN(name) name
#endif

enum { NH = 10, NX = 12, NU = 12, NZ = 240, NEQ = 140, NINEQ = 160, NKKT = 700,
       NNZ_A = 1196, NNZ_G = 280, NNZ_H = 240 };
static const double F_MAX = 500.0, LT = 0.07, LH = 0.04, GRAV = -9.81, BETA = 1e-8, DELTA = 1e-8;

/* ------------------------------------------------------------------------------------------
 * structural patterns, derived from the symbolic dependency structure (not from numbers)
 * ---------------------------------------------------------------------------------------- */
static int g_ready = 0;
static int a_colptr[NZ + 1], a_row[NNZ_A], g_colptr[NZ + 1], g_row[NNZ_G];

static void build_patterns(void) {
  static unsigned char Amask[NEQ][NZ], Gmask[NINEQ][NZ];
  unsigned char Ad[NX][NX], Bd[NX][NU];
  memset(Amask, 0, sizeof Amask); memset(Gmask, 0, sizeof Gmask);
  memset(Ad, 0, sizeof Ad); memset(Bd, 0, sizeof Bd);
  for (int i = 0; i < NX; ++i) Ad[i][i] = 1;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Ad[i][6 + j] = 1; Ad[3 + i][9 + i] = 1; }
  for (int c = 0; c < NU; ++c) { for (int i = 0; i < 3; ++i) { Bd[6 + i][c] = 1; Bd[i][c] = 1; } }   /* omega rows, euler rows */
  for (int i = 0; i < 3; ++i) { Bd[9 + i][i] = Bd[9 + i][3 + i] = 1; Bd[3 + i][i] = Bd[3 + i][3 + i] = 1; }
  for (int k = 0; k < NH; ++k) {
    for (int r = 0; r < NX; ++r) {
      Amask[12 * k + r][12 * k + r] = 1;
      for (int c = 0; c < NX; ++c) if (k >= 1 && Ad[r][c]) Amask[12 * k + r][12 * (k - 1) + c] = 1;
      for (int c = 0; c < NU; ++c) if (Bd[r][c]) Amask[12 * k + r][120 + 12 * k + c] = 1;
    }
    Amask[120 + 2 * k][120 + 12 * k + 6] = 1;
    Amask[121 + 2 * k][120 + 12 * k + 9] = 1;
    for (int ft = 0; ft < 2; ++ft) {
      const int r0 = 16 * k + 8 * ft, u0 = 120 + 12 * k, fx = 3 * ft, fy = fx + 1, fz = fx + 2, my = 7 + 3 * ft;
      Gmask[r0 + 0][u0 + fx] = Gmask[r0 + 0][u0 + fz] = 1; Gmask[r0 + 1][u0 + fx] = Gmask[r0 + 1][u0 + fz] = 1;
      Gmask[r0 + 2][u0 + fy] = Gmask[r0 + 2][u0 + fz] = 1; Gmask[r0 + 3][u0 + fy] = Gmask[r0 + 3][u0 + fz] = 1;
      Gmask[r0 + 4][u0 + fz] = Gmask[r0 + 4][u0 + my] = 1; Gmask[r0 + 5][u0 + fz] = Gmask[r0 + 5][u0 + my] = 1;
      Gmask[r0 + 6][u0 + fz] = 1; Gmask[r0 + 7][u0 + fz] = 1;
    }
  }
  int na = 0, ng = 0;
  for (int j = 0; j < NZ; ++j) {
    a_colptr[j] = na; g_colptr[j] = ng;
    for (int i = 0; i < NEQ; ++i) if (Amask[i][j]) a_row[na++] = i;
    for (int i = 0; i < NINEQ; ++i) if (Gmask[i][j]) g_row[ng++] = i;
  }
  a_colptr[NZ] = na; g_colptr[NZ] = ng;
}

/* ------------------------------------------------------------------------------------------
 * sparse LDL^T of the permuted KKT: symbolic part (done once)
 * ---------------------------------------------------------------------------------------- */
/* KKT upper-triangular entries in the permuted ordering, CSC */
static int k_perm[NKKT], k_pinv[NKKT];
static int k_nnz, *k_colptr, *k_rowidx, *k_srckind, *k_srcidx;   /* value sources */
static int l_parent[NKKT], l_nz[NKKT], l_colptr[NKKT + 1], l_nnz;
enum { SRC_H = 0, SRC_G, SRC_A, SRC_SZ, SRC_ONE, SRC_MDELTA };

static void build_ordering(void) {
  /* variable ids: dx j -> j ; ds i -> 240+i ; dz i -> 400+i ; dy i -> 560+i.
     stage-interleaved elimination order: ds_k, dz_k, u_k, dy_mx_k, dy_dyn_k, x_{k+1} */
  int n = 0;
  for (int k = 0; k < NH; ++k) {
    for (int i = 0; i < 16; ++i) k_perm[n++] = 240 + 16 * k + i;
    for (int i = 0; i < 16; ++i) k_perm[n++] = 400 + 16 * k + i;
    for (int c = 0; c < NU; ++c) k_perm[n++] = 120 + 12 * k + c;
    k_perm[n++] = 560 + 120 + 2 * k; k_perm[n++] = 560 + 121 + 2 * k;
    for (int r = 0; r < NX; ++r) k_perm[n++] = 560 + 12 * k + r;
    for (int c = 0; c < NX; ++c) k_perm[n++] = 12 * k + c;
  }
  for (int i = 0; i < NKKT; ++i) k_pinv[k_perm[i]] = i;
}

typedef struct { int r, c, kind, idx; } Trip;
static int trip_cmp(const void* a, const void* b) {
  const Trip *x = (const Trip*)a, *y = (const Trip*)b;
  if (x->c != y->c) return x->c - y->c;
  return x->r - y->r;
}

static void build_kkt_symbolic(void) {
  const int cap = NZ + NNZ_G + NNZ_A + 3 * NINEQ + NEQ;
  Trip* t = (Trip*)malloc(sizeof(Trip) * cap);
  int n = 0;
#define ADD(i, j, kd, ix) do { int pi = k_pinv[i], pj = k_pinv[j]; Trip q; q.r = pi < pj ? pi : pj; q.c = pi < pj ? pj : pi; q.kind = kd; q.idx = ix; t[n++] = q; } while (0)
  for (int j = 0; j < NZ; ++j) ADD(j, j, SRC_H, j);                                   /* H + beta I */
  for (int j = 0; j < NZ; ++j) for (int p = g_colptr[j]; p < g_colptr[j + 1]; ++p) ADD(400 + g_row[p], j, SRC_G, p);
  for (int j = 0; j < NZ; ++j) for (int p = a_colptr[j]; p < a_colptr[j + 1]; ++p) ADD(560 + a_row[p], j, SRC_A, p);
  for (int i = 0; i < NINEQ; ++i) { ADD(240 + i, 240 + i, SRC_SZ, i); ADD(400 + i, 240 + i, SRC_ONE, 0); ADD(400 + i, 400 + i, SRC_MDELTA, 0); }
  for (int i = 0; i < NEQ; ++i) ADD(560 + i, 560 + i, SRC_MDELTA, 0);
#undef ADD
  qsort(t, n, sizeof(Trip), trip_cmp);
  k_nnz = n;
  k_colptr = (int*)calloc(NKKT + 1, sizeof(int));
  k_rowidx = (int*)malloc(sizeof(int) * n); k_srckind = (int*)malloc(sizeof(int) * n); k_srcidx = (int*)malloc(sizeof(int) * n);
  for (int p = 0; p < n; ++p) { k_rowidx[p] = t[p].r; k_srckind[p] = t[p].kind; k_srcidx[p] = t[p].idx; k_colptr[t[p].c + 1]++; }
  for (int j = 0; j < NKKT; ++j) k_colptr[j + 1] += k_colptr[j];
  free(t);
  /* elimination tree and column counts of L (up-looking symbolic analysis) */
  int flag[NKKT];
  for (int k = 0; k < NKKT; ++k) {
    l_parent[k] = -1; flag[k] = k; l_nz[k] = 0;
    for (int p = k_colptr[k]; p < k_colptr[k + 1]; ++p) {
      int i = k_rowidx[p];
      if (i < k) {
        for (; flag[i] != k; i = l_parent[i]) {
          if (l_parent[i] == -1) l_parent[i] = k;
          l_nz[i]++;
          flag[i] = k;
        }
      }
    }
  }
  l_colptr[0] = 0;
  for (int k = 0; k < NKKT; ++k) l_colptr[k + 1] = l_colptr[k] + l_nz[k];
  l_nnz = l_colptr[NKKT];
}

static void ensure_ready(void) {
#pragma omp critical(srbd_oracle_init)
  {
    if (!g_ready) { build_patterns(); build_ordering(); build_kkt_symbolic(); g_ready = 1; }
  }
}

int srbd_oracle_kkt_fill(void) { ensure_ready(); return l_nnz; }
int srbd_oracle_pattern(int which, const int** colptr, const int** row) {
  ensure_ready();
  if (which == 'A') { *colptr = a_colptr; *row = a_row; return NNZ_A; }
  if (which == 'G') { *colptr = g_colptr; *row = g_row; return NNZ_G; }
  return -1;
}

/* ------------------------------------------------------------------------------------------
 * numeric LDL^T (up-looking) and solve, per environment
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  REAL *Kx, *Lx, *D, *Y;
  int *Li, *Lnz, *Pattern, *Flag;
} LdlWork;

static void ldl_alloc(LdlWork* w) {
  w->Kx = (REAL*)malloc(sizeof(REAL) * k_nnz); w->Lx = (REAL*)malloc(sizeof(REAL) * (l_nnz + 1));
  w->D = (REAL*)malloc(sizeof(REAL) * NKKT); w->Y = (REAL*)malloc(sizeof(REAL) * NKKT);
  w->Li = (int*)malloc(sizeof(int) * (l_nnz + 1)); w->Lnz = (int*)malloc(sizeof(int) * NKKT);
  w->Pattern = (int*)malloc(sizeof(int) * NKKT); w->Flag = (int*)malloc(sizeof(int) * NKKT);
}
static void ldl_free(LdlWork* w) {
  free(w->Kx); free(w->Lx); free(w->D); free(w->Y); free(w->Li); free(w->Lnz); free(w->Pattern); free(w->Flag);
}

static void ldl_numeric(LdlWork* w) {
  REAL* Y = w->Y; int* Pattern = w->Pattern; int* Flag = w->Flag; int* Lnz = w->Lnz;
  for (int k = 0; k < NKKT; ++k) {
    Y[k] = 0; int top = NKKT; Flag[k] = k; Lnz[k] = 0;
    for (int p = k_colptr[k]; p < k_colptr[k + 1]; ++p) {
      int i = k_rowidx[p];
      Y[i] += w->Kx[p];
      int len = 0;
      for (; Flag[i] != k; i = l_parent[i]) { Pattern[len++] = i; Flag[i] = k; }
      while (len > 0) Pattern[--top] = Pattern[--len];
    }
    w->D[k] = Y[k]; Y[k] = 0;
    for (; top < NKKT; ++top) {
      const int i = Pattern[top];
      const REAL yi = Y[i]; Y[i] = 0;
      const int p2 = l_colptr[i] + Lnz[i];
      for (int p = l_colptr[i]; p < p2; ++p) Y[w->Li[p]] -= w->Lx[p] * yi;
      const REAL lki = yi / w->D[i];
      w->D[k] -= lki * yi;
      w->Li[p2] = k; w->Lx[p2] = lki; Lnz[i]++;
    }
  }
}

/* x (original order) <- K^-1 b (original order) */
static void ldl_solve(const LdlWork* w, const REAL* b, REAL* x) {
  REAL t[NKKT];
  for (int i = 0; i < NKKT; ++i) t[i] = b[k_perm[i]];
  for (int j = 0; j < NKKT; ++j) { const int p2 = l_colptr[j] + w->Lnz[j]; for (int p = l_colptr[j]; p < p2; ++p) t[w->Li[p]] -= w->Lx[p] * t[j]; }
  for (int j = 0; j < NKKT; ++j) t[j] /= w->D[j];
  for (int j = NKKT - 1; j >= 0; --j) { const int p2 = l_colptr[j] + w->Lnz[j]; for (int p = l_colptr[j]; p < p2; ++p) t[j] -= w->Lx[p] * t[w->Li[p]]; }
  for (int i = 0; i < NKKT; ++i) x[k_perm[i]] = t[i];
}

static REAL ratio_step(const REAL* v, const REAL* dv, int n) {
  /* alpha = max(min(1, 0.99 * min_i{dv_i < 0 ? -v_i/dv_i : 1}), 1e-12)   (sparse_pdipm_solver.py:460-467) */
  REAL mn = (dv[0] < 0) ? -v[0] / dv[0] : (REAL)1.0;
  for (int i = 1; i < n; ++i) { const REAL c = (dv[i] < 0) ? -v[i] / dv[i] : (REAL)1.0; if (c < mn) mn = c; }
  return RFMAX(RFMIN((REAL)1.0, (REAL)0.99 * mn), (REAL)1e-12);
}

static void pdipm_env(const double* Qv, const double* Gv, const double* Av, const double* f, const double* h,
                      const double* b, const double* x0, const double* s0, const double* z0, const double* y0,
                      double* xo, double* so, double* zo, double* yo, double* res, double* muo,
                      int iters, LdlWork* w, double tol, int* it_done) {
  REAL x[NZ], s[NINEQ], z[NINEQ], y[NEQ], rx[NZ], rs[NINEQ], re[NEQ];
  REAL rhs[NKKT], aff[NKKT], cor[NKKT], ds[NINEQ], dz[NINEQ];
  for (int i = 0; i < NZ; ++i) x[i] = x0[i];
  for (int i = 0; i < NINEQ; ++i) { s[i] = s0[i]; z[i] = z0[i]; }
  for (int i = 0; i < NEQ; ++i) y[i] = y0[i];
  REAL mu_new = 0;
  int it = 0;
  for (; it < iters; ++it) {
    /* residuals (:392-402) */
    for (int j = 0; j < NZ; ++j) rx[j] = (REAL)Qv[j] * x[j] + (REAL)f[j];
    for (int i = 0; i < NINEQ; ++i) rs[i] = s[i] - (REAL)h[i];
    for (int i = 0; i < NEQ; ++i) re[i] = -(REAL)b[i];
    for (int j = 0; j < NZ; ++j) {
      for (int p = g_colptr[j]; p < g_colptr[j + 1]; ++p) { rx[j] += (REAL)Gv[p] * z[g_row[p]]; rs[g_row[p]] += (REAL)Gv[p] * x[j]; }
      for (int p = a_colptr[j]; p < a_colptr[j + 1]; ++p) { rx[j] += (REAL)Av[p] * y[a_row[p]]; re[a_row[p]] += (REAL)Av[p] * x[j]; }
    }
    REAL mu = 0;
    for (int i = 0; i < NINEQ; ++i) mu += s[i] * z[i];
    mu /= (REAL)NINEQ;
    /* KKT values (:406-439) */
    for (int p = 0; p < k_nnz; ++p) {
      REAL v;
      switch (k_srckind[p]) {
        case SRC_H: v = (REAL)Qv[k_srcidx[p]] + (REAL)BETA; break;
        case SRC_G: v = (REAL)Gv[k_srcidx[p]]; break;
        case SRC_A: v = (REAL)Av[k_srcidx[p]]; break;
        case SRC_SZ: v = ((REAL)1.0 / s[k_srcidx[p]]) * z[k_srcidx[p]] + (REAL)DELTA; break;
        case SRC_ONE: v = 1; break;
        default: v = -(REAL)DELTA; break;
      }
      w->Kx[p] = v;
    }
    ldl_numeric(w);
    /* affine direction (:443-457) */
    for (int j = 0; j < NZ; ++j) rhs[j] = -rx[j];
    for (int i = 0; i < NINEQ; ++i) { rhs[240 + i] = -(((REAL)1.0 / s[i]) * (s[i] * z[i])); rhs[400 + i] = -rs[i]; }
    for (int i = 0; i < NEQ; ++i) rhs[560 + i] = -re[i];
    ldl_solve(w, rhs, aff);
    const REAL a_pri_aff = ratio_step(s, aff + 240, NINEQ), a_dual_aff = ratio_step(z, aff + 400, NINEQ);
    REAL mu_aff = 0;
    for (int i = 0; i < NINEQ; ++i) mu_aff += (s[i] + a_pri_aff * aff[240 + i]) * (z[i] + a_dual_aff * aff[400 + i]);
    mu_aff /= (REAL)NINEQ;
    const REAL sg = mu_aff / mu, sigma = sg * sg * sg;                                    /* :474 */
    /* corrector (:478-486) */
    for (int j = 0; j < NKKT; ++j) rhs[j] = 0;
    for (int i = 0; i < NINEQ; ++i) rhs[240 + i] = -(((REAL)1.0 / s[i]) * (s[i] * z[i] + aff[240 + i] * aff[400 + i] - sigma * mu));
    ldl_solve(w, rhs, cor);
    for (int i = 0; i < NINEQ; ++i) { ds[i] = aff[240 + i] + cor[240 + i]; dz[i] = aff[400 + i] + cor[400 + i]; }
    const REAL a_pri = ratio_step(s, ds, NINEQ), a_dual = ratio_step(z, dz, NINEQ);
    for (int j = 0; j < NZ; ++j) x[j] += a_pri * (aff[j] + cor[j]);
    for (int i = 0; i < NEQ; ++i) y[i] += a_dual * (aff[560 + i] + cor[560 + i]);
    mu_new = 0;
    for (int i = 0; i < NINEQ; ++i) {
      s[i] = RFMAX(s[i] + a_pri * ds[i], (REAL)1e-8);
      z[i] = RFMAX(z[i] + a_dual * dz[i], (REAL)1e-8);
      mu_new += s[i] * z[i];
    }
    mu_new /= (REAL)NINEQ;
    /* tolerance mode: the criterion of generate_solver_function.py:187 on the quantities one call returns --
     * the residual norms of THIS iteration's starting point (:524-529) and the post-step mu (:530) */
    if (tol > 0) {
      REAL tx = 0, ts = 0, te = 0;
      for (int j = 0; j < NZ; ++j) tx += rx[j] * rx[j];
      for (int i = 0; i < NINEQ; ++i) ts += rs[i] * rs[i];
      for (int i = 0; i < NEQ; ++i) te += re[i] * re[i];
      if ((double)RSQRT(tx) < tol && (double)RSQRT(ts) < tol && (double)RSQRT(te) < tol && (double)mu_new < tol) { ++it; break; }
    }
  }
  if (it_done) *it_done = it;
  REAL nx = 0, ns = 0, ne = 0;
  if (iters > 0) {
    for (int j = 0; j < NZ; ++j) nx += rx[j] * rx[j];
    for (int i = 0; i < NINEQ; ++i) ns += rs[i] * rs[i];
    for (int i = 0; i < NEQ; ++i) ne += re[i] * re[i];
  }
  for (int i = 0; i < NZ; ++i) xo[i] = (double)x[i];
  for (int i = 0; i < NINEQ; ++i) { so[i] = (double)s[i]; zo[i] = (double)z[i]; }
  for (int i = 0; i < NEQ; ++i) yo[i] = (double)y[i];
  res[0] = (double)RSQRT(nx); res[1] = (double)RSQRT(ns); res[2] = (double)RSQRT(ne); res[3] = (double)mu_new;
  muo[0] = (double)mu_new;
}

/* in10: Q_val G_val A_val f h b x s z y ; out6: x s z y residuals(4) mu(1) ; (B, nnz) row-major.
 * cold_start != 0: x=0, s=max(h - G*0, 1), z=1, y=1 (mpc_controller_cusadi.py:138-141). */
static int pdipm_batch(const double* const* in10, double* const* out6, int B, int iters, int cold_start, int nthreads,
                       double tol, int* iters_done);
int FN(srbd_oracle_pdipm)(const double* const* in10, double* const* out6, int B, int iters, int cold_start, int nthreads) {
  return pdipm_batch(in10, out6, B, iters, cold_start, nthreads, 0.0, NULL);
}
/* the same with a convergence tolerance (runtime extension of this repo's kernels; criterion of
 * generate_solver_function.py:187): env e stops after the first iteration whose (|rx|, |rs|, |re|, mu) are all < tol,
 * at most `iters`; iters_done[e] = iterations executed. */
int FN(srbd_oracle_pdipm_tol)(const double* const* in10, double* const* out6, int* iters_done, int B, int iters, double tol,
                              int cold_start, int nthreads) {
  return pdipm_batch(in10, out6, B, iters, cold_start, nthreads, tol, iters_done);
}
static int pdipm_batch(const double* const* in10, double* const* out6, int B, int iters, int cold_start, int nthreads,
                       double tol, int* iters_done) {
  ensure_ready();
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    LdlWork w; ldl_alloc(&w);
    double cx[NZ], cs[NINEQ], cz[NINEQ], cy[NEQ];
#pragma omp for schedule(dynamic, 1)
    for (int e = 0; e < B; ++e) {
      const double *x0, *s0, *z0, *y0;
      const double* h = in10[4] + (size_t)e * NINEQ;
      if (cold_start) {
        for (int i = 0; i < NZ; ++i) cx[i] = 0;
        for (int i = 0; i < NINEQ; ++i) { cs[i] = h[i] > 1.0 ? h[i] : 1.0; cz[i] = 1.0; }
        for (int i = 0; i < NEQ; ++i) cy[i] = 1.0;
        x0 = cx; s0 = cs; z0 = cz; y0 = cy;
      } else {
        x0 = in10[6] + (size_t)e * NZ; s0 = in10[7] + (size_t)e * NINEQ; z0 = in10[8] + (size_t)e * NINEQ; y0 = in10[9] + (size_t)e * NEQ;
      }
      pdipm_env(in10[0] + (size_t)e * NNZ_H, in10[1] + (size_t)e * NNZ_G, in10[2] + (size_t)e * NNZ_A,
                in10[3] + (size_t)e * NZ, h, in10[5] + (size_t)e * NEQ, x0, s0, z0, y0,
                out6[0] + (size_t)e * NZ, out6[1] + (size_t)e * NINEQ, out6[2] + (size_t)e * NINEQ,
                out6[3] + (size_t)e * NEQ, out6[4] + (size_t)e * 4, out6[5] + (size_t)e, iters, &w, tol,
                iters_done ? iters_done + e : NULL);
    }
    ldl_free(&w);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * qp_former (srbd_constraints.py:20-227), literal: RK4 of the affine dynamics, jacobian by
 * the affine-map algebra, b = A z - eq(z), d = G z - ineq(z)
 * ---------------------------------------------------------------------------------------- */
static void matmul12(const REAL* A, const REAL* B, REAL* C, int bc) {   /* C[12][bc] = A[12][12] B[12][bc] */
  for (int i = 0; i < 12; ++i) for (int j = 0; j < bc; ++j) { REAL a = 0; for (int k = 0; k < 12; ++k) a += A[12 * i + k] * B[bc * k + j]; C[bc * i + j] = a; }
}
static void fdyn(const REAL* Ac, const REAL* Bc, const REAL* c, const REAL* x, const REAL* u, REAL* out) {
  for (int i = 0; i < 12; ++i) { REAL a = c[i]; for (int k = 0; k < 12; ++k) a += Ac[12 * i + k] * x[k] + Bc[12 * i + k] * u[k]; out[i] = a; }
}

static void qp_env(const double* const* in, double* const* out, size_t e) {
  const double* x0 = in[0] + e * 12; const double* xl = in[1] + e * 120; const double* ul = in[2] + e * 120;
  const double* xref = in[3] + e * 120; const REAL dt = in[4][e], m = in[5][e], mu = in[6][e];
  const double* Rf = in[7] + e * 9; const double* If = in[8] + e * 9; const double* body = in[9] + e * 3;
  const double* lf = in[10] + e * 3; const double* rf = in[11] + e * 3; const double* ct = in[12] + e * 20;
  const double* Q = in[13] + e * 12; const double* R = in[14] + e * 12; const double* rl = in[15] + e * 3; const double* ra = in[16] + e * 3;
  double* Hv = out[0] + e * NNZ_H; double* fo = out[1] + e * NZ; double* Avo = out[2] + e * NNZ_A;
  double* bo = out[3] + e * NEQ; double* Gvo = out[4] + e * NNZ_G; double* dout = out[5] + e * NINEQ;
  /* 3x3 matrices are column-major in nnz order: M(i,j) = flat[i + 3j] */
  REAL Rb[3][3], Iw[3][3], Ii[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Rb[i][j] = Rf[i + 3 * j]; Iw[i][j] = If[i + 3 * j]; }
  const REAL det = Iw[0][0] * (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) - Iw[0][1] * (Iw[1][0] * Iw[2][2] - Iw[1][2] * Iw[2][0])
                 + Iw[0][2] * (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    const int j1 = (j + 1) % 3, j2 = (j + 2) % 3, i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    Ii[i][j] = (Iw[j1][i1] * Iw[j2][i2] - Iw[j1][i2] * Iw[j2][i1]) / det;
  }
  REAL Ac[144], Bc[144], c[12];
  memset(Ac, 0, sizeof Ac); memset(Bc, 0, sizeof Bc); memset(c, 0, sizeof c);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Ac[12 * i + 6 + j] = Rb[i][j]; Ac[12 * (3 + i) + 9 + i] = 1; }
  for (int ft = 0; ft < 2; ++ft) {
    const double* fp = ft ? rf : lf;
    const REAL r[3] = {(REAL)fp[0] - (REAL)body[0], (REAL)fp[1] - (REAL)body[1], (REAL)fp[2] - (REAL)body[2]};
    const REAL S[3][3] = {{0, -r[2], r[1]}, {r[2], 0, -r[0]}, {-r[1], r[0], 0}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      REAL a = 0; for (int k = 0; k < 3; ++k) a += Ii[i][k] * S[k][j];
      Bc[12 * (6 + i) + 3 * ft + j] = a;
      Bc[12 * (6 + i) + 6 + 3 * ft + j] = Ii[i][j];
    }
    for (int i = 0; i < 3; ++i) Bc[12 * (9 + i) + 3 * ft + i] = (REAL)1.0 / m;
  }
  for (int i = 0; i < 3; ++i) { c[6 + i] = ra[i]; c[9 + i] = (REAL)(i == 2 ? GRAV : 0.0) + (REAL)rl[i]; }
  /* RK4 on the affine-map coefficients (srbd_centroidal_model.py:110-115) */
  REAL Kx[4][144], Ku[4][144], kc[4][12], T[144], Tu[144], tc[12];
  memcpy(Kx[0], Ac, sizeof Ac); memcpy(Ku[0], Bc, sizeof Bc); memcpy(kc[0], c, sizeof c);
  for (int st = 1; st < 4; ++st) {
    const REAL hstep = (st == 3) ? dt : dt / (REAL)2;
    for (int i = 0; i < 144; ++i) { T[i] = hstep * Kx[st - 1][i]; Tu[i] = hstep * Ku[st - 1][i]; }
    for (int i = 0; i < 12; ++i) { T[13 * i] += 1; tc[i] = hstep * kc[st - 1][i]; }
    matmul12(Ac, T, Kx[st], 12);
    matmul12(Ac, Tu, Ku[st], 12);
    for (int i = 0; i < 144; ++i) Ku[st][i] += Bc[i];
    for (int i = 0; i < 12; ++i) { REAL a = c[i]; for (int k = 0; k < 12; ++k) a += Ac[12 * i + k] * tc[k]; kc[st][i] = a; }
  }
  REAL Ad[144], Bd[144];
  const REAL s6 = dt / (REAL)6;
  for (int i = 0; i < 144; ++i) {
    Ad[i] = s6 * (Kx[0][i] + 2 * Kx[1][i] + 2 * Kx[2][i] + Kx[3][i]);
    Bd[i] = s6 * (Ku[0][i] + 2 * Ku[1][i] + 2 * Ku[2][i] + Ku[3][i]);
  }
  for (int i = 0; i < 12; ++i) Ad[13 * i] += 1;
  /* dense A, eq(z), G, ineq(z) */
  static _Thread_local REAL A[NEQ][NZ], G[NINEQ][NZ];
  REAL eq[NEQ], ineq[NINEQ], z[NZ], xcat[132];
  memset(A, 0, sizeof A); memset(G, 0, sizeof G);
  for (int i = 0; i < 120; ++i) { z[i] = xl[i]; z[120 + i] = ul[i]; }
  for (int i = 0; i < 12; ++i) xcat[i] = x0[i];
  for (int i = 0; i < 120; ++i) xcat[12 + i] = xl[i];
  for (int k = 0; k < NH; ++k) {
    for (int r = 0; r < 12; ++r) {
      A[12 * k + r][12 * k + r] += 1;
      for (int cc = 0; cc < 12; ++cc) { if (k >= 1) A[12 * k + r][12 * (k - 1) + cc] -= Ad[12 * r + cc]; A[12 * k + r][120 + 12 * k + cc] -= Bd[12 * r + cc]; }
    }
    /* literal numeric RK4 step */
    REAL xi[12], ui[12], k1[12], k2[12], k3[12], k4[12], tmp[12];
    for (int i = 0; i < 12; ++i) { xi[i] = xcat[12 * k + i]; ui[i] = ul[12 * k + i]; }
    fdyn(Ac, Bc, c, xi, ui, k1);
    for (int i = 0; i < 12; ++i) tmp[i] = xi[i] + dt / 2 * k1[i];
    fdyn(Ac, Bc, c, tmp, ui, k2);
    for (int i = 0; i < 12; ++i) tmp[i] = xi[i] + dt / 2 * k2[i];
    fdyn(Ac, Bc, c, tmp, ui, k3);
    for (int i = 0; i < 12; ++i) tmp[i] = xi[i] + dt * k3[i];
    fdyn(Ac, Bc, c, tmp, ui, k4);
    for (int i = 0; i < 12; ++i) eq[12 * k + i] = xcat[12 * (k + 1) + i] - (xi[i] + dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]));
    A[120 + 2 * k][120 + 12 * k + 6] = 1; A[121 + 2 * k][120 + 12 * k + 9] = 1;
    eq[120 + 2 * k] = ui[6]; eq[121 + 2 * k] = ui[9];
    for (int ft = 0; ft < 2; ++ft) {
      const int r0 = 16 * k + 8 * ft, u0 = 120 + 12 * k, fx = 3 * ft, fy = fx + 1, fz = fx + 2, my = 7 + 3 * ft;
      const REAL contact = ct[k + 10 * ft];                      /* column-major (10,2) */
      const int cols[8][2] = {{fx, fz}, {fx, fz}, {fy, fz}, {fy, fz}, {fz, my}, {fz, my}, {fz, -1}, {fz, -1}};
      const REAL coef[8][2] = {{-1, -mu}, {1, -mu}, {-1, -mu}, {1, -mu}, {-(REAL)LT, -1}, {-(REAL)LH, 1}, {-1, 0}, {1, 0}};
      for (int j = 0; j < 8; ++j) {
        REAL v = 0;
        for (int q = 0; q < 2; ++q) if (cols[j][q] >= 0) { G[r0 + j][u0 + cols[j][q]] = coef[j][q]; v += coef[j][q] * ui[cols[j][q]]; }
        if (j == 7) v -= (REAL)F_MAX * contact;
        ineq[r0 + j] = v;
      }
    }
  }
  for (int i = 0; i < NEQ; ++i) { REAL a = 0; for (int j = 0; j < NZ; ++j) a += A[i][j] * z[j]; bo[i] = (double)(a - eq[i]); }
  for (int i = 0; i < NINEQ; ++i) { REAL a = 0; for (int j = 0; j < NZ; ++j) a += G[i][j] * z[j]; dout[i] = (double)(a - ineq[i]); }
  for (int j = 0; j < NZ; ++j) {
    const REAL hj = j < 120 ? (REAL)Q[j % 12] : (REAL)R[j % 12];
    const REAL grad = j < 120 ? hj * (z[j] - (REAL)xref[j]) : hj * z[j];
    Hv[j] = (double)hj;
    fo[j] = (double)(grad - hj * z[j]);
    for (int p = a_colptr[j]; p < a_colptr[j + 1]; ++p) Avo[p] = (double)A[a_row[p]][j];
    for (int p = g_colptr[j]; p < g_colptr[j + 1]; ++p) Gvo[p] = (double)G[g_row[p]][j];
  }
}

int FN(srbd_oracle_qp_former)(const double* const* in17, double* const* out6, int B, int nthreads) {
  ensure_ready();
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B; ++e) qp_env(in17, out6, (size_t)e);
  return 0;
}
