// Plain argument structs of the kernels (no CUDA types: also used by the host-side
// warp emulator that runs the PDIPM device code on the CPU in tests).
#pragma once
#include "srbd_pattern.h"

namespace srbd {

struct QpFormArgs {
  const double* in[QP_N_IN];
  double* out[QP_N_OUT];
};

// Arguments of the PDIPM kernel.  in[] / out[] follow sparse_pdipm_solver.py:532-534.
struct PdipmArgs {
  const double* in[IP_N_IN];    // Q_val, G_val, A_val, f, h, b, x, s, z, y   (B, nnz) each
  double* out[IP_N_OUT];        // x, s, z, y, residuals(4), mu(1)
  double* work;                 // (B, work_stride) scratch owned by the caller
  int work_stride;              // doubles per env: PDIPM_WORK_DOUBLES_FAST (fast kernel) or PDIPM_WORK_DOUBLES
  int* inv_flag;                // optional (fast kernel, CCS inputs): the kernel verifies that the env's data is
                                //   stage-invariant (what the synchronous probe checks) and sets *inv_flag = 1 otherwise
  int* iters_done;              // optional (B) per-env Newton iterations executed
  int num_iters;                // Newton iterations per call (runtime, reference bakes 5)
  double tol;                   // <= 0: fixed iteration count; > 0: stop when rx,rs,re,mu < tol
  int precision;                // 0: fp64 throughout (the parity path).  Fast kernel only -- 1: the ten Schur pivot blocks are
                                //   formed, coupled, inverted and stored in fp32; 2 / 3: the same plus one / two steps of
                                //   iterative refinement per solve against the fp64 operator
  int cold_start;               // 1: ignore in[x,s,z,y]; start from x=0, s=max(h,1), z=1, y=1
                                //    (mpc_controller_cusadi.py:138-141)
  // ---- fused tick (fast kernel only): assemble from the 17 qp_former inputs instead of reading
  //      the CCS vectors in[0..5], and turn u0 into the body-frame wrench at the end
  int fused;                    // 1: qp_in[] is used, in[0..5] are ignored
  const double* qp_in[QP_N_IN]; // qp_former inputs (x, u may be null = ones)
  int warm_xy;                  // with cold_start, an extension (the reference cold-starts every tick,
                                //   mpc_controller_cusadi.py:138-141).  1: keep (x, y) found in out[] (the previous tick's
                                //   solution), re-initialise s = max(h - G x, 1), z = 1.  2 (fused tick): keep (x, y, z),
                                //   shifted by one stage when the contact schedule rolled on; s = max(h - G x, 1e-2); z
                                //   floored at 1e-3, re-centred (z = 0.1 / s) on the rows whose bound the schedule changed
  const int* reset;             // optional (B): envs with reset[e] != 0 are cold-started although warm_xy is set
  const float* rot;             // optional (B,3,3) fp32 row-major state-estimate rotation
  float* wrench;                // optional (B,2,6) fp32 (mpc_controller_cusadi.py:186-203)
};
constexpr int PDIPM_WORK_DOUBLES = 6224;  // general kernel, per env: directions + per-row barrier data (1 344) and the ten
                                          // stage buffers prepared once per call (10 x 488)
constexpr int PDIPM_WORK_DOUBLES_FAST = 2048;   // fast kernel (SRBD_STAGE_INVARIANT, fused tick): srbd_pdipm_fast.h FW_END; the
                                                // fp64 path touches the first 1 664

}  // namespace srbd
