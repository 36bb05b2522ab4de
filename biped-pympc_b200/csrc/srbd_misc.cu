// Small kernels around the solve: u0 -> body-frame foot wrench (the post-processing of
// biped_pympc/convex_mpc/mpc_controller_cusadi.py:186-203) and an FP64 FMA throughput
// micro-benchmark that provides the roofline denominator for the PDIPM kernel.
#include <cuda_runtime.h>

#include "srbd_kernels.h"
#include "srbd_pdipm_fast.h"
#include "srbd_stack_impl.h"

namespace srbd {

// wrench[e] = (2, 6) fp32: [-R' f_L, -R' m_L ; -R' f_R, -R' m_R] with m_x zeroed and the cast to
// fp32 done *before* the rotation, exactly as the reference does.  R is the row-major
// state-estimate rotation (B, 3, 3) in fp32.
__global__ void wrench_kernel(const double* __restrict__ x_sol, const float* __restrict__ rot,
                              float* __restrict__ wrench, int B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  const double* u0 = x_sol + (long long)e * NZ + NX * N;
  float R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = rot[(long long)e * 9 + i];
  float* o = wrench + (long long)e * 12;
#pragma unroll
  for (int idx = 0; idx < 12; ++idx) o[idx] = wrench_entry(u0, R, idx);   // shared with the fused tick
}

cudaError_t launch_wrench(const double* x_sol, const float* rot, float* wrench, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  wrench_kernel<<<(B + 127) / 128, 128, 0, stream>>>(x_sol, rot, wrench, B);
  return cudaGetLastError();
}

// Probe used by the synchronous drop-in entry: does every env of the batch carry qp_former-shaped
// data (Q / G / A values identical for every horizon stage, unit coefficients on x_{k+1} and on
// the Mx rows)?  One warp per env, 1 716 doubles read per env; *flag is OR-ed with 1 otherwise.
__global__ void stage_invariant_probe_kernel(const double* __restrict__ Q, const double* __restrict__ G,
                                             const double* __restrict__ A, int* flag, int B) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const double* q = Q + (long long)warp * NNZ_H;
  const double* g = G + (long long)warp * NNZ_G;
  const double* a = A + (long long)warp * NNZ_A;
  int bad = 0;
  for (int i = lane; i < 36 * 8; i += 32) bad |= (a[36 + i] != a[i % 36]);              // x-columns, stages 1..8
  for (int c = lane; c < 12; c += 32) bad |= (a[324 + c] != 1.0) | (a[a_xcol_off(c)] != 1.0);   // E
  for (int i = lane; i < 86 * 9; i += 32) bad |= (a[336 + 86 + i] != a[336 + i % 86]);  // u-columns
  if (lane < 2) bad |= (a[336 + a_ucol_off(6 + 3 * lane) + A_UCOL_MX_POS] != 1.0);      // Mx rows
  for (int i = lane; i < 28 * 9; i += 32) bad |= (g[28 + i] != g[i % 28]);
  for (int i = lane; i < 108; i += 32) bad |= (q[12 + i] != q[i % 12]) | (q[132 + i] != q[120 + i % 12]);
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1);
}

cudaError_t launch_stage_invariant_probe(const double* Q, const double* G, const double* A, int* flag, int B,
                                         cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  stage_invariant_probe_kernel<<<(B + 7) / 8, 256, 0, stream>>>(Q, G, A, flag, B);
  return cudaGetLastError();
}

// Input builder: the marshalling of mpc_controller_cusadi.py:54-95 (casts to fp64, flattening,
// per-env replication of the constants) as one launch instead of ~25 small torch kernels.  One warp per
// env.  Bit-exact: every output is an fp32 -> fp64 cast of one controller value.  Layout switches
// follow SURVEY.md 8b: fix_layout = 0 reproduces the reference's row-major flattening of R_body and of
// the (10,2) contact table into buffers the kernels read column-major (Q1, Q2); q_stride_bug = 1
// reproduces the 13-long Q strided by 12 (Q3).
struct MarshalArgs {
  const float *x0, *x_ref, *R_body, *I_world, *root_pos, *foot_pos, *contact, *dt_mpc, *Q, *R, *res_lin, *res_ang;
  const float *I_body, *rot;   // non-null: I_world = rot I_body rot' is computed here (mpc_controller_cusadi.py:60-63), I_world is ignored
  int nQ, fix_layout, q_stride_bug;
  double mass, mu;
  double* out[QP_N_IN];   // out[IN_X] / out[IN_U] may be null (not written: the kernels treat null as ones)
};
__global__ void marshal_kernel(MarshalArgs a, int B) {
  const long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= B) return;
  for (int i = lane; i < 120; i += 32) {
    a.out[IN_XREF][e * 120 + i] = (double)a.x_ref[e * 120 + i];
    if (a.out[IN_X]) a.out[IN_X][e * 120 + i] = 1.0;
    if (a.out[IN_U]) a.out[IN_U][e * 120 + i] = 1.0;
  }
  if (lane < 12) {
    a.out[IN_X0][e * 12 + lane] = (double)a.x0[e * 12 + lane];
    const long long qi = a.q_stride_bug ? (e * 12 + lane) % a.nQ : lane;     // Q3: env e reads Qflat[12e + i]
    a.out[IN_Q][e * 12 + lane] = (double)a.Q[qi];
    a.out[IN_R][e * 12 + lane] = (double)a.R[lane];
  }
  if (lane < 9) {   // flat[k]: fix_layout ? column-major nnz order of the true matrix : row-major as is
    const int src = a.fix_layout ? (lane % 3) * 3 + lane / 3 : lane;
    a.out[IN_RBODY][e * 9 + lane] = (double)a.R_body[e * 9 + src];
    float iw;
    if (a.I_body) {   // entry (i, j) of R I_body R' in fp32, two chained matrix products like the reference's `R @ I_body @ R.T`
      const int i = src / 3, j = src - 3 * i;
      const float* Rm = a.rot + e * 9;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float t = fmaf(Rm[3 * i + 2], a.I_body[6 + k], fmaf(Rm[3 * i + 1], a.I_body[3 + k], Rm[3 * i] * a.I_body[k]));   // (R I_body)[i][k]
        acc = (k == 0) ? t * Rm[3 * j] : fmaf(t, Rm[3 * j + k], acc);
      }
      iw = acc;
    } else {
      iw = a.I_world[e * 9 + src];
    }
    a.out[IN_IWORLD][e * 9 + lane] = (double)iw;
  }
  if (lane < 20) {  // contact (10,2) row-major; fix_layout: flat[k] = contact[k % 10, k / 10]
    const int src = a.fix_layout ? (lane % 10) * 2 + lane / 10 : lane;
    a.out[IN_CONTACT][e * 20 + lane] = (double)a.contact[e * 20 + src];
  }
  if (lane < 3) {
    a.out[IN_BODY][e * 3 + lane] = (double)a.root_pos[e * 3 + lane];
    a.out[IN_LFOOT][e * 3 + lane] = (double)a.foot_pos[e * 6 + lane];
    a.out[IN_RFOOT][e * 3 + lane] = (double)a.foot_pos[e * 6 + 3 + lane];
    a.out[IN_RESLIN][e * 3 + lane] = (double)a.res_lin[e * 3 + lane];
    a.out[IN_RESANG][e * 3 + lane] = (double)a.res_ang[e * 3 + lane];
  }
  if (lane == 0) {
    a.out[IN_DT][e] = (double)a.dt_mpc[e];
    a.out[IN_M][e] = a.mass;
    a.out[IN_MU][e] = a.mu;
  }
}

cudaError_t launch_marshal(const float* const* f32in12, const float* I_body, const float* rotation_body, int nQ, int fix_layout,
                           int q_stride_bug, double mass, double mu, double* const* out17, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  MarshalArgs a;
  a.x0 = f32in12[0]; a.x_ref = f32in12[1]; a.R_body = f32in12[2]; a.I_world = f32in12[3]; a.root_pos = f32in12[4];
  a.foot_pos = f32in12[5]; a.contact = f32in12[6]; a.dt_mpc = f32in12[7]; a.Q = f32in12[8]; a.R = f32in12[9];
  a.res_lin = f32in12[10]; a.res_ang = f32in12[11];
  a.I_body = I_body; a.rot = rotation_body;
  a.nQ = nQ; a.fix_layout = fix_layout; a.q_stride_bug = q_stride_bug; a.mass = mass; a.mu = mu;
  for (int i = 0; i < QP_N_IN; ++i) a.out[i] = out17[i];
  marshal_kernel<<<(B + 3) / 4, 128, 0, stream>>>(a, B);
  return cudaGetLastError();
}

// State builder: compute_knot_points + set_initial_state + compute_reference_trajectory of
// biped_pympc/convex_mpc/base_controller.py:164-257 (about 40 small fp32 torch kernels per tick in the
// reference) as one launch, one thread per env (srbd_stack_impl.h::reference_trajectory_env): the open-loop knot point
// (world_position_desired, yaw_desired, first_run), x0 and the ten stages of x_ref.
// Arithmetic is fp32 with the reference's operation order and NO fused multiply-add (__fmul_rn /
// __fadd_rn), so every entry equals what the torch expressions give, except the body->world velocity
// (a 3x3 torch.bmm there, whose accumulation order is the BLAS back-end's): (R0 v0 + R1 v1) + R2 v2 here.
__global__ void reference_trajectory_kernel(RefTrajArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < B) reference_trajectory_env(a, e);
}

cudaError_t launch_reference_trajectory(const float* const* f32in9, float step, float* wpd, float* yaw_des,
                                        unsigned char* first_run, float* x0, float* x_ref, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  RefTrajArgs a;
  a.euler = f32in9[0]; a.pos = f32in9[1]; a.omega_w = f32in9[2]; a.vel_w = f32in9[3]; a.rot = f32in9[4];
  a.v_des_b = f32in9[5]; a.w_des_b = f32in9[6]; a.height = f32in9[7]; a.dt_mpc = f32in9[8];
  a.step = step; a.wpd = wpd; a.yaw_des = yaw_des; a.first_run = first_run; a.x0 = x0; a.x_ref = x_ref;
  reference_trajectory_kernel<<<(B + 63) / 64, 64, 0, stream>>>(a, B);
  return cudaGetLastError();
}

// Gait generator: GaitGenerator.get_contact_sub_phase / get_swing_sub_phase / mpc_gait / update_phase of
// biped_pympc/core/gait/gait_generator.py:79-252 (about 60 small torch kernels per control step, with
// boolean-mask indexing that synchronises the host) as one launch, one thread per env.  fp32 with the
// reference's operation order (IEEE division, no FMA): bit-exact.  Outputs are computed from the phase on
// entry; the phase is advanced afterwards when `advance` is set (the order of BipedController.run_lowlevel,
// biped_controller.py:120-127).
__global__ void gait_kernel(GaitArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < B) gait_env(a, e);
}

cudaError_t launch_gait(float* phase, const float* dt_mpc, const int* ssp, const int* dsp, float dt, int advance,
                        int horizon, float* contact_sub, float* swing_sub, int* table_i32, float* table_f32, int B,
                        cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  GaitArgs a;
  a.phase = phase; a.dt_mpc = dt_mpc; a.ssp = ssp; a.dsp = dsp; a.dt = dt; a.advance = advance; a.horizon = horizon;
  a.contact_sub = contact_sub; a.swing_sub = swing_sub; a.table_i32 = table_i32; a.table_f32 = table_f32;
  gait_kernel<<<(B + 127) / 128, 128, 0, stream>>>(a, B);
  return cudaGetLastError();
}

// Host-entry prefetch: pulls one chunk of the 17 qp_former inputs (+ the rotation) from mapped pinned host memory into
// the context's device staging area with plain loads over PCIe.  It needs no shared memory and few registers, so its CTAs
// co-reside with the solve kernel of the previous chunk (which holds one 226 KB CTA per SM) and the transfer hides behind
// that solve -- what the copy engine (cudaMemcpyAsync, 16 calls per chunk) did not achieve.
struct PrefetchArgs {
  const double* src[QP_N_IN];
  double* dst[QP_N_IN];
  int n[QP_N_IN];                 // doubles per array for this chunk (0 = skip)
  const float* rot_src; float* rot_dst; int rot_n;
};
__global__ void __launch_bounds__(128) prefetch_kernel(PrefetchArgs a) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
#pragma unroll 1
  for (int i = 0; i < QP_N_IN; ++i) {
    const double* __restrict__ s = a.src[i];     // 8-byte accesses: no alignment assumption on the caller's buffers
    double* __restrict__ d = a.dst[i];
    for (int k = tid; k < a.n[i]; k += nthr) d[k] = s[k];
  }
  for (int k = tid; k < a.rot_n; k += nthr) a.rot_dst[k] = a.rot_src[k];
}

cudaError_t launch_prefetch(const double* const* src, double* const* dst, const int* n, const float* rot_src, float* rot_dst,
                            int rot_n, int ctas, cudaStream_t stream) {
  PrefetchArgs a;
  for (int i = 0; i < QP_N_IN; ++i) { a.src[i] = src[i]; a.dst[i] = dst[i]; a.n[i] = (src[i] && dst[i]) ? n[i] : 0; }
  a.rot_src = rot_src; a.rot_dst = rot_dst; a.rot_n = rot_n;
  prefetch_kernel<<<ctas, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// 8 independent DFMA chains per thread; 16 flop per inner iteration per thread.
__global__ void dfma_peak_kernel(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

cudaError_t launch_dfma_peak(double* sink, int iters, int blocks, int threads, cudaStream_t stream) {
  dfma_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
  return cudaGetLastError();
}

}  // namespace srbd
