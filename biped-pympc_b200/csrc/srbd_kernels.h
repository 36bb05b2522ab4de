// Internal declarations shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/srbd_mpc.h"
#include "srbd_args.h"
#include "srbd_pattern.h"

namespace srbd {

cudaError_t qp_form_upload_table();
cudaError_t launch_qp_form(const QpFormArgs& args, int B, int num_sms, cudaStream_t stream);

cudaError_t pdipm_upload_tables();
// leave_room: fast fp64 path only -- launch the 12-warp instance (3 per scheduler) when the batch fits one round of it, so
// that a small kernel (the host entry's prefetch) can co-reside: the 16-warp CTA holds every register of its SM
cudaError_t launch_pdipm(const PdipmArgs& args, int B, bool stage_invariant, int num_sms, cudaStream_t stream, bool leave_room = false);
int pdipm_warps_per_cta(bool stage_invariant);

// u0 -> body-frame wrench (mpc_controller_cusadi.py:186-203)
cudaError_t launch_wrench(const double* x_sol, const float* rotation_body_rowmajor, float* wrench,
                          int B, cudaStream_t stream);
// *flag |= 1 unless every env carries stage-invariant, unit-E data (what qp_former produces)
cudaError_t launch_stage_invariant_probe(const double* Q, const double* G, const double* A, int* flag, int B,
                                         cudaStream_t stream);
// controller state (fp32) -> the qp_former inputs (fp64), mpc_controller_cusadi.py:54-95
cudaError_t launch_marshal(const float* const* f32in12, const float* I_body, const float* rotation_body, int nQ, int fix_layout,
                           int q_stride_bug, double mass, double mu, double* const* out17, int B, cudaStream_t stream);
// estimator state + command (fp32) -> x0, x_ref and the open-loop knot point, base_controller.py:164-257
cudaError_t launch_reference_trajectory(const float* const* f32in9, float step, float* wpd, float* yaw_des,
                                        unsigned char* first_run, float* x0, float* x_ref, int B, cudaStream_t stream);
// gait phase -> sub-phases, MPC contact table, phase update (core/gait/gait_generator.py:79-252)
cudaError_t launch_gait(float* phase, const float* dt_mpc, const int* ssp, const int* dsp, float dt, int advance,
                        int horizon, float* contact_sub, float* swing_sub, int* table_i32, float* table_f32, int B,
                        cudaStream_t stream);
// the control step around the MPC tick (srbd_lowlevel.cu); the buffers are the public srbd_loop_buffers
struct UpdateStateArgs;
struct LowLevelArgs;
cudaError_t launch_update_state(const UpdateStateArgs& a, int B, cudaStream_t stream);
cudaError_t launch_lowlevel(const LowLevelArgs& a, int B, cudaStream_t stream);
// T1 biped: sole position (B,2,3) and contact Jacobian (B,2,6,6) from the joint angles (B,2,6); analytical IK (srbd_t1.cu)
cudaError_t launch_t1_kinematics(const float* q, float* p, float* J, int B, cudaStream_t stream);
cudaError_t launch_t1_ik(const float* p, float* q, int B, cudaStream_t stream);
// host entry: one chunk of inputs from mapped pinned host memory into device staging, co-resident with a running solve
cudaError_t launch_prefetch(const double* const* src, double* const* dst, const int* n, const float* rot_src, float* rot_dst,
                            int rot_n, int ctas, cudaStream_t stream);
// DFMA micro-benchmark used for the FP64 roofline denominator
cudaError_t launch_dfma_peak(double* sink, int iters, int blocks, int threads, cudaStream_t stream);

}  // namespace srbd
