/* srbd_mpc.h -- C ABI of the B200-native batched SRBD-MPC solve.
 *
 * This library replaces, on the hot path of rl-augmented-mpc/Biped-PyMPC, the two
 * CusADi-generated shared objects
 *
 *     libqp_former.so                          ::evaluate
 *     libsparse_pdipm_multiple_iterations.so   ::evaluate
 *
 * that biped_pympc/cusadi/src/CusadiFunction.py:28-46 loads with ctypes and that
 * biped_pympc/convex_mpc/mpc_controller_cusadi.py:97-169 calls once + four times per MPC
 * tick.  Every entry point takes plain pointers and sizes (no torch types).
 *
 * Tensor convention (identical to the reference, CusadiFunction.py:71-76):
 *   argument i is a contiguous fp64 device array of shape (B, nnz_i), env-major, element
 *   order = CasADi nnz order (column-major for dense matrices; CCS order for H, A, G).
 *
 *   qp_former (srbd_constraints.py:77-78)
 *     in  (17): x0[12] x[120] u[120] x_ref[120] dt[1] m[1] mu[1] R_body[9] I_world[9]
 *               body_pos[3] left_foot_pos[3] right_foot_pos[3] contact_table[20] Q[12] R[12]
 *               residual_lin_accel[3] residual_ang_accel[3]
 *     out  (6): H[240] f[240] A[1196] b[140] G[280] d[160]
 *   sparse_pdipm_multiple_iterations (sparse_pdipm_solver.py:532-534)
 *     in  (10): Q_val[240] G_val[280] A_val[1196] f[240] h[160] b[140] x[240] s[160] z[160] y[140]
 *     out  (6): x[240] s[160] z[160] y[140] residuals[4] mu[1]
 *
 * All functions return 0 on success or a negative SRBD_E* code; none of them calls exit().
 * Kernels are launched asynchronously on the caller's stream (a cudaStream_t passed as
 * void*; NULL = legacy default stream) unless stated otherwise.
 *
 * Threading.  The native entry points (srbd_qp_form, srbd_pdipm, srbd_marshal_inputs, srbd_wrench,
 * srbd_gait, srbd_reference_trajectory, srbd_update_state, srbd_lowlevel, srbd_t1_kinematics) keep no
 * state between calls beyond per-device one-time configuration, which is guarded (mutex / atomics):
 * they may be called concurrently from several host threads and on several devices given distinct
 * buffers.  A srbd_ctx owns buffers and two internal streams and serves ONE caller (one stream) at a
 * time.  The synchronous entries -- srbd_pdipm_is_stage_invariant and the two CusADi-compatible
 * `evaluate` functions -- share a per-device scratch area behind a per-device mutex: concurrent
 * callers on one device are serialised, callers on different devices are not.
 *
 * Alignment.  srbd_qp_form moves x, u, x_ref and its six outputs as 16-byte vectors: those pointers
 * must be 16-byte aligned (any cudaMalloc / torch allocation is; a view at an odd double offset is
 * not) -- SRBD_EINVAL otherwise.  Everything else only needs natural (8-byte) alignment.
 */
#ifndef SRBD_MPC_H_
#define SRBD_MPC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRBD_OK 0
#define SRBD_EINVAL (-1)    /* null pointer / bad size                                  */
#define SRBD_ECUDA (-2)     /* a CUDA runtime call failed: see srbd_last_cuda_error()   */
#define SRBD_ENOMEM (-3)    /* context workspace allocation failed                      */
#define SRBD_ENODEV (-4)    /* no CUDA device / not an sm_100 device                    */

/* flags for srbd_pdipm / srbd_mpc_solve */
#define SRBD_COLD_START 1u        /* start from x=0, s=max(h,1), z=1, y=1
                                     (mpc_controller_cusadi.py:138-141); in[x,s,z,y] ignored  */
#define SRBD_STAGE_INVARIANT 2u   /* caller guarantees Q/G/A values are identical for every
                                     horizon stage, with unit coefficients on x_{k+1} and on the
                                     Mx rows (true for anything qp_former produced); selects the
                                     fast kernel.  Non-unit coefficients are detected per env and
                                     reported as NaN iterates / iters_done = -1               */

/* OPTIONAL reduced-precision mode of the fast kernel (north_star: "an optional fp32 mode with a stated tolerance of <= 1e-4";
 * the reference has no such mode).  Requires SRBD_STAGE_INVARIANT; SRBD_EINVAL otherwise.
 *   SRBD_F32_FACTOR     the ten 12x12 Schur pivot blocks of a Newton step are formed, coupled, inverted and stored in fp32;
 *                       residuals, right-hand sides, the block recurrences and the iterates stay fp64.  Alone it is NOT within
 *                       1e-4: iterates within 6e-2 of the fp64 path after the reference's 5 iterations.
 *   | SRBD_F32_REFINE   plus one step of iterative refinement per solve against the fp64 operator: 3e-4 after 5 iterations
 *   | SRBD_F32_REFINE2  plus two steps: STATED TOLERANCE 1e-4 on x, s, z, y after 5 iterations (measured 1.3e-6; 2e-4 ... 6e-4
 *                       after 10).  Slower than the fp64 path: each refinement step costs a block-Thomas solve. */
#define SRBD_F32_FACTOR 4u
#define SRBD_F32_REFINE 8u
#define SRBD_F32_REFINE2 16u

#define SRBD_NZ 240
#define SRBD_NEQ 140
#define SRBD_NINEQ 160
#define SRBD_NNZ_H 240
#define SRBD_NNZ_A 1196
#define SRBD_NNZ_G 280
#define SRBD_QP_N_IN 17
#define SRBD_QP_N_OUT 6
#define SRBD_IP_N_IN 10
#define SRBD_IP_N_OUT 6

int srbd_version(void);
const char* srbd_error_string(int code);
/* text of the last CUDA error seen by this thread ("" if none) */
const char* srbd_last_cuda_error(void);

/* Upload the pattern tables to `device` (-1 = current device).  Idempotent. */
int srbd_init(int device);

/* nnz per argument, for binding code that wants to size buffers generically */
int srbd_qp_former_nnz_in(int i);
int srbd_qp_former_nnz_out(int i);
int srbd_pdipm_nnz_in(int i);
int srbd_pdipm_nnz_out(int i);
/* CCS patterns (host arrays, library-owned): colptr has ncol+1 entries */
int srbd_pattern(char which /* 'H','A','G' */, const int** colptr, const int** row, int* nrow, int* ncol, int* nnz);

/* ---- native entry points: HOST arrays of DEVICE pointers ------------------------------ */

/* replaces libqp_former.so::evaluate (generateCUDACode.py:160-183) */
int srbd_qp_form(const double* const* in17, double* const* out6, int B, void* stream);

/* scratch doubles the caller must provide to srbd_pdipm for B environments when it is called with
 * `flags`: 2 048 per env with SRBD_STAGE_INVARIANT (the fast kernel), 6 224 per env without (the general
 * kernel parks its ten prepared stage buffers there). */
size_t srbd_pdipm_work_doubles(int B, unsigned flags);

/* replaces libsparse_pdipm_multiple_iterations.so::evaluate; the iteration count is a
 * runtime argument (the reference bakes num_iters=5, generate_solver_function.py:106).
 *   tol <= 0 : run exactly num_iters Newton iterations
 *   tol  > 0 : stop an environment once |rx|,|rs|,|re| (of that iteration) and the post-step
 *              mu are all < tol (generate_solver_function.py:187), at most num_iters
 *   iters_done : optional device int[B], iterations executed per environment
 * in10[6..9] may alias out6[0..3] (in-place continuation of a previous call). */
int srbd_pdipm(const double* const* in10, double* const* out6, double* work, int* iters_done,
               int B, int num_iters, double tol, unsigned flags, void* stream);

/* SYNCHRONOUS probe: *result = 1 when every env of the batch carries stage-invariant data with
 * unit coefficients on x_{k+1} and on the Mx rows (true for anything qp_former produced), i.e. when
 * SRBD_STAGE_INVARIANT may be passed to srbd_pdipm.  Used by the CusADi-compatible entry, which is
 * synchronous anyway (generateCUDACode.py:166-177), to pick the fast kernel by itself. */
int srbd_pdipm_is_stage_invariant(const double* const* in10, int B, void* stream, int* result);

/* Input builder: the marshalling of mpc_controller_cusadi.py:54-95 as one launch.  f32in12 = HOST array
 * of 12 DEVICE fp32 pointers in this order: x0 (B,12), x_ref (B,10,12), R_body (B,3,3) [= R_body[:,0]],
 * I_world (B,3,3), root_position (B,3), foot_position (B,2,3), contact_table (B,10,2), dt_mpc (B),
 * Q (nQ >= 12), R (12), residual_lin_accel (B,3), residual_ang_accel (B,3).  out17: the 17 fp64 (B,nnz)
 * qp_former inputs; out17[1], out17[2] (x, u) may be NULL.  fix_layout = 0 reproduces the reference's
 * row-major flattening of R_body / I_world / contact_table into buffers read column-major (SURVEY 8b
 * Q1, Q2), 1 passes the intended layout; q_stride_bug = 1 reproduces the 13-long Q strided by 12 (Q3). */
int srbd_marshal_inputs(const float* const* f32in12, int nQ, int fix_layout, int q_stride_bug, double mass, double mu,
                        double* const* out17, int B, void* stream);
/* The same with I_world = R I_body R' (mpc_controller_cusadi.py:60-63: two fp32 batched matmuls per tick in the reference)
 * computed inside the launch: I_body = device fp32 (3,3) row-major shared by all envs, rotation_body = device fp32
 * (B,3,3) row-major (state_estimate_data.rotation_body); f32in12[3] is ignored and may be NULL.  fp32 fused multiply-adds
 * in the order of two chained 3x3 products: within 2 ulp of torch's `R @ I_body @ R.T`.  With this entry the MPC tick
 * is three launches (state builder, input builder, fused solve) and no torch op. */
int srbd_marshal_inputs_iw(const float* const* f32in12, const float* I_body, const float* rotation_body, int nQ, int fix_layout,
                           int q_stride_bug, double mass, double mu, double* const* out17, int B, void* stream);

/* ---- the control step around the MPC tick (HECTOR biped), SURVEY 8f rank 3 ------------------------ *
 * Device buffers of one batch of B envs; all fp32 unless noted, shapes in the comments.  The names are     *
 * the reference's attribute names (core/data/robot_data.py, controllers/*.py).                             */
typedef struct srbd_loop_buffers {
  /* gait + swing state, persistent across steps */
  float* gait_phase;                     /* (B)      GaitGenerator.gait_phase                          in/out */
  const float* dt_mpc;                   /* (B)                                                              */
  const int* ssp_durations;              /* (B,2) int32                                                      */
  const int* dsp_durations;              /* (B,2) int32                                                      */
  float* swing_time_remaining;           /* (B,2)    SwingLegController.swing_time_remaining           in/out */
  unsigned char* first_swing;            /* (B,2) torch.bool                                            in/out */
  float* p0;                             /* (B,2,3)  swing start position                              in/out */
  const float* foot_height;              /* (B)      set_swing_parameters                                    */
  const float* cp1;                      /* (B)                                                              */
  const float* cp2;                      /* (B)                                                              */
  const float* desired_velocity_b;       /* (B,3)    DesiredStateData.desired_velocity_b                     */
  /* state estimate: written by srbd_update_state, read by srbd_lowlevel (and by the MPC tick) */
  float* root_position;                  /* (B,3) */
  float* root_euler;                     /* (B,3) */
  float* rotation_body;                  /* (B,3,3) row-major */
  float* root_velocity_b;                /* (B,3) */
  float* root_angular_velocity_b;        /* (B,3) */
  float* root_velocity_w;                /* (B,3) */
  float* root_angular_velocity_w;        /* (B,3) */
  float* foot_position;                  /* (B,2,3) world frame */
  /* leg data: written by srbd_update_state */
  float* q;                              /* (B,2,5) */
  float* qd;                             /* (B,2,5) */
  float* p;                              /* (B,2,3) foot position, body frame */
  float* v;                              /* (B,2,3) */
  float* J;                              /* (B,2,6,5) contact Jacobian */
  float* contact_phase;                  /* (B,2) -1 = swing */
  float* swing_phase;                    /* (B,2) -1 = stance */
  float* contact_bool;                   /* (B,2) 0 / 1 as float, like LegControllerData.contact_bool */
  float* swing_bool;                     /* (B,2) */
  /* MPC output consumed by srbd_lowlevel */
  const float* feedforward_force;        /* (B,2,6) = the wrench srbd_mpc_solve returned */
  /* leg command: written by srbd_lowlevel */
  float* foot_placement;                 /* (B,2,3) world */
  float* foot_placement_b;               /* (B,2,3) body */
  float* pDes;                           /* (B,2,3) */
  float* vDes;                           /* (B,2,3) */
  float* tau;                            /* (B,2,5) feed-forward torque J' f of the stance legs */
  float* qDes;                           /* (B,2,5) */
  float* qdDes;                          /* (B,2,5) */
  float* kpjoint;                        /* (B,2,5) */
  float* kdjoint;                        /* (B,2,5) */
  float* action;                         /* (B,10)  BipedController.command_joint_torque */
} srbd_loop_buffers;

/* sizeof(srbd_loop_buffers) of this build, for binding code that mirrors the struct (ctypes.Structure) */
size_t srbd_loop_buffers_size(void);
/* MPCController.update_state (mpc_wrapper.py:22-33) as one launch: gait sub-phases + leg_controller.update_gait_data,
 * LegController.update_data with HECTOR's forward kinematics and contact Jacobian (leg_controller.py:28-70,
 * core/robot/hector.py:112-232), StateEstimator.set_body_state / update_foot_position (state_estimator.py:22-47).
 * state: device fp32 (B, state_stride >= 33): [position 3, quaternion wxyz 4, lin vel body 3, ang vel body 3, q 10, qd 10, ...]. */
int srbd_update_state(const srbd_loop_buffers* buf, const float* state, int state_stride, int B, void* stream);
/* BipedController.run_lowlevel + command_joint_torque (biped_controller.py:120-127, 144-176) as one launch: swing
 * time, Raibert foot placement, cubic Bezier swing trajectory (swing_leg_controller.py:146-263,
 * swing_leg_trajectory.py:97-135), J' f feed-forward torque, analytical IK and PD targets (leg_controller.py:72-119,
 * hector.py:234-291), the clamped joint torque, and -- when advance_gait != 0 -- GaitGenerator.update_phase.
 * kp, kd (5), torque_limit (10): HOST arrays.  world_frame: ControllerConf.swing_reference_frame == "world". */
int srbd_lowlevel(const srbd_loop_buffers* buf, const float* kp, const float* kd, const float* torque_limit, double dt,
                  int world_frame, int advance_gait, int B, void* stream);

/* Gait generator step (core/gait/gait_generator.py:79-252) as one launch: from the gait phase on entry the
 * contact / swing sub-phases (get_contact_sub_phase :99-137, get_swing_sub_phase :139-171; -1 = not in that
 * phase) and the MPC contact table (mpc_gait :215-252; int32 as the reference returns it and / or fp32 as
 * BaseMPCController.set_contact_table stores it), then -- when advance != 0 -- update_phase (:79-87).  Any of the
 * four outputs may be NULL.  gait_phase (B) fp32 in / out; dt_mpc (B) fp32; ssp / dsp durations (B,2) int32
 * [left, right] in MPC steps; dt = control time step.  fp32 arithmetic in the reference's order: bit-exact. */
int srbd_gait(float* gait_phase, const float* dt_mpc, const int* ssp_durations, const int* dsp_durations, double dt,
              int advance, int horizon, float* contact_sub_phase, float* swing_sub_phase, int* mpc_table_i32,
              float* contact_table_f32, int B, void* stream);

/* State builder: compute_knot_points + set_initial_state + compute_reference_trajectory of
 * convex_mpc/base_controller.py:164-257 as one launch (fp32, the reference's operation order, no FMA).
 * f32in9 = HOST array of 9 DEVICE fp32 pointers: root_euler (B,3), root_position (B,3),
 * root_angular_velocity_w (B,3), root_velocity_w (B,3), rotation_body (B,3,3 row-major),
 * desired_velocity_b (B,3), desired_angular_velocity_b (B,3), desired_height (B), dt_mpc (B).
 * step = cfg.decimation * cfg.dt.  world_position_desired (B,3), yaw_desired (B) and first_run (B bytes,
 * torch.bool) are the controller's persistent open-loop knot point: read and updated in place
 * (first_run envs are re-anchored at the current pose, :171-176).  Outputs x0 (B,12), x_ref (B,10,12). */
int srbd_reference_trajectory(const float* const* f32in9, double step, float* world_position_desired, float* yaw_desired,
                              unsigned char* first_run, float* x0, float* x_ref, int B, void* stream);

/* ---- batched kinematics of the T1 biped ------------------------------------------------------ *
 * Replaces T1.forward_kinematics / get_p0e / contact_jacobian (core/robot/t1.py:96-138), which evaluate the four
 * CasADi functions casadi/t1_kinematics.py:39-75 generates from model/t1_serial.urdf on the CPU for env 0 only.
 * joint_angle: device fp32 (B,2,6) [left, right] x [Hip_Pitch, Hip_Roll, Hip_Yaw, Knee_Pitch, Ankle_Pitch, Ankle_Roll];
 * foot_position: (B,2,3) sole position in the trunk frame; jacobian: (B,2,6,6) LOCAL_WORLD_ALIGNED frame Jacobian, rows
 * [linear, angular], columns = joints.  Either output may be NULL.  fp64 arithmetic inside, fp32 results (the reference
 * evaluates in double and casts, t1.py:113,137). */
int srbd_t1_kinematics(const float* joint_angle, float* foot_position, float* jacobian, int B, void* stream);
/* T1.analytical_IK (t1.py:139-213): foot_position (B,2,3) trunk frame -> joint_angle (B,2,6), hip yaw = ankle roll = 0. */
int srbd_t1_inverse_kinematics(const float* foot_position, float* joint_angle, int B, void* stream);

/* u0 = x[120:132] -> body-frame wrench (B,2,6) fp32, mpc_controller_cusadi.py:186-203.
 * rotation_body: device fp32 (B,3,3) row-major. */
int srbd_wrench(const double* x_sol, const float* rotation_body, float* wrench, int B, void* stream);

/* ---- fused MPC tick with a library-owned workspace ------------------------------------ */
typedef struct srbd_ctx srbd_ctx;
int srbd_ctx_create(srbd_ctx** ctx, int max_B, int device);
int srbd_ctx_destroy(srbd_ctx* ctx);
/* precision of the fused ticks this context launches (srbd_mpc_solve, srbd_mpc_solve_warm, srbd_mpc_solve_host):
 * SRBD_PRECISION_F64 (default, the parity path), SRBD_PRECISION_F32 = SRBD_F32_FACTOR above,
 * SRBD_PRECISION_F32_REFINED = SRBD_F32_FACTOR | SRBD_F32_REFINE, SRBD_PRECISION_F32_REFINED2 = SRBD_F32_FACTOR |
 * SRBD_F32_REFINE2 (the mode with the stated 1e-4 tolerance). */
#define SRBD_PRECISION_F64 0
#define SRBD_PRECISION_F32 1
#define SRBD_PRECISION_F32_REFINED 2
#define SRBD_PRECISION_F32_REFINED2 3
int srbd_ctx_set_precision(srbd_ctx* ctx, int mode);
/* form -> cold start -> num_calls x iters_per_call Newton iterations -> wrench, device
 * buffers in / out (mpc_controller_cusadi.py:97-203 without the host round trips).
 *   x_out/s_out/z_out/y_out/res_out: optional device buffers (NULL = not wanted)        */
int srbd_mpc_solve(srbd_ctx* ctx, const double* const* in17, const float* rotation_body, int B,
                   int num_calls, int iters_per_call, double tol,
                   float* wrench, double* x_out, double* s_out, double* z_out, double* y_out,
                   double* res_out, int* iters_done, void* stream);
/* EXTENSION (not in the reference, which cold-starts every tick, mpc_controller_cusadi.py:138-141): the same tick
 * warm-started from the previous tick's solution, `iters` Newton iterations in one launch.
 *   mode = SRBD_WARM_XY     (x, y) are kept, s = max(h - G x, 1) and z = 1 are re-initialised as in the cold start;
 *                           20 warm iterations per tick do what 30 cold ones do in tests/closed_loop.py.
 *   mode = SRBD_WARM_SHIFT  (x, y, z) are kept and, when the contact schedule of the tick differs from the previous
 *                           tick's (the horizon rolled on by one stage), shifted by one stage; s = max(h - G x, 1e-2);
 *                           z is floored at 1e-3 and re-centred (z = 0.1 / s) on the rows whose bound the schedule
 *                           changed (contact switches).  5-10 warm iterations per tick do what 30 cold ones do.  The
 *                           previous schedule is remembered in the context's workspace: use one context per batch
 *                           and keep the env order.
 * x_io / y_io (/ z_io): device buffers holding the previous solution on entry and the new one on return, or all NULL =
 * the context's own buffers (whose first call is then a cold start).  reset: optional device int[B]; envs with
 * reset[e] != 0 (episode resets) are cold-started. */
#define SRBD_WARM_XY 1
#define SRBD_WARM_SHIFT 2
int srbd_mpc_solve_warm(srbd_ctx* ctx, const double* const* in17, const float* rotation_body, int B,
                        int iters, double tol, int mode, const int* reset, float* wrench, double* x_io, double* s_out,
                        double* z_io, double* y_io, double* res_out, int* iters_done, void* stream);
/* the same with HOST buffers, synchronous.  The batch is cut into chunks that alternate between two internal streams.
 * Page-locked (pinned) buffers: the first chunk's inputs are pulled into device staging by one prefetch kernel (coalesced
 * reads over PCIe through the UVA mapping), the later chunks' by the copy engine while the previous chunk solves, and the
 * wrench is written to the host buffer by the solve kernel itself; pageable buffers are staged with cudaMemcpyAsync in
 * both directions.  in17_host[1] / [2] (x, u: the linearisation
 * point, dead for this LTV-linear problem) may be NULL = all ones, the constant the reference call
 * site passes (mpc_controller_cusadi.py:55-56); the same holds for in17[1] / [2] of srbd_qp_form
 * and srbd_mpc_solve. */
int srbd_mpc_solve_host(srbd_ctx* ctx, const double* const* in17_host, const float* rotation_body_host,
                        int B, int num_calls, int iters_per_call, double tol,
                        float* wrench_host, double* x_out_host, double* res_out_host);

/* ---- CusADi-compatible entry points (drop-in for CusadiFunction.py:43) ---------------- *
 * Same signature and pointer convention as the generated `evaluate`: `inputs`/`outputs`  *
 * are DEVICE arrays of device pointers, `work` is ignored (its size is not part of the   *
 * signature; the kernels keep their scratch in shared memory and in a per-device buffer  *
 * that only grows when a larger batch arrives -- nothing is allocated per call), the     *
 * call runs on the legacy default stream, is synchronous and returns the kernel time in  *
 * seconds (negative on error).  The PDIPM variant performs the reference's fixed 5       *
 * Newton iterations with the fast kernel, which verifies per env that the data is        *
 * qp_former-shaped; if an env is not, the batch is solved again by the general kernel    *
 * (arbitrary values on the frozen pattern).                                              */
float srbd_qp_former_evaluate(const double** inputs, double* work, double** outputs, int batch_size);
float srbd_sparse_pdipm_multiple_iterations_evaluate(const double** inputs, double* work, double** outputs,
                                                      int batch_size);

/* ---- measurement helper ----------------------------------------------------------------*
 * FP64 FMA micro-benchmark (8 independent chains per thread); returns achieved TFLOP/s    *
 * through *tflops.  Used by bench.py for the PDIPM roofline denominator.                  */
int srbd_dfma_peak(int iters, int blocks, int threads, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SRBD_MPC_H_ */
