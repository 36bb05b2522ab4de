// TEST-ONLY: the per-env bodies of the small fp32 kernels around the MPC tick (csrc/srbd_stack_impl.h) compiled for the
// host, looped over envs.  Lets the CPU test-suite run the device code of the control stack against the recordings of the
// reference stack (tests/golden/srbd_ref_loop_v1*.npz, srbd_ref_gait_v1.npz, srbd_ref_controller_v1.npz) without a GPU.
// Built with -ffp-contract=off so that single-precision operations round like the CUDA intrinsics.  Never linked into
// the product library.
#include <cstring>

#include "srbd_stack_impl.h"
#include "srbd_t1_impl.h"

using namespace srbd;

extern "C" {

int srbd_emul_update_state(const srbd_loop_buffers* buf, const float* state, int state_stride, int B) {
  UpdateStateArgs a;
  static_cast<srbd_loop_buffers&>(a) = *buf;
  a.state = state; a.state_stride = state_stride;
  for (long long e = 0; e < B; ++e) update_state_env(a, e);
  return 0;
}

int srbd_emul_lowlevel(const srbd_loop_buffers* buf, const float* kp, const float* kd, const float* torque_limit, double dt,
                       int world_frame, int advance_gait, int B) {
  LowLevelArgs a;
  static_cast<srbd_loop_buffers&>(a) = *buf;
  for (int i = 0; i < 5; ++i) { a.kp[i] = kp[i]; a.kd[i] = kd[i]; }
  for (int i = 0; i < 10; ++i) a.torque_limit[i] = torque_limit[i];
  a.dt = (float)dt; a.world_frame = world_frame; a.advance_gait = advance_gait; a.wrench = buf->feedforward_force;
  for (long long e = 0; e < B; ++e) lowlevel_env(a, e);
  return 0;
}

int srbd_emul_gait(float* phase, const float* dt_mpc, const int* ssp, const int* dsp, double dt, int advance, int horizon,
                   float* contact_sub, float* swing_sub, int* table_i32, float* table_f32, int B) {
  GaitArgs a;
  a.phase = phase; a.dt_mpc = dt_mpc; a.ssp = ssp; a.dsp = dsp; a.dt = (float)dt; a.advance = advance; a.horizon = horizon;
  a.contact_sub = contact_sub; a.swing_sub = swing_sub; a.table_i32 = table_i32; a.table_f32 = table_f32;
  for (long long e = 0; e < B; ++e) gait_env(a, e);
  return 0;
}

int srbd_emul_reference_trajectory(const float* const* f32in9, double step, float* wpd, float* yaw_des, unsigned char* first_run,
                                   float* x0, float* x_ref, int B) {
  RefTrajArgs a;
  a.euler = f32in9[0]; a.pos = f32in9[1]; a.omega_w = f32in9[2]; a.vel_w = f32in9[3]; a.rot = f32in9[4];
  a.v_des_b = f32in9[5]; a.w_des_b = f32in9[6]; a.height = f32in9[7]; a.dt_mpc = f32in9[8];
  a.step = (float)step; a.wpd = wpd; a.yaw_des = yaw_des; a.first_run = first_run; a.x0 = x0; a.x_ref = x_ref;
  for (long long e = 0; e < B; ++e) reference_trajectory_env(a, e);
  return 0;
}

// batched T1 kinematics (csrc/srbd_t1_impl.h), one (env, leg) pair after the other
int srbd_emul_t1_kinematics(const float* q, float* p, float* J, int B) {
  for (long long i = 0; i < 2LL * B; ++i) t1_fk_jac(q + 6 * i, (int)(i & 1), p ? p + 3 * i : nullptr, J ? J + 36 * i : nullptr);
  return 0;
}
int srbd_emul_t1_ik(const float* p, float* q, int B) {
  for (long long i = 0; i < 2LL * B; ++i) t1_ik(p + 3 * i, (int)(i & 1), q + 6 * i);
  return 0;
}

}  // extern "C"
