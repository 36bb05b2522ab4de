// The two kernels around the MPC tick of one control step: everything
// BipedController does per step besides the solve, for the HECTOR biped (5 joints per leg).
//
//   update_state_kernel   = MPCController.update_state (mpc_wrapper.py:22-33):
//        GaitGenerator.get_swing_sub_phase / get_contact_sub_phase (gait_generator.py:99-171),
//        LegController.update_gait_data / update_data (leg_controller.py:28-70) with HECTOR's
//        forward kinematics and contact Jacobian (core/robot/hector.py:112-232),
//        StateEstimator.set_body_state / update_foot_position (state_estimator.py:22-47,
//        utils/math_utils.py:53-151)
//   lowlevel_kernel       = BipedController.run_lowlevel (biped_controller.py:120-127, 144-166):
//        SwingLegController.update_swing_time / compute_foot_placement /
//        compute_foot_desired_position (swing_leg_controller.py:146-263, Raibert heuristic :9-62, cubic
//        Bezier swing_leg_trajectory.py:97-135), LegController.update_command (leg_controller.py:72-119,
//        analytical IK hector.py:234-291), GaitGenerator.update_phase (gait_generator.py:79-87) and the
//        joint torque of BipedController.command_joint_torque (biped_controller.py:168-176)
//
// The reference runs these as ~450 small torch kernels per control step, with Python loops over legs and
// boolean-mask indexing (host synchronisations).  Here: one thread per env, both legs, fp32 like the
// reference.  The results agree with a recording of the reference stack to fp32 round-off (sin / cos /
// asin / atan2 and the 4x4 matrix products are not bit-reproducible across libm / BLAS back-ends); the
// integer / boolean state (first_swing, contact flags) is identical.  Reference quirks are kept:
//   * the right leg's Jacobian takes its joint axes from the LEFT leg's transforms (hector.py:209-213);
//   * Euler angles come from the raw quaternion, the rotation matrix from the normalised one;
//   * the Bezier curve is also evaluated for stance legs (swing phase -1) and stored in pDes / vDes.
#include <cuda_runtime.h>

#include "srbd_kernels.h"
#include "srbd_stack_impl.h"

namespace srbd {

__global__ void update_state_kernel(UpdateStateArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < B) update_state_env(a, e);
}
__global__ void lowlevel_kernel(LowLevelArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < B) lowlevel_env(a, e);
}

cudaError_t launch_update_state(const UpdateStateArgs& a, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  update_state_kernel<<<(B + 127) / 128, 128, 0, stream>>>(a, B);
  return cudaGetLastError();
}
cudaError_t launch_lowlevel(const LowLevelArgs& a, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  lowlevel_kernel<<<(B + 127) / 128, 128, 0, stream>>>(a, B);
  return cudaGetLastError();
}

}  // namespace srbd
