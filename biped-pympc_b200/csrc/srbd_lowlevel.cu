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
#include <math.h>

#include "srbd_kernels.h"

namespace srbd {
namespace {

struct V3 { float x, y, z; };
struct M3 { float m[9]; };   // row-major

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 mul(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
            R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z);
}
__device__ __forceinline__ V3 mulT(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
            R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z);
}
__device__ __forceinline__ M3 mul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
__device__ __forceinline__ M3 rot_z(float a) {
  float s, c;
  sincosf(a, &s, &c);
  M3 R = {{c, -s, 0.f, s, c, 0.f, 0.f, 0.f, 1.f}};
  return R;
}
__device__ __forceinline__ V3 col2(const M3& R) { return v3(R.m[2], R.m[5], R.m[8]); }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// HECTOR link offsets after the frame permutations of initialize_kinematics (hector.py:52-110):
// p2 = R12p' p2, p3.. = R23p' R12p' p.. ; left leg.  The right leg mirrors y of p1, p2 and z of p3..p5.
struct Frames { M3 R[5]; V3 p[5]; V3 p0e; };   // T01 .. T05 and the foot sole

__device__ void hector_fk(const float* q, int leg, Frames& F) {
  const float sy = leg ? -1.f : 1.f;
  const V3 p1 = v3(-0.00f, sy * 0.047f, -0.1265f);
  const V3 p2 = v3(0.0705f, sy * 0.015f, 0.0465f);
  const V3 p3 = v3(-0.06f, 0.f, sy * 0.018f);
  const V3 p4 = v3(0.f, 0.22f, sy * 0.01805f);
  const V3 p5 = v3(0.f, 0.22f, sy * 0.f);
  const V3 p5e = v3(0.f, 0.042f, 0.f);
  const M3 R12p = {{0.f, 0.f, 1.f, 0.f, 1.f, 0.f, -1.f, 0.f, 0.f}};
  const M3 R23p = {{0.f, 1.f, 0.f, 0.f, 0.f, 1.f, 1.f, 0.f, 0.f}};
  F.R[0] = rot_z(q[0]); F.p[0] = p1;                                        // T01
  const M3 A = mul(F.R[0], R12p);
  F.R[1] = mul(A, rot_z(q[1])); F.p[1] = F.p[0] + mul(A, p2);               // T02 = T01 T12p T2p2
  const M3 Bm = mul(F.R[1], R23p);
  F.R[2] = mul(Bm, rot_z(q[2])); F.p[2] = F.p[1] + mul(Bm, p3);             // T03 = T02 T23p T3p3
  F.R[3] = mul(F.R[2], rot_z(q[3])); F.p[3] = F.p[2] + mul(F.R[2], p4);     // T04 = T03 T34
  F.R[4] = mul(F.R[3], rot_z(q[4])); F.p[4] = F.p[3] + mul(F.R[3], p5);     // T05 = T04 T45
  F.p0e = F.p[4] + mul(F.R[4], p5e);
}

// sub-phases of the gait at phase ph (same arithmetic as gait_kernel in srbd_misc.cu)
__device__ void gait_sub_phases(float ph, int s0, int s1, int d0, int d1, float* contact, float* swing) {
  const float Lf = (float)((d0 + s0) + (d1 + s1));
  const float sp0 = __fdiv_rn((float)s0, Lf), sp1 = __fdiv_rn((float)s1, Lf);
  const float dp0 = __fdiv_rn((float)d0, Lf), dp1 = __fdiv_rn((float)d1, Lf);
  float cl = -1.f, cr = -1.f, wl = -1.f, wr = -1.f;
  const float th1 = __fadd_rn(sp0, dp0), th2 = __fadd_rn(th1, sp1);
  if (ph < th1) cl = __fdiv_rn(ph, th1);
  if (ph >= th2) cl = __fdiv_rn(__fsub_rn(ph, th2), dp0);
  if (!(ph < sp1)) cr = __fdiv_rn(__fsub_rn(ph, sp1), __fadd_rn(__fadd_rn(dp0, sp1), dp1));
  const float st = __fadd_rn(sp1, dp0), en = __fadd_rn(st, sp0);
  if (ph >= st && ph < en) wl = __fdiv_rn(__fsub_rn(ph, st), sp0);
  if (ph < sp1) wr = __fdiv_rn(ph, sp1);
  contact[0] = cl; contact[1] = cr; swing[0] = wl; swing[1] = wr;
}

}  // namespace

__global__ void update_state_kernel(UpdateStateArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= B) return;
  // all loads first (the stores below may alias them as far as the compiler knows): one round trip to L2
  float st[33];
  {
    const float* g = a.state + e * a.state_stride;
#pragma unroll
    for (int i = 0; i < 33; ++i) st[i] = g[i];
  }
  const float phase = a.gait_phase[e];
  const int s0 = a.ssp_durations[2 * e], s1 = a.ssp_durations[2 * e + 1], d0 = a.dsp_durations[2 * e], d1 = a.dsp_durations[2 * e + 1];
  // ---- gait sub-phases, leg_controller.update_gait_data ------------------------------------
  float cph[2], sph[2];
  gait_sub_phases(phase, s0, s1, d0, d1, cph, sph);
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    a.contact_phase[2 * e + l] = cph[l]; a.swing_phase[2 * e + l] = sph[l];
    a.contact_bool[2 * e + l] = cph[l] != -1.f ? 1.f : 0.f;
    a.swing_bool[2 * e + l] = sph[l] != -1.f ? 1.f : 0.f;
  }
  // ---- state estimator ------------------------------------------------------------------------
  const V3 pos = ld3(st);
  const float qw = st[3], qx = st[4], qy = st[5], qz = st[6];
  M3 R;
  {
    const float n = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    const float w = qw / n, x = qx / n, y = qy / n, z = qz / n;
    const float ww = w * w, xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R.m[0] = ww + xx - yy - zz; R.m[1] = 2.f * (xy - wz); R.m[2] = 2.f * (xz + wy);
    R.m[3] = 2.f * (xy + wz); R.m[4] = ww - xx + yy - zz; R.m[5] = 2.f * (yz - wx);
    R.m[6] = 2.f * (xz - wy); R.m[7] = 2.f * (yz + wx); R.m[8] = ww - xx - yy + zz;
  }
  {   // quat_to_euler on the raw quaternion (math_utils.py:111-151); atan2(sin, cos) re-wraps
    float roll = atan2f(2.f * (qw * qx + qy * qz), 1.f - 2.f * (qx * qx + qy * qy));
    roll = atan2f(sinf(roll), cosf(roll));
    float pitch = asinf(clampf(2.f * (qw * qy - qz * qx), -1.f, 1.f));
    pitch = atan2f(sinf(pitch), cosf(pitch));
    float yaw = atan2f(2.f * (qw * qz + qx * qy), 1.f - 2.f * (qy * qy + qz * qz));
    yaw = atan2f(sinf(yaw), cosf(yaw));
    st3(a.root_euler + 3 * e, v3(roll, pitch, yaw));
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) a.rotation_body[9 * e + i] = R.m[i];
  st3(a.root_position + 3 * e, pos);
  st3(a.root_velocity_b + 3 * e, ld3(st + 7));
  st3(a.root_angular_velocity_b + 3 * e, ld3(st + 10));
  st3(a.root_velocity_w + 3 * e, mul(R, ld3(st + 7)));
  st3(a.root_angular_velocity_w + 3 * e, mul(R, ld3(st + 10)));
  // ---- leg kinematics ---------------------------------------------------------------------------
  Frames FL, FR;
  hector_fk(st + 13, 0, FL);
  hector_fk(st + 18, 1, FR);
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const Frames& F = l ? FR : FL;
    const float* qd = st + 23 + 5 * l;
    float* J = a.J + (e * 2 + l) * 30;          // (6,5) row-major
    V3 v = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const V3 z = col2(FL.R[i]);               // hector.py:209-213: axes of the LEFT transforms for both legs
      const V3 jv = cross(z, F.p0e - F.p[i]);
      J[i] = jv.x; J[5 + i] = jv.y; J[10 + i] = jv.z; J[15 + i] = z.x; J[20 + i] = z.y; J[25 + i] = z.z;
      v = v + qd[i] * jv;
    }
    st3(a.p + (e * 2 + l) * 3, F.p0e);
    st3(a.v + (e * 2 + l) * 3, v);
    st3(a.foot_position + (e * 2 + l) * 3, mul(R, F.p0e) + pos);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      a.q[(e * 2 + l) * 5 + i] = st[13 + 5 * l + i];
      a.qd[(e * 2 + l) * 5 + i] = qd[i];
    }
  }
}

__device__ void hector_ik(V3 p_des, int leg, float* q) {   // hector.py:234-291
  const float side = leg ? 1.f : -1.f;
  V3 f = p_des - v3(-0.00f + 0.0465f - 0.06f, -side * (0.047f + 0.015f), -0.126f - 0.0705f);
  f.z += 0.042f;
  const float d2 = sqrtf(f.y * f.y + f.z * f.z);
  const float dh = 0.018f + 0.01805f;
  const float q1 = asinf(clampf(f.y / d2, -1.f, 1.f)) + asinf(clampf(dh * side / d2, -1.f, 1.f));
  float s, c;
  sincosf(q1, &s, &c);
  const V3 g = v3(f.x, c * f.y - s * f.z + 0.018f * side, s * f.y + c * f.z);    // rot_x(q1) f + offset
  const float r = sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
  const float lt = 0.22f, lc = 0.22f;
  const float c2 = clampf((r * r - lt * lt - lc * lc) / (2.0f * lt * lc), -1.f, 1.f);
  const float s2 = clampf(-sqrtf(fmaxf(1.0f - c2 * c2, 1e-6f)), -1.f, 1.f);
  q[0] = 0.f;
  q[1] = q1;
  q[3] = atan2f(s2, c2);
  q[2] = atan2f(-g.x, -g.z) - atan2f(lc * s2, lt + lc * c2);
  q[4] = -q[2] - q[3];
}

__global__ void lowlevel_kernel(LowLevelArgs a, int B) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= B) return;
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = a.rotation_body[9 * e + i];
  const V3 pos = ld3(a.root_position + 3 * e), vw = ld3(a.root_velocity_w + 3 * e);
  const V3 vdw = mul(R, ld3(a.desired_velocity_b + 3 * e));
  const float dtm = a.dt_mpc[e];
  const float height = a.foot_height[e], cp1 = a.cp1[e], cp2 = a.cp2[e];
  const float fbx = clampf(0.03f * (vw.x - vdw.x), -0.3f, 0.3f), fby = clampf(0.03f * (vw.y - vdw.y), -0.3f, 0.3f);
  float action[10];
  // all loads of both legs first (the stores below may alias them as far as the compiler knows)
  struct LegIn { float cph, sph, rem, contact, p0[3], src[3], J[30], ff[6], q[5], qd[5]; int ssp; bool first; } in[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const long long el = e * 2 + l;
    LegIn& L = in[l];
    L.cph = a.contact_phase[el]; L.sph = a.swing_phase[el]; L.rem = a.swing_time_remaining[el]; L.contact = a.contact_bool[el];
    L.ssp = a.ssp_durations[el]; L.first = a.first_swing[el] != 0;
    const float* src = (a.world_frame ? a.foot_position : a.p) + el * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) { L.p0[i] = a.p0[el * 3 + i]; L.src[i] = src[i]; }
#pragma unroll
    for (int i = 0; i < 30; ++i) L.J[i] = a.J[el * 30 + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.ff[i] = a.wrench[el * 6 + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) { L.q[i] = a.q[el * 5 + i]; L.qd[i] = a.qd[el * 5 + i]; }
  }
  const float phase = a.gait_phase[e];
  const int gs0 = a.ssp_durations[2 * e], gs1 = a.ssp_durations[2 * e + 1], gd0 = a.dsp_durations[2 * e], gd1 = a.dsp_durations[2 * e + 1];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const long long el = e * 2 + l;
    const LegIn& L = in[l];
    const float cph = L.cph, sph = L.sph;
    const float dur = (float)L.ssp * dtm;                          // gait.swing_durations_sec
    // ---- update_swing_time (:146-157) -----------------------------------------------------------
    bool first = L.first;
    float rem = first ? dur : L.rem - a.dt;
    if (cph >= 0.f) first = true;
    // ---- compute_foot_placement (:159-216) --------------------------------------------------------
    const V3 hip = v3((-0.00f + 0.0465f) - 0.06f, (l ? -1.f : 1.f) * ((0.047f + 0.015f) + 0.036f), 0.f);
    V3 fp = (pos + mul(R, hip)) + rem * (0.5f * vw);
    fp.x += fbx; fp.y += fby; fp.z = 0.f;
    const V3 fpb = mulT(R, fp - pos);
    st3(a.foot_placement + el * 3, fp);
    st3(a.foot_placement_b + el * 3, fpb);
    const V3 pf = a.world_frame ? fp : fpb;
    // ---- compute_foot_desired_position (:218-263) -----------------------------------------------
    V3 p0 = ld3(L.p0);
    if (first && sph >= 0.f) p0 = ld3(L.src);
    if (sph >= 0.f) first = false;
    if (cph >= 0.f) first = true;
    st3(a.p0 + el * 3, p0);
    a.first_swing[el] = first ? 1 : 0;
    a.swing_time_remaining[el] = rem;
    V3 pd, vd;
    {   // cubic Bezier (swing_leg_trajectory.py:97-135), also for phase = -1
      V3 c1 = p0 + cp1 * (pf - p0), c2 = p0 + cp2 * (pf - p0);
      const float za = p0.z + height;
      c1.z = (8.f * za - p0.z - pf.z) / 6.0f; c2.z = c1.z;
      const float u = 1.f - sph, u2 = u * u, s2 = sph * sph;
      pd = ((u2 * u) * p0 + (3.f * u2 * sph) * c1) + (3.f * u * s2) * c2 + (s2 * sph) * pf;
      const V3 num = ((3.f * u2) * (c1 - p0) + (6.f * u * sph) * (c2 - c1)) + (3.f * s2) * (pf - c2);
      vd = v3(num.x / dur, num.y / dur, num.z / dur);      // dur = 0 for the standing gait: inf / nan like the reference
      if (a.world_frame) { pd = mulT(R, pd - pos); vd = mulT(R, vd - vw); }
    }
    st3(a.pDes + el * 3, pd);
    st3(a.vDes + el * 3, vd);
    // ---- leg_controller.update_command (:72-119) ---------------------------------------------------
    const bool contact = L.contact != 0.f;
    const float* J = L.J;
    const float* ff = L.ff;
    float qdes[5];
    hector_ik(pd, l, qdes);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      float tau = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) tau += J[5 * k + i] * ff[k];                 // J' f
      float qdd = J[i] * vd.x + J[5 + i] * vd.y + J[10 + i] * vd.z;          // Jv' vDes
      if (i == 0 || i == 4) qdd = 0.f;
      const float tq = contact ? tau : 0.f, qD = contact ? 0.f : qdes[i], qdD = contact ? 0.f : qdd;
      const float kp = contact ? 0.f : a.kp[i], kd = a.kd[i];
      a.tau[el * 5 + i] = tq; a.qDes[el * 5 + i] = qD; a.qdDes[el * 5 + i] = qdD;
      a.kpjoint[el * 5 + i] = kp; a.kdjoint[el * 5 + i] = kd;
      const float t = tq + (kp * (qD - L.q[i]) + kd * (qdD - L.qd[i]));
      const float lim = a.torque_limit[5 * l + i];
      action[5 * l + i] = clampf(t, -lim, lim);
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) a.action[e * 10 + i] = action[i];
  if (a.advance_gait) {   // gait_generator.update_phase, same arithmetic as gait_kernel
    const float Lf = (float)((gd0 + gs0) + (gd1 + gs1));
    const float ph = phase;
    float np = __fadd_rn(ph, __fmul_rn(__frcp_rn(__fmul_rn(Lf, dtm)), a.dt));
    np = __fsub_rn(np, np > 1.f ? 1.f : 0.f);
    a.gait_phase[e] = np;
  }
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
