// Per-environment bodies of the small fp32 kernels around the MPC tick, shared by the
// CUDA kernels (srbd_misc.cu, srbd_lowlevel.cu: one thread per env) and, for tests only, by a host build
// (tests/emul/stack_emul.cpp) that lets the CPU test-suite run the same code against the recordings of the
// reference stack.  See the .cu files for what each function mirrors (reference file:line).
#pragma once
#include <math.h>

#include "../../include/srbd_mpc.h"

#ifdef __CUDACC__
#define SRBD_STACK_FN __device__ __forceinline__
#else
// host build: IEEE single-precision operations without contraction (compile with -ffp-contract=off)
#define SRBD_STACK_FN static inline
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
#endif

namespace srbd {

// controller-side arguments of the four kernels (device pointers; the public srbd_loop_buffers for the last two)
struct RefTrajArgs {
  const float *euler, *pos, *omega_w, *vel_w, *rot, *v_des_b, *w_des_b, *height, *dt_mpc;
  float step;                       // cfg.decimation * cfg.dt, rounded to fp32 like torch's scalar
  float *wpd, *yaw_des;             // (B,3), (B): in / out
  unsigned char* first_run;         // (B) torch.bool: in / out
  float *x0, *x_ref;                // (B,12), (B,10,12)
};
struct GaitArgs {
  float* phase;                  // (B) in / out
  const float* dt_mpc;           // (B)
  const int *ssp, *dsp;          // (B,2) single / double support durations in MPC steps [left, right]
  float dt;
  int advance, horizon;
  float *contact_sub, *swing_sub;      // (B,2) or null
  int* table_i32;                // (B,horizon,2) or null
  float* table_f32;              // (B,horizon,2) or null
};
struct UpdateStateArgs : srbd_loop_buffers { const float* state; int state_stride; };
struct LowLevelArgs : srbd_loop_buffers { float kp[5], kd[5], torque_limit[10]; float dt; int world_frame, advance_gait; const float* wrench; };


struct V3 { float x, y, z; };
struct M3 { float m[9]; };   // row-major

SRBD_STACK_FN V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SRBD_STACK_FN V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SRBD_STACK_FN V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SRBD_STACK_FN V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
SRBD_STACK_FN V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SRBD_STACK_FN V3 mul(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
            R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z);
}
SRBD_STACK_FN V3 mulT(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
            R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z);
}
SRBD_STACK_FN M3 mul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
SRBD_STACK_FN M3 rot_z(float a) {
  float s, c;
  sincosf(a, &s, &c);
  M3 R = {{c, -s, 0.f, s, c, 0.f, 0.f, 0.f, 1.f}};
  return R;
}
SRBD_STACK_FN V3 col2(const M3& R) { return v3(R.m[2], R.m[5], R.m[8]); }
SRBD_STACK_FN V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
SRBD_STACK_FN void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
SRBD_STACK_FN float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// HECTOR link offsets after the frame permutations of initialize_kinematics (hector.py:52-110):
// p2 = R12p' p2, p3.. = R23p' R12p' p.. ; left leg.  The right leg mirrors y of p1, p2 and z of p3..p5.
struct Frames { M3 R[5]; V3 p[5]; V3 p0e; };   // T01 .. T05 and the foot sole

SRBD_STACK_FN void hector_fk(const float* q, int leg, Frames& F) {
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
SRBD_STACK_FN void gait_sub_phases(float ph, int s0, int s1, int d0, int d1, float* contact, float* swing) {
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


SRBD_STACK_FN void gait_env(const GaitArgs& a, long long e) {
  const int s0 = a.ssp[2 * e], s1 = a.ssp[2 * e + 1], d0 = a.dsp[2 * e], d1 = a.dsp[2 * e + 1];
  const int L = (d0 + s0) + (d1 + s1);                                   // gait_cycle_length, :52
  const float Lf = (float)L, ph = a.phase[e];
  const float sp0 = __fdiv_rn((float)s0, Lf), sp1 = __fdiv_rn((float)s1, Lf);   // :67-68
  const float dp0 = __fdiv_rn((float)d0, Lf), dp1 = __fdiv_rn((float)d1, Lf);
  if (a.contact_sub) {                                                   // :99-137
    float cl = -1.f, cr = -1.f;
    const float th1 = __fadd_rn(sp0, dp0), th2 = __fadd_rn(th1, sp1);
    if (ph < th1) cl = __fdiv_rn(ph, th1);
    if (ph >= th2) cl = __fdiv_rn(__fsub_rn(ph, th2), dp0);
    if (!(ph < sp1)) cr = __fdiv_rn(__fsub_rn(ph, sp1), __fadd_rn(__fadd_rn(dp0, sp1), dp1));
    a.contact_sub[2 * e] = cl; a.contact_sub[2 * e + 1] = cr;
  }
  if (a.swing_sub) {                                                     // :139-171
    float wl = -1.f, wr = -1.f;
    const float st = __fadd_rn(sp1, dp0), en = __fadd_rn(st, sp0);
    if (ph >= st && ph < en) wl = __fdiv_rn(__fsub_rn(ph, st), sp0);
    if (ph < sp1) wr = __fdiv_rn(ph, sp1);
    a.swing_sub[2 * e] = wl; a.swing_sub[2 * e + 1] = wr;
  }
  if (a.table_i32 || a.table_f32) {                                      // :215-252
    const long long step0 = (long long)(int)__fmul_rn(ph, Lf);           // .int() truncates
    for (int k = 0; k < a.horizon; ++k) {
      long long g = (step0 + k) % L;
      if (g < 0) g += L;                                                 // torch's % follows the divisor's sign
      const bool p1 = g < s1, p2 = g >= s1 && g < s1 + d0, p3 = g >= s1 + d0 && g < s1 + d0 + s0;
      const bool rest = !(p1 || p2 || p3);
      const int cl = (p1 || p2 || rest) ? 1 : 0, cr = (p2 || p3 || rest) ? 1 : 0;
      const long long o = ((long long)e * a.horizon + k) * 2;
      if (a.table_i32) { a.table_i32[o] = cl; a.table_i32[o + 1] = cr; }
      if (a.table_f32) { a.table_f32[o] = (float)cl; a.table_f32[o + 1] = (float)cr; }
    }
  }
  if (a.advance) {                                                       // update_phase, :79-87
    const float gds = __fmul_rn(Lf, a.dt_mpc[e]);                        // gait_durations_sec, :77
    float np = __fadd_rn(ph, __fmul_rn(__frcp_rn(gds), a.dt));          // `dt / tensor` is tensor.reciprocal() * dt in torch
    np = __fsub_rn(np, np > 1.f ? 1.f : 0.f);
    a.phase[e] = np;
  }
}


// compute_knot_points + set_initial_state + compute_reference_trajectory for one env (base_controller.py:164-257)
SRBD_STACK_FN void reference_trajectory_env(const RefTrajArgs& a, long long e) {
  const float px = a.pos[e * 3], py = a.pos[e * 3 + 1];
  const float vbx = a.v_des_b[e * 3], vby = a.v_des_b[e * 3 + 1], vbz = a.v_des_b[e * 3 + 2];
  const float wz = a.w_des_b[e * 3 + 2], h = a.height[e], dtm = a.dt_mpc[e];
  float x0[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) { x0[i] = a.euler[e * 3 + i]; x0[3 + i] = a.pos[e * 3 + i]; x0[6 + i] = a.omega_w[e * 3 + i]; x0[9 + i] = a.vel_w[e * 3 + i]; }
  float R[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) R[i] = a.rot[e * 9 + i];
  // knot point (base_controller.py:164-176, 221-225)
  float wx_d = a.wpd[e * 3], wy_d = a.wpd[e * 3 + 1], yaw_d = a.yaw_des[e];
  if (a.first_run[e]) { wx_d = px; wy_d = py; yaw_d = x0[2]; }
  wx_d = __fadd_rn(wx_d, __fmul_rn(a.step, vbx));
  wy_d = __fadd_rn(wy_d, __fmul_rn(a.step, vby));
  yaw_d = __fadd_rn(yaw_d, __fmul_rn(a.step, wz));
  // desired velocity in the world frame (base_controller.py:239)
  const float vwx = __fadd_rn(__fadd_rn(__fmul_rn(R[0], vbx), __fmul_rn(R[1], vby)), __fmul_rn(R[2], vbz));
  const float vwy = __fadd_rn(__fadd_rn(__fmul_rn(R[3], vbx), __fmul_rn(R[4], vby)), __fmul_rn(R[5], vbz));
  const bool stationary = fabsf(vbx) < 1e-2f;     // :227
  for (int k = 0; k < 10; ++k) {
    const float tb = __fmul_rn(dtm, (float)k);                                  // :218
    float* o = a.x_ref + (e * 10 + k) * 12;
    o[0] = 0.f; o[1] = 0.f;
    o[2] = __fadd_rn(yaw_d, __fmul_rn(wz, tb));                                 // :234 (:237-238 is the same expression)
    o[3] = __fadd_rn(stationary ? wx_d : px, __fmul_rn(vwx, tb));               // :240, :245
    o[4] = __fadd_rn(stationary ? wy_d : py, __fmul_rn(vwy, tb));               // :241, :246
    o[5] = h; o[6] = 0.f; o[7] = 0.f; o[8] = wz; o[9] = vwx; o[10] = vwy; o[11] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) a.x0[e * 12 + i] = x0[i];                         // set_initial_state, :204-211
  a.wpd[e * 3] = wx_d; a.wpd[e * 3 + 1] = wy_d; a.wpd[e * 3 + 2] = h;            // :224
  a.yaw_des[e] = yaw_d;
  a.first_run[e] = 0;
}

SRBD_STACK_FN void update_state_env(const UpdateStateArgs& a, long long e) {
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

SRBD_STACK_FN void hector_ik(V3 p_des, int leg, float* q) {   // hector.py:234-291
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

SRBD_STACK_FN void lowlevel_env(const LowLevelArgs& a, long long e) {
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

}  // namespace srbd
