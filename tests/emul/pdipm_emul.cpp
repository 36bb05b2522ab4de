// TEST-ONLY: runs the PDIPM *device* code (srbd_pdipm_core.h) on the CPU under the
// fiber-based warp emulator, one emulated warp per environment.  Used by the CPU test
// suite to check the kernel's indexing / data flow against the oracle without a GPU.
// Never linked into the product library.
#define SRBD_HOST_EMUL 1
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "srbd_stage_table.h"
#include "srbd_pdipm_fast.h"

namespace srbd {
static std::map<std::string, std::vector<double>> g_trace;
static bool g_trace_on = false;
void emul_trace(const char* tag, const double* ptr, int n) {
  if (!g_trace_on) return;
  WarpEmulState* st = emul_state();
  if (st && st->cur != 0) return;   // lane 0 records
  g_trace[tag].assign(ptr, ptr + n);
}
}  // namespace srbd

using namespace srbd;

template <bool kUniform>
static void run_env(const PdipmArgs& A, long long e, const StageSlot* slots, int* iters) {
  std::vector<double> smem(PdipmSmem<kUniform>::SIZE, 0.0);
  run_warp([&](int lane) {
    PdipmWarp<kUniform> W;
    W.w.lane_ = lane;
    W.sm = smem.data();
    W.slots = slots;
    int it = W.run(A, e);
    if (lane == 0 && iters) *iters = it;
  });
}

static int g_f32_factor = 0;   // PdipmArgs::precision: 1 / 2 = the reduced-precision pivot-block mode (PdipmFastT<true>)
template <bool kF32>
static void run_env_fast_t(const PdipmArgs& A, long long e, const StageSlot* slots) {
  std::vector<double> smem(FastSmemT<!kF32>::SIZE, 0.0);
  std::vector<double> tmem(32 * 60, 0.0);     // stands in for the warp's tensor-memory slots (lane-private rows)
  run_warp([&](int lane) {
    PdipmFastT<kF32> W;
    W.w.lane_ = lane;
    W.sm = smem.data();
    W.slots = slots;
    W.tme = tmem.data();
    W.run(A, e);
  });
}
static void run_env_fast(const PdipmArgs& A, long long e, const StageSlot* slots) {
  if (g_f32_factor) run_env_fast_t<true>(A, e, slots);
  else run_env_fast_t<false>(A, e, slots);
}
extern "C" void srbd_emul_set_f32_factor(int on) { g_f32_factor = on; }

extern "C" int srbd_emul_pdipm(const double* const* in10, double* const* out6, double* work, int* iters_done,
                               int B, int num_iters, double tol, int cold_start, int uniform, int trace) {
  static StageSlot slots[SB_SIZE];
  static StageSlot fslots[FB_NSLOT];
  build_stage_slots(slots);
  build_fast_slots(slots, fslots);
  PdipmArgs A;
  for (int i = 0; i < IP_N_IN; ++i) A.in[i] = in10[i];
  for (int i = 0; i < IP_N_OUT; ++i) A.out[i] = out6[i];
  A.work = work; A.work_stride = PDIPM_WORK_DOUBLES; A.inv_flag = nullptr; A.iters_done = iters_done; A.num_iters = num_iters; A.tol = tol; A.cold_start = cold_start;
  A.precision = g_f32_factor; A.fused = 0; A.rot = nullptr; A.wrench = nullptr; A.warm_xy = 0; A.reset = nullptr;
  for (int i = 0; i < QP_N_IN; ++i) A.qp_in[i] = nullptr;
  g_trace_on = trace != 0;
  g_trace.clear();
  for (long long e = 0; e < B; ++e) {
    if (uniform == 2) run_env_fast(A, e, fslots);
    else if (uniform) run_env<true>(A, e, slots, nullptr);
    else run_env<false>(A, e, slots, nullptr);
  }
  return 0;
}

// fused tick (assembly from the 17 qp_former inputs + PDIPM + wrench) under the emulator
extern "C" int srbd_emul_tick(const double* const* in17, const float* rot, double* const* out6, double* work,
                              float* wrench, int* iters_done, int B, int num_iters, double tol, int pair,
                              int warm, const int* reset) {
  static StageSlot slots[SB_SIZE];
  static StageSlot fslots[FB_NSLOT];
  build_stage_slots(slots);
  build_fast_slots(slots, fslots);
  PdipmArgs A;
  for (int i = 0; i < IP_N_IN; ++i) A.in[i] = nullptr;
  for (int i = 0; i < QP_N_IN; ++i) A.qp_in[i] = in17[i];
  for (int i = 0; i < IP_N_OUT; ++i) A.out[i] = out6[i];
  A.work = work; A.work_stride = PDIPM_WORK_DOUBLES; A.inv_flag = nullptr; A.iters_done = iters_done; A.num_iters = num_iters; A.tol = tol; A.cold_start = 1;
  A.precision = g_f32_factor; A.fused = 1; A.rot = rot; A.wrench = wrench; A.warm_xy = warm; A.reset = reset;
  g_trace_on = false;
  (void)pair;
  for (long long e = 0; e < B; ++e) run_env_fast(A, e, fslots);
  return 0;
}

// lane scheduling order of the emulated warp between barriers: +1 ascending, -1 descending
extern "C" void srbd_emul_set_order(int order) { emul_order() = order < 0 ? -1 : 1; }

extern "C" int srbd_emul_trace_get(const char* tag, double* out, int n) {
  auto it = g_trace.find(tag);
  if (it == g_trace.end()) return -1;
  int m = (int)it->second.size() < n ? (int)it->second.size() : n;
  std::memcpy(out, it->second.data(), sizeof(double) * m);
  return m;
}
