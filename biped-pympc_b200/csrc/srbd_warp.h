// Warp-level primitives used by the PDIPM device code.
//
// Under nvcc these are the real thing (__syncwarp / __shfl_*).  When the device code is
// compiled by the host compiler with SRBD_HOST_EMUL defined (tests only), a warp is
// emulated by 32 cooperative fibers (ucontext): a lane runs until it reaches a
// warp-synchronous point and then yields to the next lane.  The emulator exists so that
// the indexing / data-flow logic of the kernel can be debugged on a machine without a
// GPU; it is never part of the product path.
#pragma once

#if defined(__CUDACC__) && !defined(SRBD_HOST_EMUL)

#define SRBD_DEV __device__ __forceinline__
#define SRBD_TRACE(tag, ptr, n) ((void)0)

namespace srbd {
struct Warp {
  SRBD_DEV int lane() const { return (int)(threadIdx.x & 31u); }
  SRBD_DEV void sync() const { __syncwarp(); }
  SRBD_DEV double shfl_xor(double v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
  SRBD_DEV double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  SRBD_DEV int shfl_i(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  SRBD_DEV bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
};
SRBD_DEV double ldg(const double* p) { return __ldg(p); }
}  // namespace srbd

#else  // ---------------------------------------------------------------- host emulation

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#define SRBD_DEV inline
#ifndef SRBD_TRACE
#define SRBD_TRACE(tag, ptr, n) ::srbd::emul_trace(tag, ptr, n)
#endif

namespace srbd {

void emul_trace(const char* tag, const double* ptr, int n);   // provided by the test harness

struct WarpEmulState {
  static constexpr int kLanes = 32;
  ucontext_t main_ctx;
  ucontext_t lane_ctx[kLanes];
  std::vector<char> stacks[kLanes];
  bool done[kLanes];
  int cur = 0;
  int barrier_count = 0;
  long long barrier_gen = 0;
  double pub_d[2][kLanes];
  int pub_i[2][kLanes];
  int pub_sel = 0;
  std::function<void(int)> body;
};

inline WarpEmulState*& emul_state() {
  static thread_local WarpEmulState* s = nullptr;
  return s;
}
inline int& emul_order() {
  static int order = 1;
  return order;
}

struct Warp {
  int lane_;
  int lane() const { return lane_; }
  // yield to the next unfinished lane (round robin; emul_order() = -1 walks the lanes in the
  // opposite direction: a kernel without intra-interval races gives bit-identical results)
  static void yield() {
    WarpEmulState* st = emul_state();
    int me = st->cur;
    int nxt = me;
    const int dir = emul_order() < 0 ? WarpEmulState::kLanes - 1 : 1;
    for (int i = 1; i <= WarpEmulState::kLanes; ++i) {
      int c = (me + i * dir) % WarpEmulState::kLanes;
      if (!st->done[c]) { nxt = c; break; }
    }
    if (nxt == me) return;
    st->cur = nxt;
    swapcontext(&st->lane_ctx[me], &st->lane_ctx[nxt]);
  }
  void sync() const {
    WarpEmulState* st = emul_state();
    long long gen = st->barrier_gen;
    if (++st->barrier_count == WarpEmulState::kLanes) {
      st->barrier_count = 0;
      st->barrier_gen++;
    } else {
      while (st->barrier_gen == gen) yield();
    }
  }
  double shfl(double v, int src) const {
    WarpEmulState* st = emul_state();
    int sel = st->pub_sel;
    st->pub_d[sel][lane_] = v;
    sync();
    double r = st->pub_d[sel][src & 31];
    if (lane_ == 0) st->pub_sel ^= 1;
    sync();
    return r;
  }
  double shfl_xor(double v, int m) const { return shfl(v, lane_ ^ m); }
  int shfl_i(int v, int src) const {
    WarpEmulState* st = emul_state();
    int sel = st->pub_sel;
    st->pub_i[sel][lane_] = v;
    sync();
    int r = st->pub_i[sel][src & 31];
    if (lane_ == 0) st->pub_sel ^= 1;
    sync();
    return r;
  }
  bool any(bool p) const {
    int acc = 0;
    for (int src = 0; src < WarpEmulState::kLanes; ++src) acc |= shfl_i(p ? 1 : 0, src);
    return acc != 0;
  }
};
inline double ldg(const double* p) { return *p; }

inline void emul_lane_entry(int lane) {
  WarpEmulState* st = emul_state();
  st->body(lane);
  st->done[lane] = true;
  // hand over to another unfinished lane, or back to main
  const int dir = emul_order() < 0 ? WarpEmulState::kLanes - 1 : 1;
  for (int i = 1; i <= WarpEmulState::kLanes; ++i) {
    int c = (lane + i * dir) % WarpEmulState::kLanes;
    if (!st->done[c]) {
      st->cur = c;
      setcontext(&st->lane_ctx[c]);
    }
  }
  setcontext(&st->main_ctx);
}

// Run `body(lane)` for the 32 lanes of one emulated warp.
inline void run_warp(const std::function<void(int)>& body, size_t stack_bytes = 1 << 20) {
  WarpEmulState st;
  st.body = body;
  emul_state() = &st;
  for (int l = 0; l < WarpEmulState::kLanes; ++l) {
    st.done[l] = false;
    st.stacks[l].resize(stack_bytes);
    getcontext(&st.lane_ctx[l]);
    st.lane_ctx[l].uc_stack.ss_sp = st.stacks[l].data();
    st.lane_ctx[l].uc_stack.ss_size = stack_bytes;
    st.lane_ctx[l].uc_link = &st.main_ctx;
    makecontext(&st.lane_ctx[l], (void (*)())emul_lane_entry, 1, l);
  }
  st.cur = emul_order() < 0 ? WarpEmulState::kLanes - 1 : 0;
  swapcontext(&st.main_ctx, &st.lane_ctx[st.cur]);
  emul_state() = nullptr;
}

}  // namespace srbd
#endif
