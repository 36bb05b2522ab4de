// QP assembly kernel: the B200-native replacement of the reference's CusADi-generated
// `qp_former` (biped_pympc/casadi/srbd_constraints.py:20-81, emitted as straight-line
// CUDA by biped_pympc/cusadi/src/generateCUDACode.py:62-190).
//
// Same boundary: 17 fp64 inputs (B, nnz_in) -> 6 fp64 CCS value vectors (B, nnz_out),
// env-major, CasADi nnz order (column-major).  HBM-bound: 3 616 B read + 18 048 B
// written per env.  Design:
//   * one CTA = kEnvs envs; warp w builds env w's 228-entry "value table" V in shared
//     memory from closed forms (A_c is nilpotent, so RK4 is exact:
//     A_d = I + dt A_c, B_d = dt B_c + dt^2/2 A_c B_c, c_d = dt c + dt^2/2 A_c c);
//   * the whole CTA then streams the tile's outputs with 16-byte coalesced stores:
//     out[o][env, k] = V[env][tab_o[k]], where tab_o (2 016 bytes, built once on the
//     host from srbd_pattern.h) encodes the frozen CCS patterns.  No scratch `work`
//     array, no pre-zeroed outputs: every output element is written.
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>

#include "srbd_kernels.h"
#include "srbd_pattern.h"
#include "srbd_qp_table.h"

namespace srbd {

constexpr int TAB_H = 0, TAB_A = TAB_H + NNZ_H, TAB_B = TAB_A + NNZ_A, TAB_G = TAB_B + NEQ,
              TAB_D = TAB_G + NNZ_G, TAB_SIZE = TAB_D + NINEQ;   // 2016

__device__ uint8_t g_qp_tab[TAB_SIZE];

void qp_form_build_table(uint8_t* t) {
  for (int j = 0; j < NZ; ++j) t[TAB_H + j] = (uint8_t)((j < 120 ? V_Q : V_R) + j % 12);
  for (int i = 0; i < NNZ_A; ++i) t[TAB_A + i] = 255;
  for (int k = 0; k < N; ++k) {
    for (int c = 0; c < NX; ++c) {
      const int st = a_xcol_start(k, c);
      t[TAB_A + st] = V_ONE;
      if (k < 9) {
        for (int r = 0; r < NX; ++r) {
          const int p = a_xcol_pos(r, c);
          if (p < 0) continue;
          uint8_t v;
          if (r == c) v = V_NEG1;
          else if (r < 3) v = (uint8_t)(V_NADTR + 3 * r + (c - 6));
          else v = V_NADT;
          t[TAB_A + st + p] = v;
        }
      }
    }
    for (int c = 0; c < NU; ++c) {
      const int st = a_ucol_start(k, c);
      for (int r = 0; r < NX; ++r) {
        const int p = a_ucol_pos(r, c);
        if (p >= 0) t[TAB_A + st + p] = (uint8_t)(V_NBD + 12 * r + c);
      }
      if (c == 6 || c == 9) t[TAB_A + st + A_UCOL_MX_POS] = V_ONE;
    }
    for (int r = 0; r < NX; ++r) t[TAB_B + 12 * k + r] = (uint8_t)((k == 0 ? V_B0 : V_CD) + r);
    t[TAB_B + 120 + 2 * k] = V_ZERO;
    t[TAB_B + 121 + 2 * k] = V_ZERO;
    for (int ft = 0; ft < 2; ++ft) {
      const int gb = g_foot_base(k, ft), mb = g_my_base(k, ft);
      // column fx: rows 0,1 ; fy: rows 2,3 ; fz: rows 0..7 ; my: rows 4,5
      t[TAB_G + gb + 0] = V_NEG1;  t[TAB_G + gb + 1] = V_ONE;
      t[TAB_G + gb + 2] = V_NEG1;  t[TAB_G + gb + 3] = V_ONE;
      t[TAB_G + gb + 4] = V_NEGMU; t[TAB_G + gb + 5] = V_NEGMU;
      t[TAB_G + gb + 6] = V_NEGMU; t[TAB_G + gb + 7] = V_NEGMU;
      t[TAB_G + gb + 8] = V_NEGLT; t[TAB_G + gb + 9] = V_NEGLH;
      t[TAB_G + gb + 10] = V_NEG1; t[TAB_G + gb + 11] = V_ONE;
      t[TAB_G + mb + 0] = V_NEG1;  t[TAB_G + mb + 1] = V_ONE;
      for (int j = 0; j < 8; ++j)
        t[TAB_D + 16 * k + 8 * ft + j] = (uint8_t)(j == 7 ? V_FC + 2 * k + ft : V_ZERO);
    }
  }
}

cudaError_t qp_form_upload_table() {
  uint8_t h[TAB_SIZE];
  qp_form_build_table(h);
  for (int i = 0; i < TAB_SIZE; ++i)
    if (h[i] == 255) return cudaErrorUnknown;   // a pattern slot was left unassigned
  return cudaMemcpyToSymbol(g_qp_tab, h, TAB_SIZE);
}

template <int kEnvs>
__global__ void __launch_bounds__(32 * kEnvs)
qp_form_kernel(QpFormArgs a, int B) {
  __shared__ double sV[kEnvs][V_COUNT];
  __shared__ double sRaw[kEnvs][RW_COUNT + 1];
  __shared__ uint8_t sTab[TAB_SIZE];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < TAB_SIZE / 4; i += 32 * kEnvs)
    reinterpret_cast<uint32_t*>(sTab)[i] = reinterpret_cast<const uint32_t*>(g_qp_tab)[i];

  for (long long tile = blockIdx.x; tile * kEnvs < B; tile += gridDim.x) {
    const int env0 = (int)(tile * kEnvs);
    const int nenv = min(kEnvs, B - env0);
    __syncthreads();   // previous tile's readers of sV are done (and sTab is loaded)
    if (warp < nenv) {
      const long long e = env0 + warp;
      double* V = sV[warp];
      double* W = sRaw[warp];
      qp_value_table(a.in, e, lane, 32, Warp(), V, W);
    }
    __syncthreads();

    // ---- stream the tile's outputs (16-byte stores, contiguous across the tile) ----------
    const int nthr = 32 * kEnvs;
    auto emit = [&](double* __restrict__ out, int nnz, int tab) {
      double2* o2 = reinterpret_cast<double2*>(out + (long long)env0 * nnz);
      const int half = nnz >> 1;
      for (int i = tid; i < nenv * half; i += nthr) {
        const int e = i / half, k = (i - e * half) * 2;
        double2 v;
        v.x = sV[e][sTab[tab + k]];
        v.y = sV[e][sTab[tab + k + 1]];
        o2[i] = v;
      }
    };
    emit(a.out[OUT_A], NNZ_A, TAB_A);
    emit(a.out[OUT_G], NNZ_G, TAB_G);
    emit(a.out[OUT_H], NNZ_H, TAB_H);
    emit(a.out[OUT_B], NEQ, TAB_B);
    emit(a.out[OUT_D], NINEQ, TAB_D);
    {   // f = grad - H z   (srbd_constraints.py:67-69), evaluated literally
      // the linearisation point (x, u) only enters through round-off; NULL = all ones, the constant
      // the reference call site passes (mpc_controller_cusadi.py:55-56)
      const double2* x2 = a.in[IN_X] ? reinterpret_cast<const double2*>(a.in[IN_X] + (long long)env0 * 120) : nullptr;
      const double2* u2 = a.in[IN_U] ? reinterpret_cast<const double2*>(a.in[IN_U] + (long long)env0 * 120) : nullptr;
      const double2* r2 = reinterpret_cast<const double2*>(a.in[IN_XREF] + (long long)env0 * 120);
      double2* f2 = reinterpret_cast<double2*>(a.out[OUT_F] + (long long)env0 * NZ);
      const double2 ones = make_double2(1.0, 1.0);
      for (int i = tid; i < nenv * 120; i += nthr) {
        const int e = i / 120, k = (i - e * 120) * 2;   // k in 0..238 (even)
        double2 v;
        if (k < 120) {
          const double2 xv = x2 ? x2[e * 60 + (k >> 1)] : ones, rv = r2[e * 60 + (k >> 1)];
          const double q0 = sV[e][V_Q + k % 12], q1 = sV[e][V_Q + (k + 1) % 12];
          v.x = q0 * (xv.x - rv.x) - q0 * xv.x;
          v.y = q1 * (xv.y - rv.y) - q1 * xv.y;
        } else {
          const double2 uv = u2 ? u2[e * 60 + ((k - 120) >> 1)] : ones;
          const double r0 = sV[e][V_R + k % 12], r1 = sV[e][V_R + (k + 1) % 12];
          const double t0 = r0 * uv.x, t1 = r1 * uv.y;
          v.x = t0 - t0;
          v.y = t1 - t1;
        }
        f2[i] = v;
      }
    }
  }
}

cudaError_t launch_qp_form(const QpFormArgs& args, int B, int num_sms, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  constexpr int kEnvs = 4;
  const int tiles = (B + kEnvs - 1) / kEnvs;
  // Persistent CTAs, exactly as many as are resident at once (48 registers x 128 threads and 12.9 KB of shared
  // memory allow 10 per SM on sm_100a): a larger grid leaves a second, nearly empty wave that runs its tiles
  // at a fifth of the occupancy (ncu with 12 per SM: 1.20 waves, 4.3 TB/s at 65 536 envs; 5.4 TB/s with 10).  One
  // tile per CTA (hardware scheduling) measures the same to 3 %.
  // (a property of the kernel image and the architecture, identical on every sm_100a device; the cached value is an
  //  atomic so that concurrent first calls from several host threads are benign)
  static std::atomic<int> ctas_per_sm{0};
  int cps = ctas_per_sm.load(std::memory_order_relaxed);
  if (cps == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, qp_form_kernel<kEnvs>, 32 * kEnvs, 0) != cudaSuccess || n < 1) n = 8;
    ctas_per_sm.store(n, std::memory_order_relaxed);
    cps = n;
  }
  const int max_ctas = num_sms * cps;
  const int grid = tiles < max_ctas ? tiles : max_ctas;
  qp_form_kernel<kEnvs><<<grid, 32 * kEnvs, 0, stream>>>(args, B);
  return cudaGetLastError();
}

}  // namespace srbd
