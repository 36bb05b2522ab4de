// Developer tool: what a shared-memory load costs on sm_100a for the access patterns of the PDIPM kernel
// (DESIGN.md section 4.2).  One kernel per (width, pattern); every warp issues kLoads loads per loop trip.
// Run plain for ns per warp-level load (14 warps x 148 CTAs, as the PDIPM kernel is launched), and under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum
// for wavefronts per load instruction.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_bcast_bench tools/smem_bcast_bench.cu
#include <cuda_runtime.h>

#include <cstdio>

enum Pattern : int {
  kSame = 0,        // all 32 lanes one address
  kHalves,          // the two half-warps 192 B apart (the sweep's pivot-row broadcast)
  kRows112,         // lane l reads row min(l & 15, 11), rows 112 B apart (a row of Bbar, 14 doubles)
  kRows104,         // the same with rows of 13 doubles (only 8-byte aligned)
  kRows96x6,        // six rows 96 B apart, per half-warp window 576 B apart (the hand-over window)
  kRows96x12,       // twelve rows 96 B apart (an unpacked 12 x 12 block)
  kLanePrivate,     // lane l reads its own 8 / 16 bytes (perfectly coalesced)
  kNumPatterns
};
static const char* kNames[] = {"same", "halves", "rows112", "rows104", "rows96x6", "rows96x12", "lane_private"};

__device__ __forceinline__ unsigned lane_offset(int pattern, int lane, int width) {
  const int h = lane >> 4, r = lane & 15, rr = r < 12 ? r : 11;
  switch (pattern) {
    case kSame: return 0;
    case kHalves: return 192 * h;
    case kRows112: return 112 * rr;
    case kRows104: return 104 * rr;
    case kRows96x6: return 576 * h + 96 * (rr % 6);
    case kRows96x12: return 96 * rr;
    default: return width * lane;
  }
}

template <int kWidth>   // bytes per lane: 8 or 16
__global__ void __launch_bounds__(448, 1) bench_kernel(int pattern, int trips, double* out) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* mine = smem + warp * 2066;   // the PDIPM kernel's per-warp stride
  for (int i = lane; i < 2066; i += 32) mine[i] = (double)(i & 7);
  __syncwarp();
  unsigned addr = (unsigned)__cvta_generic_to_shared(mine) + lane_offset(pattern, lane, kWidth);
  unsigned acc = 0;
#pragma unroll 1
  for (int t = 0; t < trips; ++t) {
    const unsigned at = addr + ((unsigned)(t & 1) << 10);   // same banks, loop-variant (nothing to hoist)
#pragma unroll
    for (int i = 0; i < 6; ++i) {   // twelve doubles, like ld12; every loaded word is used (or the compiler narrows the load)
      if (kWidth == 16) {
        unsigned long long a, b;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(at + 16u * i) : "memory");
        acc ^= (unsigned)a ^ (unsigned)(a >> 32) ^ (unsigned)b ^ (unsigned)(b >> 32);
      } else {
        unsigned long long a, b;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(at + 16u * i) : "memory");
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(b) : "r"(at + 16u * i + 8u) : "memory");
        acc ^= (unsigned)a ^ (unsigned)(a >> 32) ^ (unsigned)b ^ (unsigned)(b >> 32);
      }
    }
  }
  if (acc == 0xdeadbeefu) out[0] = (double)acc;
}

int main() {
  const int smem = 2066 * 14 * 8, trips = 4000;
  cudaFuncSetAttribute(bench_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  double* out;
  cudaMalloc(&out, 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("# %d SMs, 14 warps per SM, %d trips x 12 doubles per lane; nominal clock %d MHz\n", sms, trips, khz / 1000);
  printf("width_bytes,pattern,ms,sm_cycles_per_12_doubles_per_warp_at_nominal_clock\n");
  for (int width = 8; width <= 16; width += 8) {
    for (int p = 0; p < kNumPatterns; ++p) {
      if (width == 16 && p == kRows104) continue;   // not 16-byte aligned
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (width == 8) bench_kernel<8><<<sms, 448, smem>>>(p, trips, out);
        else bench_kernel<16><<<sms, 448, smem>>>(p, trips, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
      }
      // cycles the SM spends per warp-level group of 12 doubles: time * clock / (trips * 14 warps)
      const double cyc = (double)best * 1e-3 * (double)khz * 1e3 / ((double)trips * 14.0);
      printf("%d,%s,%.4f,%.2f\n", width, kNames[p], best, cyc);
    }
  }
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
  return 0;
}
