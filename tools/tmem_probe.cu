// Developer tool: is tensor memory usable as lane-private scratch from an ordinary (non-MMA) kernel?
// 14 warps of one CTA per SM; warp w owns lanes 32 (w % 4) .. +31 (its sub-partition) and columns 128 (w / 4) .. +119:
// five slots of 24 32-bit columns = one row of twelve doubles per lane and slot.  Every warp writes lane- and
// slot-specific doubles with tcgen05.st.32x32b (x16 + x8), reads them back with tcgen05.ld and compares; then a timing
// loop (clock64) of dependent load -> FMA -> store round trips next to the same round trip through shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st24(uint32_t addr, const double v[12]) {
  uint32_t r[24];
#pragma unroll
  for (int i = 0; i < 12; ++i) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(addr + 16), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]) : "memory");
}
__device__ __forceinline__ void tm_ld24(uint32_t addr, double v[12]) {
  uint32_t r[24];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]) : "r"(addr + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(448, 1) probe(int* bad, long long* cyc, int trips) {
  __shared__ uint32_t base_s;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s;
  const uint32_t mine = base + (((uint32_t)(warp & 3) * 32u) << 16) + 128u * (uint32_t)(warp >> 2);
  int nbad = 0;
  // ---- correctness: five slots, lane- and slot-specific values, written in one order and read in another --------
  for (int s = 0; s < 5; ++s) {
    double v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = 1000.0 * blockIdx.x + 100.0 * warp + 10.0 * s + lane + 0.001 * i;
    tm_st24(mine + 24 * s, v);
  }
  tm_wait_st();
  for (int s = 4; s >= 0; --s) {
    double v[12];
    tm_ld24(mine + 24 * s, v);
#pragma unroll
    for (int i = 0; i < 12; ++i) nbad += (v[i] != 1000.0 * blockIdx.x + 100.0 * warp + 10.0 * s + lane + 0.001 * i);
  }
  if (nbad) atomicAdd(bad, nbad);
  // ---- timing: dependent round trips through TMEM and through shared memory (lane-private rows, 96 B apart) ------
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = lane + i;
  __syncthreads();
  long long t0 = clock64();
  for (int t = 0; t < trips; ++t) {
    tm_st24(mine + 24 * (t % 5), acc);
    tm_wait_st();
    double v[12];
    tm_ld24(mine + 24 * (t % 5), v);
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = fma(v[i], 1.0000001, 1e-9);
  }
  long long t1 = clock64();
  double* row = smem + warp * 32 * 12 + lane * 12;
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < 12; i += 2) *reinterpret_cast<double2*>(row + i) = make_double2(acc[i], acc[i + 1]);
    __syncwarp();
    double v[12];
#pragma unroll
    for (int i = 0; i < 12; i += 2) { double2 q = *reinterpret_cast<double2*>(row + i); v[i] = q.x; v[i + 1] = q.y; }
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = fma(v[i], 1.0000001, 1e-9);
    __syncwarp();
  }
  long long t2 = clock64();
  double sum = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) sum += acc[i];
  if (sum == 123.456) bad[1] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(512));
}

int main() {
  int* bad; long long* cyc;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&cyc, 16);
  bad[0] = bad[1] = 0;
  const int trips = 2000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 32 * 12 * 8);
  probe<<<148, 448, 14 * 32 * 12 * 8>>>(bad, cyc, trips);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s, mismatches %d\n", cudaGetErrorString(e), bad[0]);
  printf("round trip (store 12 doubles per lane, wait, load, wait) per trip, 14 warps per SM: TMEM %.1f cycles, shared memory %.1f cycles\n",
         (double)cyc[0] / trips, (double)cyc[1] / trips);
  return e != cudaSuccess || bad[0];
}
