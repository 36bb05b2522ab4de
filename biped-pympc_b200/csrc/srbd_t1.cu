// Batched T1 kinematics: forward kinematics + contact Jacobian of both feet, and the analytical
// inverse kinematics, for every env in one launch each -- one thread per (env, leg).  The reference evaluates the same
// quantities through serialized CasADi functions on the CPU for env 0 only (core/robot/t1.py:96-138); see
// srbd_t1_impl.h for the chain and the arithmetic.
#include <cuda_runtime.h>

#include "srbd_kernels.h"
#include "srbd_t1_impl.h"

namespace srbd {

__global__ void t1_kinematics_kernel(const float* __restrict__ q, float* __restrict__ p, float* __restrict__ J, int n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (env, leg) pair
  if (i >= n) return;
  t1_fk_jac(q + 6 * i, (int)(i & 1), p ? p + 3 * i : nullptr, J ? J + 36 * i : nullptr);
}
__global__ void t1_ik_kernel(const float* __restrict__ p, float* __restrict__ q, int n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  t1_ik(p + 3 * i, (int)(i & 1), q + 6 * i);
}

cudaError_t launch_t1_kinematics(const float* q, float* p, float* J, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int n = 2 * B;
  t1_kinematics_kernel<<<(n + 127) / 128, 128, 0, stream>>>(q, p, J, n);
  return cudaGetLastError();
}
cudaError_t launch_t1_ik(const float* p, float* q, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int n = 2 * B;
  t1_ik_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p, q, n);
  return cudaGetLastError();
}

}  // namespace srbd
