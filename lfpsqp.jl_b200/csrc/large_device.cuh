// large_device.cuh -- device helper functions shared by the large-n kernels and the communicator kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "large_ctrl.h"

namespace lfpsqp {

// NaN-propagating max (Julia's norm(x, Inf) returns NaN if any entry is NaN; fmax would drop it and a diverged
// retraction would look converged)
__device__ __forceinline__ double nanmax(double a, double b) { return (b > a || isnan(b)) ? b : a; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { double w = __shfl_xor_sync(0xffffffffu, v, o); v = (w > v || isnan(w)) ? w : v; }
  return v;
}
// sum over the CTA (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double *sh /* >= 33 doubles */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = (lane < nw) ? sh[lane] : 0.0;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_max(double v, double *sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = (lane < nw) ? sh[lane] : 0.0;
  r = warp_max(r);
  return r;
}
// fixed-order sum of np partials by the whole CTA
__device__ __forceinline__ double reduce_partials(const double *part, int np, double *sh) {
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) s += part[i];
  return block_sum(s, sh);
}
__device__ __forceinline__ double reduce_partials_max(const double *part, int np, double *sh) {
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) { double w = part[i]; s = (w > s || isnan(w)) ? w : s; }
  return block_max(s, sh);
}

// Streaming 128-bit load for data that is read once per pass (J, Q, A: GiB-sized): not kept in L1 and marked
// evict-first in L2, so that the m x m factors and the n-vectors, which ARE re-read every iteration, stay resident in
// the 126 MB L2 while the stream passes through.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_stream2(const double *p) {
  double2 v;
  const unsigned long long pol = l2_evict_first_policy();
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;\n" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}

}  // namespace lfpsqp
