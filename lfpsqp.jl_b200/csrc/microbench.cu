// microbench.cu -- FP64 peak microbenchmarks used as roofline denominators for the FP64-issue-bound batched
// kernels and the DMMA Gram (MEASURED_PEAKS.json carries no FP64 figure; SURVEY.md 8d asks for a measured one).
#include "ctx.h"

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// mma.sync m8n8k4 f64: D(8x8) += A(8x4) B(4x8); per warp 512 flops per instruction
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double av = a + threadIdx.x * 1e-6, bv = b;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

}  // namespace

// which: 0 = DFMA (vector FP64), 1 = DMMA (mma.sync.m8n8k4.f64). Returns TFLOP/s in *tflops.
extern "C" int lfpsqp_bench_fp64_peak(lfpsqp_ctx *c, int which, double *tflops) {
  if (!c || !tflops) return LFPSQP_ERR_ARG;
  cudaSetDevice(c->device);
  const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
  double *out = (double *)c->arena(20, (size_t)blocks * threads * 8);
  if (!out) return c->fail(LFPSQP_ERR_NOMEM, "device allocation failed");
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(c->ev0, c->stream);
    if (which == 0) dfma_kernel<<<blocks, threads, 0, c->stream>>>(out, iters, 1.0000001, 1e-9);
    else dmma_kernel<<<blocks, threads, 0, c->stream>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(c->ev1, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return c->cuda_fail(e, "fp64 microbenchmark");
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    double flops = (which == 0) ? (double)blocks * threads * iters * 64 * 2.0
                                : (double)blocks * (threads / 32) * iters * 32 * 512.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return LFPSQP_OK;
}
