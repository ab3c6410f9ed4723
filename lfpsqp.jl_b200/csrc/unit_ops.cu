// unit_ops.cu -- unit-level exports of the bound embedding (src/inequality_helper.jl, y_retract! src/retractions.jl:451-500)
// on ONE instance, so that the reference's own component tests (test/test_inequalities.jl) can be re-run against the
// device code through the C ABI.  Runs the same Solver methods the batched warp kernel uses.
#include <vector>
#include "ctx.h"
#include "batched_warp.cuh"

using namespace lfpsqp;

namespace {
__global__ void __launch_bounds__(32) ineq_op_kernel(int op, int n, int m, const double *bnd_g, const double *Jg, const double *in,
                                                     double *out, lfpsqp_params prm, int *okflag) {
  extern __shared__ double smem[];
  const WarpLayout L(n, m, 0, 1, 0);
  double *bnd = smem;
  for (int i = threadIdx.x; i < 5 * n; i += 32) bnd[i] = bnd_g[i];
  double *ws = smem + ((5 * n + 1) & ~1);
  for (int i = threadIdx.x; i < L.total; i += 32) ws[i] = 0.0;
  __syncwarp();
  WarpGroup g(threadIdx.x);
  Solver<FamBoxQuad, WarpGroup> S(g, L, prm, ws, bnd);
  const int N = 2 * n;
  for (int i = threadIdx.x; i < m * n; i += 32) S.J[i] = Jg ? Jg[i] : 0.0;
  __syncwarp();
  if (op == 0) {                                  // generate_initial_y! (:92-109)
    for (int i = threadIdx.x; i < n; i += 32) S.x[i] = in[i];
    __syncwarp();
    S.generate_initial_y(S.x);
    for (int i = threadIdx.x; i < N; i += 32) out[i] = S.x[i];
  } else if (op == 1) {                           // calculate_h! (:112-122)
    for (int i = threadIdx.x; i < N; i += 32) S.x[i] = in[i];
    __syncwarp();
    S.calculate_h(S.cvaug, S.x);
    for (int i = threadIdx.x; i < n; i += 32) out[i] = S.cvaug[i];
  } else if (op == 2) {                           // inequality_gradient! (:125-141)
    for (int i = threadIdx.x; i < N; i += 32) S.x[i] = in[i];
    __syncwarp();
    S.inequality_gradient(S.x);
    for (int i = threadIdx.x; i < n; i += 32) { out[i] = S.Dx[i]; out[n + i] = S.Dy[i]; out[2 * n + i] = S.S[i]; }
  } else if (op == 3) {                           // y_retract! (retractions.jl:451-500)
    for (int i = threadIdx.x; i < N; i += 32) { S.x[i] = in[i]; S.xnew[i] = in[N + i]; }
    __syncwarp();
    S.y_retract(S.xnew, S.x);
    for (int i = threadIdx.x; i < N; i += 32) out[i] = S.xnew[i];
  } else {
    for (int i = threadIdx.x; i < N; i += 32) S.x[i] = in[i];
    __syncwarp();
    S.inequality_gradient(S.x);
    if (op == 4) {                                // bigA * v  (:215-231)
      for (int i = threadIdx.x; i < n + m; i += 32) S.tm[i] = in[N + i];
      __syncwarp();
      S.fullJ_mulT(S.d, S.tm, 1.0, 0.0);
      for (int i = threadIdx.x; i < N; i += 32) out[i] = S.d[i];
    } else if (op == 5) {                         // bigA' * w (:254-271)
      for (int i = threadIdx.x; i < N; i += 32) S.d[i] = in[N + i];
      __syncwarp();
      S.fullJ_mul(S.tm, S.d);
      for (int i = threadIdx.x; i < n + m; i += 32) out[i] = S.tm[i];
    } else {                                      // d - Q Q' d, lambda, lambda_y (optimize.jl:316-317,:332; :286-308)
      for (int i = threadIdx.x; i < N; i += 32) S.d[i] = in[N + i];
      __syncwarp();
      bool ok = true;
      if (m > 0) S.factor();
      if (threadIdx.x == 0) okflag[0] = ok ? 1 : 0;
      if (ok) {
        S.project(S.d, true);
        for (int i = threadIdx.x; i < N; i += 32) out[i] = S.d[i];
        for (int i = threadIdx.x; i < m; i += 32) out[N + i] = S.lam[i];
        for (int i = threadIdx.x; i < n; i += 32) out[N + m + i] = S.lamy[i];
      }
    }
  }
}
}  // namespace

int build_bounds_public(int64_t n, int64_t p, const double *xl, const double *xu, std::vector<double> &bnd);  // abi.cu

extern "C" int lfpsqp_ineq_op(lfpsqp_ctx *c, int op, int64_t n, int64_t m, const double *xl, const double *xu, const double *J,
                              const double *in, int64_t in_len, double *out, int64_t out_len) {
  if (!c || !xl || !xu || !in || !out || n < 1 || m < 0 || op < 0 || op > 6) return c ? c->fail(LFPSQP_ERR_ARG, "bad arguments") : LFPSQP_ERR_ARG;
  cudaSetDevice(c->device);
  const int64_t N = 2 * n;
  const int64_t need_in[] = {n, N, N, 2 * N, N + n + m, 2 * N, 2 * N}, need_out[] = {N, n, 3 * n, N, N, n + m, N + m + n};
  if (in_len != need_in[op] || out_len != need_out[op]) return c->fail(LFPSQP_ERR_ARG, "lfpsqp_ineq_op: wrong buffer lengths for op %d", op);
  std::vector<double> bnd;
  int ineq = build_bounds_public(n, 0, xl, xu, bnd);
  if (ineq < 0) return c->fail(ineq, "Infeasible: lower bounds cannot be greater than upper bounds");
  if (ineq == 0) {  // all bounds infinite: the embedding is still well defined (every pair on a line)
    bnd.assign(5 * n, 0.0);
  }
  WarpLayout L((int)n, (int)m, 0, 1, 0);
  const size_t smem = (size_t)(((5 * n + 1) & ~1) + L.total) * 8;
  if (smem > (size_t)c->smem_optin) return c->fail(LFPSQP_ERR_NOMEM, "instance too large for the unit-level op");
  cudaFuncSetAttribute(ineq_op_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double *d_bnd = (double *)c->arena(10, bnd.size() * 8), *d_J = (double *)c->arena(11, (size_t)std::max<int64_t>(m * n, 1) * 8),
         *d_in = (double *)c->arena(12, in_len * 8), *d_out = (double *)c->arena(13, out_len * 8 + 8);
  if (!d_bnd || !d_J || !d_in || !d_out) return c->fail(LFPSQP_ERR_NOMEM, "device allocation failed");
  cudaStream_t s = c->stream;
  cudaMemcpyAsync(d_bnd, bnd.data(), bnd.size() * 8, cudaMemcpyHostToDevice, s);
  if (m > 0 && J) cudaMemcpyAsync(d_J, J, (size_t)m * n * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_in, in, in_len * 8, cudaMemcpyHostToDevice, s);
  lfpsqp_params prm; lfpsqp_default_params(&prm);
  int *d_ok = (int *)(d_out + out_len);
  cudaMemsetAsync(d_ok, 0xff, 4, s);
  ineq_op_kernel<<<1, 32, smem, s>>>(op, (int)n, (int)m, d_bnd, (m > 0 && J) ? d_J : nullptr, d_in, d_out, prm, d_ok);
  cudaMemcpyAsync(out, d_out, out_len * 8, cudaMemcpyDeviceToHost, s);
  int ok = 1;
  cudaMemcpyAsync(&ok, d_ok, 4, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return c->cuda_fail(e, "lfpsqp_ineq_op");
  if (op == 6 && ok == 0) return c->fail(LFPSQP_ERR_ARG, "rank-deficient projected Jacobian");
  return LFPSQP_OK;
}
