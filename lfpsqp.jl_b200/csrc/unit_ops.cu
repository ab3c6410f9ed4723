// unit_ops.cu -- unit-level exports of the bound embedding (src/inequality_helper.jl, y_retract! src/retractions.jl:451-500)
// on ONE instance, so that the reference's own component tests (test/test_inequalities.jl) can be re-run against the
// device code through the C ABI.  Runs the same Solver methods the batched warp kernel uses.
#include <algorithm>
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


namespace {
// armijo! / exact_linesearch! (src/linesearch.jl:32-89, :107-339) on ONE instance, called as the driver calls them
// (src/optimize.jl:396-420): g = grad f(x), fval = f(x), the factorisation taken at x, retraction chosen by the same rule.
// io: [x (N) | d (N)] in, [xnew (N) | newf, f_diff, step_diff, alpha, tot_iter1, tot_iter2, flag] out.
template <class Fam>
__global__ void __launch_bounds__(32) linesearch_kernel(int which, int n, int m, int ineq, const double *bnd_g, const double *fam_params,
                                                        const double *in, double *out, lfpsqp_params prm) {
  extern __shared__ double smem[];
  const int use_nr = prm.do_project_retract ? 0 : 1;
  const WarpLayout L(n, m, 0, ineq, use_nr, 1);
  double *bnd = smem;
  const int nb = ineq ? 5 * n : 0;
  for (int i = threadIdx.x; i < nb; i += 32) bnd[i] = bnd_g[i];
  double *ws = smem + ((nb + 1) & ~1);
  for (int i = threadIdx.x; i < L.total; i += 32) ws[i] = 0.0;
  __syncwarp();
  WarpGroup g(threadIdx.x);
  Solver<Fam, WarpGroup> S(g, L, prm, ws, bnd);
  S.st = lfpsqp_stats{};
  S.fc.prm = fam_params;
  const int N = L.N;
  for (int i = threadIdx.x; i < N; i += 32) { S.x[i] = in[i]; S.d[i] = in[N + i]; S.gr[i] = 0.0; }
  __syncwarp();
  S.grad_aux(S.gr, S.x);
  const double fval = S.f_aux(S.x);
  if (ineq) S.inequality_gradient(S.x);
  if (m > 0) { S.jac_aux(S.cval, S.x); S.factor(); }
  int kind;
  if (m > 0) kind = (S.rank == m && !prm.do_project_retract) ? 2 : 3; else kind = ineq ? 1 : 0;
  double newf = 0, f_diff = 0, step_diff = 0;
  const int flag = (which == 0) ? S.armijo(kind, fval, &newf, &f_diff, &step_diff)
                                : S.exact_linesearch(kind, fval, &newf, &f_diff, &step_diff);
  for (int i = threadIdx.x; i < N; i += 32) out[i] = S.xnew[i];
  if (threadIdx.x == 0) {
    double *o = out + N;
    o[0] = newf; o[1] = f_diff; o[2] = step_diff; o[3] = S.ls_alpha; o[4] = S.ls_it1; o[5] = S.ls_it2; o[6] = flag;
  }
}

// augmented_hess_lag_vec! (src/inequality_helper.jl:144-158): dest = [H(x, lam) src_x + 2 lam_y q src_x ; 2 lam_y s src_y]
// io: [xaug (2n) | src (2n) | lam (m) | lam_y (n)] in, dest (2n) out
template <class Fam>
__global__ void __launch_bounds__(32) aug_hess_kernel(int n, int m, const double *bnd_g, const double *fam_params, const double *in,
                                                      double *out, lfpsqp_params prm) {
  extern __shared__ double smem[];
  const WarpLayout L(n, m, 0, 1, 0);
  double *bnd = smem;
  for (int i = threadIdx.x; i < 5 * n; i += 32) bnd[i] = bnd_g[i];
  double *ws = smem + ((5 * n + 1) & ~1);
  for (int i = threadIdx.x; i < L.total; i += 32) ws[i] = 0.0;
  __syncwarp();
  WarpGroup g(threadIdx.x);
  Solver<Fam, WarpGroup> S(g, L, prm, ws, bnd);
  S.fc.prm = fam_params;
  const int N = 2 * n;
  for (int i = threadIdx.x; i < N; i += 32) { S.x[i] = in[i]; S.w0[i] = in[N + i]; }
  for (int i = threadIdx.x; i < m; i += 32) S.lam[i] = in[2 * N + i];
  for (int i = threadIdx.x; i < n; i += 32) S.lamy[i] = in[2 * N + m + i];
  __syncwarp();
  S.hess_aux(S.w1, S.w0);
  for (int i = threadIdx.x; i < N; i += 32) out[i] = S.w1[i];
}

template <class K>
int run_unit(lfpsqp_ctx *c, K kern_launch, size_t smem, const std::vector<double> &bnd, const double *fam_params, int64_t npar,
             const double *in, int64_t in_len, double *out, int64_t out_len, const char *what) {
  if (smem > (size_t)c->smem_optin) return c->fail(LFPSQP_ERR_NOMEM, "instance too large for the unit-level op");
  double *d_bnd = (double *)c->arena(10, std::max<size_t>(bnd.size(), 1) * 8), *d_par = (double *)c->arena(11, (size_t)std::max<int64_t>(npar, 1) * 8),
         *d_in = (double *)c->arena(12, in_len * 8), *d_out = (double *)c->arena(13, out_len * 8 + 8);
  if (!d_bnd || !d_par || !d_in || !d_out) return c->fail(LFPSQP_ERR_NOMEM, "device allocation failed");
  cudaStream_t s = c->stream;
  if (!bnd.empty()) cudaMemcpyAsync(d_bnd, bnd.data(), bnd.size() * 8, cudaMemcpyHostToDevice, s);
  if (npar > 0) cudaMemcpyAsync(d_par, fam_params, (size_t)npar * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_in, in, in_len * 8, cudaMemcpyHostToDevice, s);
  kern_launch(d_bnd, npar > 0 ? d_par : nullptr, d_in, d_out, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return c->cuda_fail(e, what);
  cudaMemcpyAsync(out, d_out, out_len * 8, cudaMemcpyDeviceToHost, s);
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return c->cuda_fail(e, what);
  return LFPSQP_OK;
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

extern "C" int lfpsqp_linesearch(lfpsqp_ctx *c, int which, int family, int64_t n, int64_t m, const double *fam_params,
                                 const double *x, const double *d, const double *xl, const double *xu, const lfpsqp_params *prm,
                                 double *xnew_out, double *out6, int *flag) {
  if (!c) return LFPSQP_ERR_ARG;
  if (!x || !d || !prm || !xnew_out || !out6 || n < 1 || m < 0 || which < 0 || which > 1) return c->fail(LFPSQP_ERR_ARG, "lfpsqp_linesearch: bad arguments");
  if (family != LFPSQP_FAM_BOXQUAD && family != LFPSQP_FAM_SIN)
    return c->fail(LFPSQP_ERR_FAMILY, "lfpsqp_linesearch: the unit-level export covers the BOXQUAD and SIN families");
  if (family == LFPSQP_FAM_BOXQUAD ? !FamBoxQuad::valid(n, m, 0) : !FamSin::valid(n, m, 0)) return c->fail(LFPSQP_ERR_FAMILY, "family does not support n=%lld m=%lld", (long long)n, (long long)m);
  cudaSetDevice(c->device);
  std::vector<double> bnd;
  const int ineq = build_bounds_public(n, 0, xl, xu, bnd);
  if (ineq == LFPSQP_ERR_BOUNDS) return c->fail(ineq, "Infeasible: lower bounds cannot be greater than upper bounds");
  if (ineq < 0) return c->fail(ineq, "xl and xu must both be given or both be NULL");
  lfpsqp_params p2 = *prm;
  const int use_nr = p2.do_project_retract ? 0 : 1;
  WarpLayout L((int)n, (int)m, 0, ineq, use_nr, 1);
  const int64_t N = L.N;
  const size_t smem = (size_t)((((ineq ? 5 * n : 0) + 1) & ~1) + L.total) * 8;
  const int64_t npar = (family == LFPSQP_FAM_BOXQUAD) ? 2 * n + 1 : n;
  if (!fam_params) return c->fail(LFPSQP_ERR_ARG, "family needs a parameter blob");
  std::vector<double> in(2 * N), out(N + 8);
  for (int64_t i = 0; i < N; i++) { in[i] = x[i]; in[N + i] = d[i]; }
  auto launch = [&](double *d_bnd, double *d_par, double *d_in, double *d_out, cudaStream_t s) {
    if (family == LFPSQP_FAM_BOXQUAD) {
      cudaFuncSetAttribute(linesearch_kernel<FamBoxQuad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      linesearch_kernel<FamBoxQuad><<<1, 32, smem, s>>>(which, (int)n, (int)m, ineq, d_bnd, d_par, d_in, d_out, p2);
    } else {
      cudaFuncSetAttribute(linesearch_kernel<FamSin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      linesearch_kernel<FamSin><<<1, 32, smem, s>>>(which, (int)n, (int)m, ineq, d_bnd, d_par, d_in, d_out, p2);
    }
  };
  int rc = run_unit(c, launch, smem, bnd, fam_params, npar, in.data(), 2 * N, out.data(), N + 8, "lfpsqp_linesearch");
  if (rc) return rc;
  for (int64_t i = 0; i < N; i++) xnew_out[i] = out[i];
  for (int i = 0; i < 6; i++) out6[i] = out[N + i];
  if (flag) *flag = (int)out[N + 6];
  return LFPSQP_OK;
}

extern "C" int lfpsqp_aug_hess_vec(lfpsqp_ctx *c, int family, int64_t n, int64_t m, const double *fam_params, const double *xl,
                                   const double *xu, const double *xaug, const double *lam, const double *lamy, const double *src,
                                   double *dest) {
  if (!c) return LFPSQP_ERR_ARG;
  if (!xl || !xu || !xaug || !lamy || !src || !dest || n < 1 || m < 0 || (m > 0 && !lam)) return c->fail(LFPSQP_ERR_ARG, "lfpsqp_aug_hess_vec: bad arguments");
  if (family != LFPSQP_FAM_BOXQUAD && family != LFPSQP_FAM_SIN)
    return c->fail(LFPSQP_ERR_FAMILY, "lfpsqp_aug_hess_vec: the unit-level export covers the BOXQUAD and SIN families");
  if (family == LFPSQP_FAM_BOXQUAD ? !FamBoxQuad::valid(n, m, 0) : !FamSin::valid(n, m, 0)) return c->fail(LFPSQP_ERR_FAMILY, "family does not support n=%lld m=%lld", (long long)n, (long long)m);
  if (!fam_params) return c->fail(LFPSQP_ERR_ARG, "family needs a parameter blob");
  cudaSetDevice(c->device);
  std::vector<double> bnd;
  int ineq = build_bounds_public(n, 0, xl, xu, bnd);
  if (ineq < 0) return c->fail(ineq, "Infeasible: lower bounds cannot be greater than upper bounds");
  if (ineq == 0) bnd.assign(5 * n, 0.0);
  WarpLayout L((int)n, (int)m, 0, 1, 0);
  const int64_t N = 2 * n;
  const size_t smem = (size_t)(((5 * n + 1) & ~1) + L.total) * 8;
  const int64_t npar = (family == LFPSQP_FAM_BOXQUAD) ? 2 * n + 1 : n;
  std::vector<double> in(2 * N + m + n);
  for (int64_t i = 0; i < N; i++) { in[i] = xaug[i]; in[N + i] = src[i]; }
  for (int64_t i = 0; i < m; i++) in[2 * N + i] = lam[i];
  for (int64_t i = 0; i < n; i++) in[2 * N + m + i] = lamy[i];
  lfpsqp_params prm; lfpsqp_default_params(&prm);
  auto launch = [&](double *d_bnd, double *d_par, double *d_in, double *d_out, cudaStream_t s) {
    if (family == LFPSQP_FAM_BOXQUAD) {
      cudaFuncSetAttribute(aug_hess_kernel<FamBoxQuad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      aug_hess_kernel<FamBoxQuad><<<1, 32, smem, s>>>((int)n, (int)m, d_bnd, d_par, d_in, d_out, prm);
    } else {
      cudaFuncSetAttribute(aug_hess_kernel<FamSin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      aug_hess_kernel<FamSin><<<1, 32, smem, s>>>((int)n, (int)m, d_bnd, d_par, d_in, d_out, prm);
    }
  };
  return run_unit(c, launch, smem, bnd, fam_params, npar, in.data(), (int64_t)in.size(), dest, N, "lfpsqp_aug_hess_vec");
}
