// batched_tiny.cuh -- one THREAD per instance for tiny unconstrained problems (BASELINE config C3: Rosenbrock, n=2).
//
// With m = 0 and no bounds the reference's driver (src/optimize.jl:257-435) reduces to: gradient, d = -g,
// projcg! with an empty projector (src/projcg.jl:40-121), armijo! with the Euclidean retraction
// (src/linesearch.jl:32-89, src/retractions.jl:61-65) and the termination tests (optimize.jl:347-359).
// All state lives in registers (NT is a compile-time constant); HBM only sees x0 in and the results out.
#pragma once
#include "common.cuh"
#include "families.cuh"

namespace lfpsqp {

struct ThreadGroup {  // a "group" of one thread, so the family callbacks can be shared with the warp kernel
  static constexpr int SIZE = 1;
  static constexpr int lane = 0;
  LFPSQP_DEV void sync() const {}
  LFPSQP_DEV double sum(double v) const { return v; }
  LFPSQP_DEV double maxabs(double v) const { return v; }
};

template <class Fam, int NT>
__global__ void __launch_bounds__(128) batched_tiny_kernel(const BatchedArgs A) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.B) return;
  const lfpsqp_params &prm = A.prm;
  ThreadGroup g;
  FamCtx fc; fc.n = NT; fc.m = 0; fc.p = 0; fc.prm = A.fam_params ? A.fam_params + k * A.fam_stride : nullptr;
  double x[NT], xnew[NT], gr[NT], d[NT], nd[NT], r[NT], dc[NT], Ad[NT];
  lfpsqp_stats st = lfpsqp_stats{};
#pragma unroll
  for (int i = 0; i < NT; i++) x[i] = A.x0[k * NT + i];
  int64_t it = 0, nobj = 0;
  double f_diff = INFINITY, step_diff = INFINITY, kkt_diff = INFINITY, prev_grad_norm = 0.0;
  double fval = Fam::f(g, fc, x); st.f_evals++;
  if (nobj < A.H) A.obj_hist[k * A.H + nobj] = fval;
  nobj++;
  int cond = LFPSQP_F_TOL, status = 0;
  while (true) {
    Fam::grad(g, fc, gr, x);                                               // optimize.jl:259
    kkt_diff = 0.0;
#pragma unroll
    for (int i = 0; i < NT; i++) d[i] = -1.0 * gr[i];                                              // :262
    if (A.noise) {                                                                                   // :264-273
      const double nc = noise_coef(prm, it, A.noise_T);
      if (nc != 0.0) { const double *nz = A.noise + ((int64_t)k * A.noise_T + it) * NT;
#pragma unroll
        for (int i = 0; i < NT; i++) d[i] += nc * nz[i]; }
    }
#pragma unroll
    for (int i = 0; i < NT; i++) kkt_diff = pmax(fabs(d[i]), kkt_diff);                            // :320
    if (f_diff <= prm.eps_f) { cond = LFPSQP_F_TOL; break; }               // :347-359
    else if (step_diff <= prm.eps_x) { cond = LFPSQP_X_TOL; break; }
    else if (it >= prm.maxiter) { cond = LFPSQP_MAX_ITER; break; }
    else if (kkt_diff <= prm.eps_kkt) { cond = LFPSQP_KKT_TOL; break; }
    if (!(kkt_diff == kkt_diff)) { status |= LFPSQP_ST_NONFINITE; cond = LFPSQP_MAX_ITER; break; }
    if (prm.do_newton) {                                                   // :364-390
      double gn = 0.0;
#pragma unroll
      for (int i = 0; i < NT; i++) gn += d[i] * d[i];
      gn = sqrt(gn);
      const double tol = prm.tn_kappa * fmin(1.0, gn / prev_grad_norm) * gn;
      prev_grad_norm = gn;
      // projcg! with U = n x 0 (projcg.jl:55-62): x = 0, r = g = -b, d = b
#pragma unroll
      for (int i = 0; i < NT; i++) { nd[i] = 0.0; r[i] = -d[i]; dc[i] = -1.0 * r[i]; }
      int i = 0;
      const int64_t lim = prm.tn_maxiter < NT ? prm.tn_maxiter : NT;
      bool negcurv = false;
      while (i < lim) {
        i++;
        Fam::hess(g, fc, Ad, dc, x, nullptr, nullptr);
        double dAd = 0.0, rg = 0.0;
#pragma unroll
        for (int q = 0; q < NT; q++) { dAd += dc[q] * Ad[q]; rg += r[q] * r[q]; }
        if (dAd <= 0.0) {                                                  // projcg.jl:77-82
          double nrm = 0.0;
#pragma unroll
          for (int q = 0; q < NT; q++) nrm += dc[q] * dc[q];
          nrm = sqrt(nrm);
#pragma unroll
          for (int q = 0; q < NT; q++) nd[q] = dc[q] / nrm;
          negcurv = true; st.projcg_negcurv++;
          break;
        }
        if (rg <= 0.0) break;
        const double alpha = rg / dAd;
        double rpgp = 0.0;
#pragma unroll
        for (int q = 0; q < NT; q++) { nd[q] += alpha * dc[q]; r[q] = r[q] + alpha * Ad[q]; rpgp += r[q] * r[q]; }
        const double beta = rpgp / rg;
#pragma unroll
        for (int q = 0; q < NT; q++) dc[q] = beta * dc[q] - r[q];
        if (sqrt(rpgp) < tol) break;                                       // projcg.jl:103-111
      }
      (void)negcurv;
      st.projcg_iters += i;
      double nd_d = 0.0;
#pragma unroll
      for (int q = 0; q < NT; q++) nd_d += nd[q] * d[q];
      if (nd_d > 0.0) {                                                    // optimize.jl:386-389
#pragma unroll
        for (int q = 0; q < NT; q++) d[q] = nd[q];
        st.newton_accepted++;
      }
    }
    // armijo! with Euclidean retraction (linesearch.jl:32-89)
    double alpha = prm.alpha, ar_dot = 0.0, newf = 0.0;
#pragma unroll
    for (int q = 0; q < NT; q++) ar_dot += d[q] * gr[q];
    f_diff = INFINITY; step_diff = INFINITY;
    int flag = 0;
    while (step_diff > prm.eps_x) {
#pragma unroll
      for (int q = 0; q < NT; q++) xnew[q] = x[q] + alpha * d[q];
      st.armijo_trials++;
      newf = Fam::f(g, fc, xnew); st.f_evals++;
      double s2 = 0.0;
#pragma unroll
      for (int q = 0; q < NT; q++) { double t = xnew[q] - x[q]; s2 += t * t; }
      step_diff = sqrt(s2);
      f_diff = fabs(newf - fval);
      if (prm.disable_linesearch) break;
      if ((newf - fval) <= prm.sigma * alpha * ar_dot) break;
      alpha *= prm.s;
      if (alpha < 1e-100) { flag = 99; break; }
    }
    st.flag_last = flag;
#pragma unroll
    for (int q = 0; q < NT; q++) x[q] = xnew[q];
    fval = newf;
    if (nobj < A.H) A.obj_hist[k * A.H + nobj] = fval;
    nobj++;
    it++;
  }
#pragma unroll
  for (int i = 0; i < NT; i++) A.x_out[k * NT + i] = x[i];
  A.obj_len[k] = nobj;
  lfpsqp_term t; t.condition = cond; t.status = status; t.f_diff = f_diff; t.step_diff = step_diff;
  t.kkt_diff = kkt_diff; t.iter = it;
  A.term[k] = t;
  if (A.stats) A.stats[k] = st;
}

}  // namespace lfpsqp
