// large_ctrl.h -- device-resident control block of the large-n mode and the reduction-slot geometry.
#pragma once
#include <stdint.h>

namespace lfpsqp {

constexpr int MAXP = 2048;   // max per-CTA partials per reduction slot
constexpr int NSLOT = 8;

// device-resident control block (one per ctx); mirrored to pinned host memory when the host needs a decision
struct LargeCtrl {
  double s[16];              // generic scalar results (finalize_kernel)
  double rg, dAd, alpha, beta, nr, tol, rpgp, gg, mu, rho, pz, norm_res;
  int iter, lim, status;     // projcg: 0 running, 1 nr<tol, 2 negative curvature, 3 rg<=0, 4 iteration limit
  int pcg_iter, pcg_lim, pcg_status;  // pcg: 0 running, 1 converged (norm_res<=tol), 4 limit
  int rankflag, pad;
};

// bound embedding of the large-n mode (kernels in large_ineq.cuh); device arrays, each nx doubles
struct IneqDev {
  const double *q = nullptr, *r = nullptr, *s = nullptr, *t = nullptr;   // InequalityData (inequality_helper.jl:1-8, :54-82)
  double *Dx = nullptr, *Dy = nullptr, *S = nullptr, *lamy = nullptr;   // inequality_gradient! output, lambda_y
  int64_t nx = 0;
};

}  // namespace lfpsqp
