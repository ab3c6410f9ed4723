// batched_reg.cuh -- register-resident warp-per-instance solver for SEPARABLE families (BASELINE config C2).
//
// Same algorithm, same reference line mapping as batched_warp.cuh (the shared-memory solver), but every vector of
// the instance lives in REGISTERS: element j of an N_A-vector sits in lane j%32, slot j/32 (NPL slots per lane), the
// x- and y-halves of the 2n-embedding (src/inequality_helper.jl) are two register arrays.  Elementwise work needs no
// synchronisation at all; the only cross-lane traffic is the xor-butterfly of the dot products.  The m_E (<= 2) rows
// of J, the m_E x m_E Gram/Cholesky factor and all m_E-vectors are replicated in registers of every lane.
//
// "Separable" = f(x) = sum_j f_j(x_j), every constraint c_a(x) = sum_j c_aj(x_j) - off_a, hence a diagonal Lagrangian
// Hessian: README equality / inequality examples (README.md:41-76), the bounded quadratic, diagonal-quadratic rows.
// Anything else (Thomson, Rosenbrock, sin system, m_E > 2, n_A > 128) runs in batched_warp.cuh.
#pragma once
#include "common.cuh"

namespace lfpsqp {

// ---------------------------------------------------------------- separable family descriptions
// row a < m is an equality c_a, row a >= m an inequality d_{a-m}; ep = this element's cached parameters (KP doubles)
struct SepReadmeIneq {  // f = coeff.x ; d = x.x - 1     (README.md:57-76)
  static constexpr int kId = LFPSQP_FAM_README_INEQ, KP = 1;
  static LFPSQP_DEV void load(const FamCtx &fc, int j, double *ep) { ep[0] = __ldg(fc.prm + j); }
  static LFPSQP_DEV double f(const FamCtx &, int, double x, const double *ep) { return ep[0] * x; }
  static LFPSQP_DEV double g(const FamCtx &, int, double, const double *ep) { return ep[0]; }
  static LFPSQP_DEV double h(const FamCtx &, int, double, const double *, const double *lam) { return 2.0 * lam[0]; }
  static LFPSQP_DEV double c(const FamCtx &, int, int, double x, const double *) { return x * x; }
  static LFPSQP_DEV double off(const FamCtx &, int) { return 1.0; }
  static LFPSQP_DEV double jac(const FamCtx &, int, int, double x, const double *) { return 2.0 * x; }
};
struct SepReadmeEq {  // f = x.x ; c = x[1] - 0.75          (README.md:41-54)
  static constexpr int kId = LFPSQP_FAM_README_EQ, KP = 0;
  static LFPSQP_DEV void load(const FamCtx &, int, double *) {}
  static LFPSQP_DEV double f(const FamCtx &, int, double x, const double *) { return x * x; }
  static LFPSQP_DEV double g(const FamCtx &, int, double x, const double *) { return 2.0 * x; }
  static LFPSQP_DEV double h(const FamCtx &, int, double, const double *, const double *) { return 2.0; }
  static LFPSQP_DEV double c(const FamCtx &, int, int j, double x, const double *) { return j == 0 ? x : 0.0; }
  static LFPSQP_DEV double off(const FamCtx &, int) { return 0.75; }
  static LFPSQP_DEV double jac(const FamCtx &, int, int j, double, const double *) { return j == 0 ? 1.0 : 0.0; }
};
struct SepBoxQuad {  // f = |x-t|^2 ; optional c = a.x - b ; params [t(n), a(n), b]
  static constexpr int kId = LFPSQP_FAM_BOXQUAD, KP = 2;
  static LFPSQP_DEV void load(const FamCtx &fc, int j, double *ep) { ep[0] = __ldg(fc.prm + j); ep[1] = __ldg(fc.prm + fc.n + j); }
  static LFPSQP_DEV double f(const FamCtx &, int, double x, const double *ep) { double u = x - ep[0]; return u * u; }
  static LFPSQP_DEV double g(const FamCtx &, int, double x, const double *ep) { return 2.0 * (x - ep[0]); }
  static LFPSQP_DEV double h(const FamCtx &, int, double, const double *, const double *) { return 2.0; }
  static LFPSQP_DEV double c(const FamCtx &, int, int, double x, const double *ep) { return ep[1] * x; }
  static LFPSQP_DEV double off(const FamCtx &fc, int) { return __ldg(fc.prm + 2 * fc.n); }
  static LFPSQP_DEV double jac(const FamCtx &, int, int, double, const double *ep) { return ep[1]; }
};

LFPSQP_DEV double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
LFPSQP_DEV double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = pmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// two independent butterflies interleaved (same instruction count, half the dependent-latency chain)
LFPSQP_DEV void wsum2(double &a, double &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a += ta; b += tb;
  }
}

#define LF_UNROLL _Pragma("unroll")

template <class Fam, int NPL, int ME, bool INEQ>
struct RegSolver {
  static constexpr int KP = Fam::KP > 0 ? Fam::KP : 1;
  static constexpr int MEA = ME > 0 ? ME : 1;
  struct Vec { double x[NPL]; double y[INEQ ? NPL : 1]; };

  const int lane;
  const lfpsqp_params &prm;
  FamCtx fc;
  const int n, m, p, NA;
  // bound data of this lane's elements (inequality_helper.jl:1-8); kind: 0 line, 1 parabola, 2 circle
  double bq[NPL], br[NPL], bs[NPL], bt[NPL]; int bkind[NPL];
  bool valid[NPL];       // element index < NA
  double ep[NPL][KP];    // cached per-element family parameters
  // instance state
  Vec x;
  double J[MEA][NPL], Lc[MEA][MEA], cval[MEA], lam[MEA];
  double Dx[NPL], Dy[NPL], S[NPL], Sinv[NPL], lamy[NPL];
  double cvh[NPL], cvc[MEA];   // cvalaug = [h ; c] (retractions.jl:29), persistent across PP calls (stale-tail quirk)
  int st_projcg, st_negcurv, st_trials, st_rout, st_rpcg, st_bt, st_newton, st_fact, st_feval, status;
  int rank; bool pinv;   // numerical rank (optimize.jl:297-302); Lc holds the truncated pseudo-inverse G^+ when pinv

  LFPSQP_DEV RegSolver(int lane_, const lfpsqp_params &prm_, int n_, int m_, int p_)
      : lane(lane_), prm(prm_), n(n_), m(m_), p(p_), NA(n_ + p_) { fc.n = n_; fc.m = m_; fc.p = p_; fc.prm = nullptr; }

  LFPSQP_DEV int idx(int s) const { return s * 32 + lane; }

  // ---------------------------------------------------------------- vector helpers
  LFPSQP_DEV double dot(const Vec &a, const Vec &b) const {
    double s = 0.0;
    LF_UNROLL for (int k = 0; k < NPL; k++) s += a.x[k] * b.x[k];
    if (INEQ) { LF_UNROLL for (int k = 0; k < NPL; k++) s += a.y[k] * b.y[k]; }
    return wsum(s);
  }
  LFPSQP_DEV void zero(Vec &a) const {
    LF_UNROLL for (int k = 0; k < NPL; k++) { a.x[k] = 0.0; if (INEQ) a.y[k] = 0.0; }
  }
  // out[a] = sum_j J[a][j] v_j  (J acts on the x-half only)
  LFPSQP_DEV void rowdots(double *out, const double *vx) const {
    LF_UNROLL for (int a = 0; a < ME; a++) {
      double s = 0.0;
      LF_UNROLL for (int k = 0; k < NPL; k++) s += J[a][k] * vx[k];
      out[a] = wsum(s);
    }
  }
  LFPSQP_DEV double coldot(const double *u, int k) const {
    double s = 0.0;
    LF_UNROLL for (int a = 0; a < ME; a++) s += J[a][k] * u[a];
    return s;
  }
  LFPSQP_DEV void solveG(double *u) const {  // u <- (L L')^-1 u (or G^+ u), replicated scalar code
    if (pinv) {
      double t[MEA];
      LF_UNROLL for (int a = 0; a < ME; a++) { double s = 0.0; LF_UNROLL for (int b = 0; b < ME; b++) s += Lc[a][b] * u[b]; t[a] = s; }
      LF_UNROLL for (int a = 0; a < ME; a++) u[a] = t[a];
      return;
    }
    LF_UNROLL for (int k = 0; k < ME; k++) {
      double s = u[k];
      LF_UNROLL for (int t = 0; t < k; t++) s -= Lc[k][t] * u[t];
      u[k] = s / Lc[k][k];
    }
    LF_UNROLL for (int k = ME - 1; k >= 0; k--) {
      double s = u[k];
      LF_UNROLL for (int t = k + 1; t < ME; t++) s -= Lc[t][k] * u[t];
      u[k] = s / Lc[k][k];
    }
  }

  // ---------------------------------------------------------------- callbacks on the slack-augmented problem (optimize.jl:38-51)
  LFPSQP_DEV double f_aux(const Vec &v) {
    st_feval++;
    double s = 0.0;
    LF_UNROLL for (int k = 0; k < NPL; k++) if (idx(k) < n) s += Fam::f(fc, idx(k), v.x[k], ep[k]);
    return wsum(s);
  }
  LFPSQP_DEV void c_aux(double *cv, const Vec &v) const {
    LF_UNROLL for (int a = 0; a < ME; a++) {
      double s = 0.0;
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        int j = idx(k);
        if (j < n) s += Fam::c(fc, a, j, v.x[k], ep[k]);
        else if (a >= m && j == n + (a - m)) s -= v.x[k];       // d_k(x) - s_k
      }
      cv[a] = wsum(s) - Fam::off(fc, a);
    }
  }
  LFPSQP_DEV void jac_aux(double *cv, const Vec &v) {   // jac!(Jc, cval, x): fills J and cval
    LF_UNROLL for (int a = 0; a < ME; a++)
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        int j = idx(k);
        J[a][k] = (j < n) ? Fam::jac(fc, a, j, v.x[k], ep[k]) : ((a >= m && j == n + (a - m)) ? -1.0 : 0.0);
      }
    c_aux(cv, v);
  }
  LFPSQP_DEV void hess_aux(Vec &dest, const Vec &src) const {   // Lagrangian Hessian at the current (x, lam, lamy)
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      int j = idx(k);
      double hx = (j < n) ? Fam::h(fc, j, x.x[k], ep[k], lam) * src.x[k] : 0.0;
      if (INEQ) {
        double ly2 = 2.0 * lamy[k];
        hx += ly2 * bq[k] * src.x[k];
        dest.y[k] = ly2 * bs[k] * src.y[k];
      }
      dest.x[k] = hx;
    }
  }

  // ---------------------------------------------------------------- bound embedding (inequality_helper.jl)
  LFPSQP_DEV void generate_initial_y(Vec &v) const {  // :92-109
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      double xv = v.x[k], yv;
      if (bkind[k] == 0) yv = xv;
      else if (bkind[k] == 1) yv = sqrt(fmax(-(xv - bt[k]) / bs[k], 0.0)) + br[k];
      else yv = sqrt(fmax(bt[k] - (xv - br[k]) * (xv - br[k]), 0.0)) + br[k];
      v.y[k] = valid[k] ? yv : 0.0;
    }
  }
  LFPSQP_DEV void calculate_h(double *out, const Vec &v) const {  // :112-122
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      double q = bq[k], s = bs[k], r = br[k], dx = v.x[k] - r, dy = v.y[k] - r;
      // padded elements have q = s = r = t = 0 and x = y = 0: the formula itself gives 0 there
      out[k] = q * (dx * dx) + (1.0 - q * q) * v.x[k] + s * (dy * dy) - (1.0 - s * s) * v.y[k] - bt[k];
    }
  }
  LFPSQP_DEV void inequality_gradient(const Vec &v) {  // :125-141
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      double q = bq[k], s = bs[k], r = br[k];
      double dx = 2.0 * q * (v.x[k] - r) + (q == 0.0 ? 1.0 : 0.0);
      double dy = 2.0 * s * (v.y[k] - r) - (s == 0.0 ? 1.0 : 0.0);
      double sv = sqrt(dx * dx + dy * dy), inv = 1.0 / sv;   // one reciprocal instead of two divisions (<= 1 ulp apart)
      S[k] = sv; Sinv[k] = inv; Dx[k] = dx * inv; Dy[k] = dy * inv;
    }
  }
  LFPSQP_DEV void y_retract(Vec &vn, const Vec &vb) const {  // retractions.jl:451-500
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      if (!valid[k]) continue;
      if (bkind[k] == 0) { vn.x[k] = vn.y[k]; }
      else if (bkind[k] == 1) {
        double s = bs[k], r = br[k];
        double g1 = -s, g2 = -2.0 * (vb.y[k] - r), ng = sqrt(g1 * g1 + g2 * g2);
        double ux = vb.x[k] - vn.x[k] + g1 / ng, uy = vb.y[k] - vn.y[k] + g2 / ng;
        double yn = vn.y[k] - r;
        double a = s * uy * uy, b = ux + 2.0 * s * yn * uy, c = vn.x[k] + s * yn * yn - r;
        double a1 = -b / (2.0 * a), a2 = sqrt(b * b - 4.0 * a * c) / (2.0 * a);
        double gam = fmin(a1 + a2, a1 - a2);
        vn.x[k] += gam * ux; vn.y[k] += gam * uy;
      } else {
        double c = br[k], rho = sqrt(bt[k]);
        double ex = vn.x[k] - c, ey = vn.y[k] - c, dist = sqrt(ex * ex + ey * ey);
        vn.y[k] = c + rho * ey / dist;
        vn.x[k] = c + rho * ex / dist;
      }
    }
  }
  // fulljac * v (retractions.jl:324): J v, or bigA' v (inequality_helper.jl:254-271) -> [oh ; oc]
  LFPSQP_DEV void fullJ_mul(double *oh, double *oc, const Vec &v) const {
    if (INEQ) { LF_UNROLL for (int k = 0; k < NPL; k++) oh[k] = S[k] * (Dx[k] * v.x[k] + Dy[k] * v.y[k]); }
    rowdots(oc, v.x);
  }
  // dest = a * fulljac' [wh ; wc] + b * dest (inequality_helper.jl:215-251)
  LFPSQP_DEV void fullJ_mulT(Vec &dest, const double *wh, const double *wc, double a, double b) const {
    LF_UNROLL for (int k = 0; k < NPL; k++) {
      double t = coldot(wc, k);
      if (INEQ) {
        double sw = S[k] * wh[k];
        dest.x[k] = a * (t + Dx[k] * sw) + (b == 0.0 ? 0.0 : b * dest.x[k]);
        dest.y[k] = (b == 0.0 ? 0.0 : b * dest.y[k]) + a * Dy[k] * sw;
      } else dest.x[k] = a * t + (b == 0.0 ? 0.0 : b * dest.x[k]);
    }
  }

  // ---------------------------------------------------------------- Gram + Cholesky (replaces ksvd!, optimize.jl:288-302)
  LFPSQP_DEV bool factor() {
    st_fact++;
    rank = ME; pinv = false;
    double maxdiag = 0.0, Gs[MEA][MEA];
    LF_UNROLL for (int a = 0; a < ME; a++)
      LF_UNROLL for (int b = 0; b <= a; b++) {
        double s = 0.0;
        LF_UNROLL for (int k = 0; k < NPL; k++) {
          double w = INEQ ? Dy[k] * Dy[k] : 1.0;
          s += J[a][k] * w * J[b][k];
        }
        s = wsum(s);
        Lc[a][b] = s; Gs[a][b] = s; Gs[b][a] = s;
        if (a == b) maxdiag = fmax(maxdiag, s);
      }
    const double thresh = fmax(prm.eps_rank * prm.eps_rank, 1e-14 * maxdiag);
    bool ok = true;
    LF_UNROLL for (int k = 0; k < ME; k++) {
      LF_UNROLL for (int i = k; i < ME; i++) {
        double s = Lc[i][k];
        LF_UNROLL for (int t = 0; t < k; t++) s -= Lc[i][t] * Lc[k][t];
        Lc[i][k] = s;
      }
      double piv = Lc[k][k];
      if (!(piv > thresh)) { ok = false; piv = 1.0; }
      double rinv = 1.0 / sqrt(piv);
      Lc[k][k] = sqrt(piv);
      LF_UNROLL for (int i = k + 1; i < ME; i++) Lc[i][k] *= rinv;
    }
    if (ok) return true;
    // rank-deficient (optimize.jl:297-302): truncated pseudo-inverse of G through its eigen-decomposition (ME <= 2:
    // one Jacobi rotation is exact); same thresholds as batched_warp.cuh::factor_rank_deficient
    double V[MEA][MEA], ev[MEA];
    LF_UNROLL for (int a = 0; a < ME; a++) LF_UNROLL for (int b = 0; b < ME; b++) V[a][b] = (a == b) ? 1.0 : 0.0;
    if (ME == 2 && Gs[0][MEA - 1] != 0.0) {
      const double apq = Gs[0][MEA - 1], app = Gs[0][0], aqq = Gs[MEA - 1][MEA - 1];
      const double theta = (aqq - app) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      V[0][0] = c; V[0][MEA - 1] = sn; V[MEA - 1][0] = -sn; V[MEA - 1][MEA - 1] = c;
      Gs[0][0] = app - t * apq; Gs[MEA - 1][MEA - 1] = aqq + t * apq;
    }
    double lmax = 0.0;
    LF_UNROLL for (int k = 0; k < ME; k++) lmax = fmax(lmax, Gs[k][k]);
    const double thr = fmax(prm.eps_rank * prm.eps_rank, 1e-13 * lmax);
    int r = 0;
    LF_UNROLL for (int k = 0; k < ME; k++) { bool keep = Gs[k][k] >= thr; ev[k] = keep ? 1.0 / Gs[k][k] : 0.0; r += keep ? 1 : 0; }
    LF_UNROLL for (int a = 0; a < ME; a++) LF_UNROLL for (int b = 0; b < ME; b++) {
      double s = 0.0;
      LF_UNROLL for (int k = 0; k < ME; k++) s += V[a][k] * ev[k] * V[b][k];
      Lc[a][b] = s;
    }
    rank = r; pinv = true;
    return true;
  }

  // v <- v - Q Q' v (optimize.jl:306-307 / :316-317, projcg.jl:59-60,:96-97); multipliers as by-product (:331-343)
  LFPSQP_DEV void project(Vec &v, bool want_mult) {
    double u[MEA];
    if (INEQ) {
      double aa[NPL], bb[NPL];
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        aa[k] = Dx[k] * v.x[k] + Dy[k] * v.y[k];
        bb[k] = Dy[k] * (Dy[k] * v.x[k] - Dx[k] * v.y[k]);
      }
      if (ME > 0) { rowdots(u, bb); solveG(u); }
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        double wj = (ME > 0) ? coldot(u, k) : 0.0;
        v.x[k] -= Dx[k] * aa[k] + Dy[k] * Dy[k] * wj;
        v.y[k] -= Dy[k] * aa[k] - Dx[k] * Dy[k] * wj;
        if (want_mult) lamy[k] = (-1.0 * Dx[k] * Sinv[k]) * wj + aa[k] * Sinv[k];
      }
    } else if (ME > 0) {
      rowdots(u, v.x); solveG(u);
      LF_UNROLL for (int k = 0; k < NPL; k++) v.x[k] -= coldot(u, k);
    }
    if (want_mult) { LF_UNROLL for (int a = 0; a < ME; a++) lam[a] = u[a]; }
  }

  // ---------------------------------------------------------------- projcg! (projcg.jl:40-121), c = 0
  LFPSQP_DEV void projcg(Vec &xs, const Vec &b, double tol, int64_t maxit) {
    Vec r, dc, Ad, rp, gp;
    LF_UNROLL for (int k = 0; k < NPL; k++) { xs.x[k] = 0.0; r.x[k] = -b.x[k]; if (INEQ) { xs.y[k] = 0.0; r.y[k] = -b.y[k]; } }
    project(r, false);
    LF_UNROLL for (int k = 0; k < NPL; k++) { dc.x[k] = -1.0 * r.x[k]; if (INEQ) dc.y[k] = -1.0 * r.y[k]; }
    int i = 0;
    const int N = INEQ ? 2 * NA : NA;
    int64_t lim = (int64_t)N + (INEQ ? NA + rank : rank); if (maxit < lim) lim = maxit;   // c has length rank / n+rank
    double rg = dot(r, r);                                                      // r == g after every projection
    while (i < lim) {
      i++;
      hess_aux(Ad, dc);
      double dAd = dot(dc, Ad);
      if (dAd <= 0.0) {                                                         // :77-82
        double nrm = sqrt(dot(dc, dc));
        LF_UNROLL for (int k = 0; k < NPL; k++) { xs.x[k] = dc.x[k] / nrm; if (INEQ) xs.y[k] = dc.y[k] / nrm; }
        st_negcurv++;
        break;
      }
      if (rg <= 0.0) break;                                                     // :87-89
      double alpha = rg / dAd;
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        xs.x[k] += alpha * dc.x[k];
        double t = r.x[k] + alpha * Ad.x[k]; rp.x[k] = t; gp.x[k] = t;
        if (INEQ) { xs.y[k] += alpha * dc.y[k]; double ty = r.y[k] + alpha * Ad.y[k]; rp.y[k] = ty; gp.y[k] = ty; }
      }
      project(gp, false);                                                       // :95-97
      double rpgp = 0.0, gg = 0.0;                                              // rp.gp and gp.gp in one interleaved pass
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        rpgp += rp.x[k] * gp.x[k]; gg += gp.x[k] * gp.x[k];
        if (INEQ) { rpgp += rp.y[k] * gp.y[k]; gg += gp.y[k] * gp.y[k]; }
      }
      wsum2(rpgp, gg);
      double beta = rpgp / rg;
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        dc.x[k] = beta * dc.x[k] - gp.x[k]; r.x[k] = gp.x[k];
        if (INEQ) { dc.y[k] = beta * dc.y[k] - gp.y[k]; r.y[k] = gp.y[k]; }
      }
      rg = gg;                                                                  // next iteration's r.g (:84) == |g|^2
      if (sqrt(gg) < tol) break;                                                // :103-111
    }
    st_projcg += i;
  }

  // ---------------------------------------------------------------- retract!(::ProjPenalty) (retractions.jl:265-441) with pcg! (:179-246)
  LFPSQP_DEV int retract_pp(Vec &xnew, const Vec &xtil, int *it1, int *it2) {
    Vec r, pv, z, dx, gv;
    int flag = 0;
    xnew = xtil;
    double mu = prm.mu0;
    int i = 0, pcg_total = 0;
    while (i < prm.maxiter_retract) {
      jac_aux(cval, xnew);                                              // :340
      double curtol = 0.0;
      LF_UNROLL for (int a = 0; a < ME; a++) curtol = pmax(fabs(cval[a]), curtol);
      if (INEQ) {
        inequality_gradient(xnew);                                      // :344
        calculate_h(cvh, xnew);                                         // :350
        double hm = 0.0;
        LF_UNROLL for (int k = 0; k < NPL; k++) hm = pmax(fabs(cvh[k]), hm);
        hm = wmax(hm);
        LF_UNROLL for (int a = 0; a < ME; a++) hm = pmax(fabs(cvc[a]), hm);   // :352 incl. the stale tail of cvalaug
        curtol = pmax(curtol, hm);
      }
      LF_UNROLL for (int a = 0; a < ME; a++) cvc[a] = cval[a];          // :356
      if (curtol < prm.eps_c) break;                                    // :359-361
      double gg = 0.0, hh = 0.0;
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        gv.x[k] = xnew.x[k] - xtil.x[k]; gg += gv.x[k] * gv.x[k];
        if (INEQ) { gv.y[k] = xnew.y[k] - xtil.y[k]; gg += gv.y[k] * gv.y[k]; hh += cvh[k] * cvh[k]; }
      }
      double cc = 0.0;
      LF_UNROLL for (int a = 0; a < ME; a++) cc += cvc[a] * cvc[a];
      if (INEQ) wsum2(hh, gg); else gg = wsum(gg);
      const double prev_obj = (hh + cc) + mu * gg;                      // :366
      fullJ_mulT(gv, cvh, cvc, 1.0, mu);                                // :369
      r = gv; zero(dx); zero(pv);
      // pcg! (:179-246), M! = copy.  rho_k = r.r is reduced once per iteration (it is both norm(r)^2 of :235 and
      // dot(z,r) of :213) together with J r, from which J p follows by the p-recurrence (J p = J r + beta J p_old).
      int pi = 0; double norm_res = INFINITY, rho_prev = 1.0, rho, Jr[MEA], Jp[MEA];
      {
        double a0 = 0.0, b0 = 0.0;
        LF_UNROLL for (int k = 0; k < NPL; k++) { a0 += r.x[k] * r.x[k]; if (INEQ) a0 += r.y[k] * r.y[k]; if (ME > 0) b0 += J[0][k] * r.x[k]; }
        if (ME > 0) { wsum2(a0, b0); Jr[0] = b0; } else a0 = wsum(a0);
        rho = a0;
        LF_UNROLL for (int a = 1; a < ME; a++) { double s = 0.0; LF_UNROLL for (int k = 0; k < NPL; k++) s += J[a][k] * r.x[k]; Jr[a] = wsum(s); }
        LF_UNROLL for (int a = 0; a < MEA; a++) Jp[a] = 0.0;
      }
      while (norm_res > prm.eps_c && pi < prm.maxiter_pcg) {
        const double beta = rho / rho_prev;
        LF_UNROLL for (int k = 0; k < NPL; k++) { pv.x[k] = r.x[k] + beta * pv.x[k]; if (INEQ) pv.y[k] = r.y[k] + beta * pv.y[k]; }
        LF_UNROLL for (int a = 0; a < ME; a++) Jp[a] = Jr[a] + beta * Jp[a];
        double th[NPL];
        if (INEQ) { LF_UNROLL for (int k = 0; k < NPL; k++) th[k] = S[k] * (Dx[k] * pv.x[k] + Dy[k] * pv.y[k]); }
        z = pv;
        fullJ_mulT(z, th, Jp, 1.0, mu);
        const double alpha = rho / dot(pv, z);
        double a0 = 0.0, b0 = 0.0;
        LF_UNROLL for (int k = 0; k < NPL; k++) {
          dx.x[k] += alpha * pv.x[k]; r.x[k] -= alpha * z.x[k]; a0 += r.x[k] * r.x[k];
          if (ME > 0) b0 += J[0][k] * r.x[k];
          if (INEQ) { dx.y[k] += alpha * pv.y[k]; r.y[k] -= alpha * z.y[k]; a0 += r.y[k] * r.y[k]; }
        }
        if (ME > 0) { wsum2(a0, b0); Jr[0] = b0; } else a0 = wsum(a0);
        LF_UNROLL for (int a = 1; a < ME; a++) { double s = 0.0; LF_UNROLL for (int k = 0; k < NPL; k++) s += J[a][k] * r.x[k]; Jr[a] = wsum(s); }
        rho_prev = rho; rho = a0;
        norm_res = sqrt(rho);
        pi++;
      }
      pcg_total += pi;
      if (pi == prm.maxiter_pcg) { flag = 2; break; }                   // :240-243, :377-381
      const double ar_dot = -dot(gv, dx);
      double alpha = 1.0, s2 = 0.0;
      pv = xnew;                                                        // :384
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        xnew.x[k] -= alpha * dx.x[k]; gv.x[k] = xnew.x[k] - xtil.x[k]; s2 += gv.x[k] * gv.x[k];
        if (INEQ) { xnew.y[k] -= alpha * dx.y[k]; gv.y[k] = xnew.y[k] - xtil.y[k]; s2 += gv.y[k] * gv.y[k]; }
      }
      double dist2 = wsum(s2);
      c_aux(cval, xnew);                                                // :392
      if (INEQ) calculate_h(cvh, xnew);
      LF_UNROLL for (int a = 0; a < ME; a++) cvc[a] = cval[a];
      int armijo_count = 0;
      while (true) {
        double h2 = 0.0, c2 = 0.0;
        if (INEQ) { LF_UNROLL for (int k = 0; k < NPL; k++) h2 += cvh[k] * cvh[k]; h2 = wsum(h2); }
        LF_UNROLL for (int a = 0; a < ME; a++) c2 += cvc[a] * cvc[a];
        if (!((h2 + c2) + mu * dist2 > prev_obj + 1e-4 * alpha * ar_dot)) break;   // :403
        alpha /= 2; s2 = 0.0;
        LF_UNROLL for (int k = 0; k < NPL; k++) {
          xnew.x[k] = pv.x[k] - alpha * dx.x[k]; gv.x[k] = xnew.x[k] - xtil.x[k]; s2 += gv.x[k] * gv.x[k];
          if (INEQ) { xnew.y[k] = pv.y[k] - alpha * dx.y[k]; gv.y[k] = xnew.y[k] - xtil.y[k]; s2 += gv.y[k] * gv.y[k]; }
        }
        dist2 = wsum(s2);
        if (INEQ) calculate_h(cvh, xnew);   // :410-417: only the bound part is refreshed, the c-part stays frozen
        armijo_count++; st_bt++;
        if (armijo_count == 100) { flag = 3; break; }
      }
      i++;
      double nn = 0.0;
      if (INEQ) { LF_UNROLL for (int k = 0; k < NPL; k++) nn += cvh[k] * cvh[k]; nn = wsum(nn); }
      LF_UNROLL for (int a = 0; a < ME; a++) nn += cvc[a] * cvc[a];
      mu = fmin(mu * 0.1, sqrt(nn));                                    // :431
    }
    if (i == prm.maxiter_retract) flag = 1;
    *it1 = i; *it2 = pcg_total;
    return flag;
  }

  // ---------------------------------------------------------------- retract!(::NR) (retractions.jl:75-177), Cholesky-QR basis
  LFPSQP_DEV int retract_nr(Vec &xnew, const Vec &xtil, int *it1) {
    double D[MEA][MEA], t1[MEA], t2[MEA], dcv[MEA];
    xnew = xtil;
    if (INEQ) y_retract(xnew, x);
    c_aux(cval, xnew);
    LF_UNROLL for (int c = 0; c < ME; c++)        // D0 = L^-1
      LF_UNROLL for (int i = 0; i < ME; i++) {
        double s = (i == c) ? 1.0 : 0.0;
        LF_UNROLL for (int t = 0; t < i; t++) if (t >= c) s -= Lc[i][t] * D[t][c];
        D[i][c] = (i < c) ? 0.0 : s / Lc[i][i];
      }
    int i = 0;
    while (i < prm.maxiter_retract) {
      double cm = 0.0;
      LF_UNROLL for (int a = 0; a < ME; a++) cm = pmax(fabs(cval[a]), cm);
      if (cm < prm.eps_c) break;
      LF_UNROLL for (int a = 0; a < ME; a++) { double s = 0.0; LF_UNROLL for (int b = 0; b < ME; b++) s += D[a][b] * cval[b]; t1[a] = -s; }
      double yv[MEA];                              // xnew += Q delta, Q = PJct L^-T
      LF_UNROLL for (int a = 0; a < ME; a++) yv[a] = t1[a];
      LF_UNROLL for (int k = ME - 1; k >= 0; k--) { double s = yv[k]; LF_UNROLL for (int t = k + 1; t < ME; t++) s -= Lc[t][k] * yv[t]; yv[k] = s / Lc[k][k]; }
      LF_UNROLL for (int k = 0; k < NPL; k++) {
        double wj = coldot(yv, k);
        if (INEQ) { xnew.x[k] += Dy[k] * Dy[k] * wj; xnew.y[k] -= Dx[k] * Dy[k] * wj; } else xnew.x[k] += wj;
      }
      if (INEQ) y_retract(xnew, x);
      c_aux(t2, xnew);
      LF_UNROLL for (int a = 0; a < ME; a++) { dcv[a] = t2[a] - cval[a]; cval[a] = t2[a]; }
      LF_UNROLL for (int a = 0; a < ME; a++) { double s = 0.0; LF_UNROLL for (int b = 0; b < ME; b++) s += D[b][a] * t1[b]; t2[a] = s; }
      double den = 0.0;
      LF_UNROLL for (int a = 0; a < ME; a++) den += t2[a] * dcv[a];
      LF_UNROLL for (int a = 0; a < ME; a++) { double s = 0.0; LF_UNROLL for (int b = 0; b < ME; b++) s += D[a][b] * dcv[b]; t1[a] -= s; }
      double al = 1.0 / den;
      LF_UNROLL for (int a = 0; a < ME; a++) LF_UNROLL for (int b = 0; b < ME; b++) D[a][b] += al * t1[a] * t2[b];
      i++;
    }
    *it1 = i;
    return (i == prm.maxiter_retract) ? 1 : 0;
  }

  // ---------------------------------------------------------------- the driver (optimize.jl:176-443)
  LFPSQP_DEV void run(const BatchedArgs &A, int64_t k, const double *bnd) {
    fc.prm = A.fam_params ? A.fam_params + k * A.fam_stride : nullptr;
    st_projcg = st_negcurv = st_trials = st_rout = st_rpcg = st_bt = st_newton = st_fact = st_feval = 0; status = 0;
    rank = ME; pinv = false;
    LF_UNROLL for (int s = 0; s < NPL; s++) {
      int j = idx(s);
      valid[s] = j < NA;
      if (INEQ) {
        bkind[s] = valid[s] ? (int)bnd[j] : 0; bq[s] = valid[s] ? bnd[NA + j] : 0.0; br[s] = valid[s] ? bnd[2 * NA + j] : 0.0;
        bs[s] = valid[s] ? bnd[3 * NA + j] : 0.0; bt[s] = valid[s] ? bnd[4 * NA + j] : 0.0;
      } else { bkind[s] = 0; bq[s] = br[s] = bs[s] = bt[s] = 0.0; }
      LF_UNROLL for (int q = 0; q < KP; q++) ep[s][q] = 0.0;
      if (j < n) Fam::load(fc, j, ep[s]);
      x.x[s] = (j < n) ? A.x0[k * n + j] : 0.0;
      if (INEQ) x.y[s] = 0.0;
      cvh[s] = 0.0; Dx[s] = 0.0; Dy[s] = 0.0; S[s] = 1.0; Sinv[s] = 1.0; lamy[s] = 0.0;
    }
    LF_UNROLL for (int a = 0; a < MEA; a++) { cvc[a] = 0.0; cval[a] = 0.0; lam[a] = 0.0; }
    if (ME > 0 && p > 0) {  // slack start values s0 = d(x0) (optimize.jl:26-28): c_aux with s = 0 gives d(x0) in rows >= m
      double cv0[MEA];
      c_aux(cv0, x);
      LF_UNROLL for (int s = 0; s < NPL; s++) { int j = idx(s); LF_UNROLL for (int a = 0; a < ME; a++) if (a >= m && j == n + (a - m)) x.x[s] = cv0[a]; }
    }
    if (INEQ) generate_initial_y(x);
    int64_t it = 0, nobj = 0;
    double f_diff = INFINITY, step_diff = INFINITY, kkt_diff = INFINITY, prev_grad_norm = 0.0;
    double fval = f_aux(x);
    if (lane == 0 && nobj < A.H) A.obj_hist[k * A.H + nobj] = fval;
    nobj++;
    if (ME > 0) c_aux(cval, x);
    int cond = LFPSQP_F_TOL, last_flag = 0;
    Vec g, d;
    while (true) {
      LF_UNROLL for (int s = 0; s < NPL; s++) {
        int j = idx(s);
        g.x[s] = (j < n) ? Fam::g(fc, j, x.x[s], ep[s]) : 0.0;                // optimize.jl:259
        d.x[s] = -1.0 * g.x[s];
        if (INEQ) { g.y[s] = 0.0; d.y[s] = -0.0; }
      }
      if (INEQ) inequality_gradient(x);
      if (ME > 0) {
        jac_aux(cval, x);
        factor();
        if (pinv) status |= LFPSQP_ST_RANK_DEFICIENT;   // informational: truncated path of optimize.jl:297-302
      }
      project(d, true);
      double km = 0.0;
      LF_UNROLL for (int s = 0; s < NPL; s++) { km = pmax(fabs(d.x[s]), km); if (INEQ) km = pmax(fabs(d.y[s]), km); }
      kkt_diff = wmax(km);                                                     // :320
      if (f_diff <= prm.eps_f) { cond = LFPSQP_F_TOL; break; }
      else if (step_diff <= prm.eps_x) { cond = LFPSQP_X_TOL; break; }
      else if (it >= prm.maxiter) { cond = LFPSQP_MAX_ITER; break; }
      else if (kkt_diff <= prm.eps_kkt) { cond = LFPSQP_KKT_TOL; break; }
      if (!(kkt_diff == kkt_diff)) { status |= LFPSQP_ST_NONFINITE; cond = LFPSQP_MAX_ITER; break; }
      if (prm.do_newton) {
        double gn = sqrt(dot(d, d));
        double tol = prm.tn_kappa * fmin(1.0, gn / prev_grad_norm) * gn;
        prev_grad_norm = gn;
        Vec nd;
        projcg(nd, d, tol, prm.tn_maxiter);
        if (dot(nd, d) > 0.0) { d = nd; st_newton++; }
      }
      int kind;
      if (ME > 0) kind = (rank == ME && !prm.do_project_retract) ? 2 : 3; else kind = INEQ ? 1 : 0;
      // armijo! (linesearch.jl:32-89)
      double alpha = prm.alpha, newf = 0.0;
      f_diff = INFINITY; step_diff = INFINITY;
      const double ar_dot = dot(d, g);
      int flag = 0;
      Vec xnew, xtil;
      while (step_diff > prm.eps_x) {
        LF_UNROLL for (int s = 0; s < NPL; s++) { xtil.x[s] = x.x[s] + alpha * d.x[s]; if (INEQ) xtil.y[s] = x.y[s] + alpha * d.y[s]; }
        int i1 = 0, i2 = 0;
        if (kind == 0) { xnew = xtil; flag = 0; }
        else if (kind == 1) { xnew = xtil; y_retract(xnew, x); flag = 0; }
        else if (kind == 2) flag = retract_nr(xnew, xtil, &i1);
        else flag = retract_pp(xnew, xtil, &i1, &i2);
        st_rout += i1; st_rpcg += i2; st_trials++;
        // linesearch.jl:57-60 has no lower bound on alpha in this branch: when the retraction fails at EVERY alpha the
        // reference spins forever once alpha has underflowed to 0.  Stop at the floor the other branch uses (:82-85):
        // flag 98, LFPSQP_ST_NONFINITE.
        if (flag > 0) { if (alpha < 1e-100) { flag = 98; break; } alpha *= prm.s; continue; }
        newf = f_aux(xnew);
        double s2 = 0.0;
        LF_UNROLL for (int s = 0; s < NPL; s++) { double t = xnew.x[s] - x.x[s]; s2 += t * t; }   // first n_A entries (:66)
        step_diff = sqrt(wsum(s2));
        f_diff = fabs(newf - fval);
        if (prm.disable_linesearch) break;
        if ((newf - fval) <= prm.sigma * alpha * ar_dot) break;
        alpha *= prm.s;
        if (alpha < 1e-100) { flag = 99; break; }
      }
      last_flag = flag;
      if (flag == 98) { status |= LFPSQP_ST_NONFINITE; cond = LFPSQP_MAX_ITER; break; }
      x = xnew;
      fval = newf;
      if (lane == 0 && nobj < A.H) A.obj_hist[k * A.H + nobj] = fval;
      nobj++;
      it++;
    }
    LF_UNROLL for (int s = 0; s < NPL; s++) { int j = idx(s); if (j < n) A.x_out[k * n + j] = x.x[s]; }
    if (lane == 0) {
      LF_UNROLL for (int a = 0; a < ME; a++) A.lambda[k * ME + a] = lam[a];
      A.obj_len[k] = nobj;
      lfpsqp_term t; t.condition = cond; t.status = status; t.f_diff = f_diff; t.step_diff = step_diff;
      t.kkt_diff = kkt_diff; t.iter = it;
      A.term[k] = t;
      if (A.stats) {
        lfpsqp_stats st;
        st.projcg_iters = st_projcg; st.projcg_negcurv = st_negcurv; st.armijo_trials = st_trials; st.retract_outer = st_rout;
        st.retract_pcg = st_rpcg; st.pp_backtracks = st_bt; st.newton_accepted = st_newton; st.factorizations = st_fact;
        st.f_evals = st_feval; st.flag_last = last_flag;
        A.stats[k] = st;
      }
    }
  }
};

#ifndef LFPSQP_REG_MINBLOCKS
#define LFPSQP_REG_MINBLOCKS 3
#endif
template <class Fam, int NPL, int ME, bool INEQ>
__global__ void __launch_bounds__(128, LFPSQP_REG_MINBLOCKS) batched_reg_kernel(const BatchedArgs A) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int NA = A.n + A.p;
  double *bnd = smem;
  const int nb = INEQ ? 5 * NA : 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) bnd[i] = A.bnd[i];
  __syncthreads();
  RegSolver<Fam, NPL, ME, INEQ> S(lane, A.prm, A.n, A.m, A.p);
  for (;;) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(A.work_counter, 1ULL);
    k = __shfl_sync(0xffffffffu, k, 0);
    if ((int64_t)k >= A.B) break;
    S.run(A, (int64_t)k, bnd);
  }
}

}  // namespace lfpsqp
