/*
 * oracle/lfpsqp_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see lfpsqp_oracle.h).
 *
 * CPU restatement of LFPSQP.jl's driver + hot path, following the reference line by line
 * (SVD-based, as the reference; the product replaces SVD by Gram+Cholesky).  Derivatives are
 * analytic (the reference uses ReverseDiff/ForwardDiff, exact to rounding: test/test_autodiff.jl).
 * LAPACK dgesvd is bound at run time from scipy's bundled OpenBLAS (la_helper.jl:22-28 calls the same routine).
 */
#include "lfpsqp_oracle.h"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <thread>
#include <atomic>
#include <vector>

typedef std::vector<double> vec;
static const double INF = std::numeric_limits<double>::infinity();
static const double QNAN = std::numeric_limits<double>::quiet_NaN();

/* caller-supplied noise rows for beta > 0 (optimize.jl:264-273): the reference draws randn!(tmp_n) from Julia's global RNG; here
   the test supplies the same numbers to the oracle and to the device path.  Row i of the current instance at t_noise + i*N. */
static const double *g_noise = nullptr; static int64_t g_noise_T = 0, g_noise_stride = 0;
static thread_local const double *t_noise = nullptr;
extern "C" void orc_set_noise(const double *noise, int64_t T, int64_t stride_per_instance) { g_noise = noise; g_noise_T = T; g_noise_stride = stride_per_instance; t_noise = noise; }
static thread_local double g_flops = 0.0;
static thread_local orc_stats *g_stats = nullptr;

/* ------------------------------------------------------------------ LAPACK binding */
typedef void (*dgesvd_fn)(const char *, const char *, const int *, const int *, double *, const int *, double *,
                          double *, const int *, double *, const int *, double *, const int *, int *);
static dgesvd_fn p_dgesvd = nullptr;
typedef void (*setthreads_fn)(int);

extern "C" int orc_set_lapack(const char *libpath) {
  void *h = dlopen(libpath, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", libpath, dlerror()); return -1; }
  const char *names[] = {"scipy_dgesvd_", "dgesvd_", "dgesvd_64_"};
  for (const char *nm : names) {
    p_dgesvd = (dgesvd_fn)dlsym(h, nm);
    if (p_dgesvd) break;
  }
  if (!p_dgesvd) { fprintf(stderr, "oracle: no dgesvd symbol in %s\n", libpath); return -2; }
  /* single-threaded BLAS: the batched baseline parallelises over instances instead */
  const char *tn[] = {"scipy_openblas_set_num_threads", "openblas_set_num_threads"};
  for (const char *nm : tn) {
    setthreads_fn f = (setthreads_fn)dlsym(h, nm);
    if (f) { f(1); break; }
  }
  return 0;
}

/* ksvd! -- src/la_helper.jl:8-34: thin SVD, jobu=jobvt='S', destroys A. A is M x N col-major (M>=N). */
struct SvdWork {
  vec work;
};
static void ksvd(double *A, int M, int N, double *U, double *S, double *VT, SvdWork &w) {
  if (!p_dgesvd) { fprintf(stderr, "oracle: LAPACK not bound (call orc_set_lapack)\n"); abort(); }
  int info = 0, lda = M, ldu = M, ldvt = N;
  if (w.work.empty()) { /* workspace query, la_helper.jl:17-19,30-33 */
    double q = 0; int lwork = -1;
    p_dgesvd("S", "S", &M, &N, A, &lda, S, U, &ldu, VT, &ldvt, &q, &lwork, &info);
    w.work.resize((size_t)std::max(1.0, q));
  }
  int lwork = (int)w.work.size();
  p_dgesvd("S", "S", &M, &N, A, &lda, S, U, &ldu, VT, &ldvt, w.work.data(), &lwork, &info);
  /* info is ignored by the reference (la_helper.jl:12,28) */
  g_flops += (double)N * (N + 1) * M + (double)N * N * N / 3.0;
  if (g_stats) g_stats->svd_calls++;
}

/* ------------------------------------------------------------------ small BLAS-like helpers (counted) */
static inline double dot(const double *a, const double *b, int64_t n) {
#ifdef ORC_SEQ_SUM
  /* rounding-sensitivity probe (Makefile: liblfpsqp_oracle_seq.so): plain left-to-right summation, as Julia's generic
     fallback would do; the reference does not pin the summation order (BLAS ddot) */
  { double s = 0; for (int64_t i = 0; i < n; i++) s += a[i] * b[i]; g_flops += 2.0 * n; return s; }
#endif
  /* four partial sums, as an optimised BLAS ddot would keep (summation order is not pinned by the reference) */
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int64_t i = 0;
  for (; i + 4 <= n; i += 4) { s0 += a[i] * b[i]; s1 += a[i+1] * b[i+1]; s2 += a[i+2] * b[i+2]; s3 += a[i+3] * b[i+3]; }
  for (; i < n; i++) s0 += a[i] * b[i];
  g_flops += 2.0 * n; return (s0 + s1) + (s2 + s3);
}
static inline double nrm2(const double *a, int64_t n) { return std::sqrt(dot(a, a, n)); }
static inline double nrminf(const double *a, int64_t n) {
  double s = 0; for (int64_t i = 0; i < n; i++) { double v = std::fabs(a[i]); if (v > s || std::isnan(v)) s = v; }
  return s;
}
static inline void axpy(double a, const double *x, double *y, int64_t n) {
  for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
  g_flops += 2.0 * n;
}
/* y = alpha*op(A)*x + beta*y, A is M x N col-major with leading dim lda */
static void gemv(char t, int64_t M, int64_t N, double alpha, const double *A, int64_t lda, const double *x, double beta,
                 double *y) {
  if (t == 'N') {
    for (int64_t i = 0; i < M; i++) y[i] = (beta == 0.0) ? 0.0 : beta * y[i];
    for (int64_t j = 0; j < N; j++) {
      double xj = alpha * x[j]; const double *col = A + j * lda;
      for (int64_t i = 0; i < M; i++) y[i] += col[i] * xj;
    }
  } else {
    for (int64_t j = 0; j < N; j++) {
#ifdef ORC_SEQ_SUM
      { const double *cj = A + j * lda; double s = 0; for (int64_t i = 0; i < M; i++) s += cj[i] * x[i];
        y[j] = alpha * s + ((beta == 0.0) ? 0.0 : beta * y[j]); continue; }
#endif
      const double *col = A + j * lda; double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int64_t i = 0;
      for (; i + 4 <= M; i += 4) { s0 += col[i] * x[i]; s1 += col[i+1] * x[i+1]; s2 += col[i+2] * x[i+2]; s3 += col[i+3] * x[i+3]; }
      for (; i < M; i++) s0 += col[i] * x[i];
      y[j] = alpha * ((s0 + s1) + (s2 + s3)) + ((beta == 0.0) ? 0.0 : beta * y[j]);
    }
  }
  g_flops += 2.0 * M * N;
}

/* ------------------------------------------------------------------ problem families (analytic callbacks) */
struct Family {
  int64_t n = 0, m = 0, p = 0;
  virtual ~Family() {}
  virtual double f(const double *x) = 0;
  virtual void grad(double *g, const double *x) = 0;
  virtual void c(double *cval, const double *x) {}
  /* Jc: m x n col-major with leading dimension ld */
  virtual void jac(double *Jc, int64_t ld, double *cval, const double *x) {}
  virtual void d(double *dval, const double *x) {}
  virtual void jacd(double *Jd, int64_t ld, double *dval, const double *x) {}
  /* dest = (Hess f + sum lam_c[i] Hess c_i + sum lam_d[k] Hess d_k) src */
  virtual void hess(double *dest, const double *src, const double *x, const double *lam_c, const double *lam_d) = 0;
};

struct Rosenbrock : Family { /* README.md:18-22 */
  Rosenbrock() { n = 2; }
  double f(const double *x) override {
    double a = 1 - x[0], b = x[1] - x[0] * x[0]; g_flops += 8; return a * a + 100 * b * b;
  }
  void grad(double *g, const double *x) override {
    double b = x[1] - x[0] * x[0];
    g[0] = -2 * (1 - x[0]) - 400 * x[0] * b; g[1] = 200 * b; g_flops += 10;
  }
  void hess(double *dest, const double *src, const double *x, const double *, const double *) override {
    double h11 = 2 - 400 * (x[1] - x[0] * x[0]) + 800 * x[0] * x[0], h12 = -400 * x[0], h22 = 200;
    double v0 = src[0], v1 = src[1];
    dest[0] = h11 * v0 + h12 * v1; dest[1] = h12 * v0 + h22 * v1; g_flops += 14;
  }
};

struct ReadmeEq : Family { /* README.md:41-54 */
  ReadmeEq(int64_t n_) { n = n_; m = 1; }
  double f(const double *x) override { return dot(x, x, n); }
  void grad(double *g, const double *x) override { for (int64_t i = 0; i < n; i++) g[i] = 2 * x[i]; }
  void c(double *cv, const double *x) override { cv[0] = x[0] - 0.75; }
  void jac(double *J, int64_t ld, double *cv, const double *x) override {
    for (int64_t j = 0; j < n; j++) J[j * ld] = 0; J[0] = 1; cv[0] = x[0] - 0.75;
  }
  void hess(double *dest, const double *src, const double *, const double *, const double *) override {
    for (int64_t i = 0; i < n; i++) dest[i] = 2 * src[i];
  }
};

struct ReadmeIneq : Family { /* README.md:57-76 */
  const double *coeff;
  ReadmeIneq(int64_t n_, const double *co) : coeff(co) { n = n_; m = 0; p = 1; }
  double f(const double *x) override { return dot(coeff, x, n); }
  void grad(double *g, const double *) override { for (int64_t i = 0; i < n; i++) g[i] = coeff[i]; }
  void d(double *dv, const double *x) override { dv[0] = dot(x, x, n) - 1.0; }
  void jacd(double *J, int64_t ld, double *dv, const double *x) override {
    for (int64_t j = 0; j < n; j++) J[j * ld] = 2 * x[j]; dv[0] = dot(x, x, n) - 1.0;
  }
  void hess(double *dest, const double *src, const double *, const double *, const double *lam_d) override {
    for (int64_t i = 0; i < n; i++) dest[i] = 2 * lam_d[0] * src[i]; g_flops += 2.0 * n;
  }
};

struct Thomson : Family { /* SURVEY 8(d) C4 */
  int64_t N;
  Thomson(int64_t n_) { n = n_; N = n_ / 3; m = N; }
  double f(const double *x) override {
    double s = 0;
    for (int64_t i = 0; i < N; i++) for (int64_t j = i + 1; j < N; j++) {
      double a = x[3*i]-x[3*j], b = x[3*i+1]-x[3*j+1], cc = x[3*i+2]-x[3*j+2];
      s += 1.0 / std::sqrt(a*a + b*b + cc*cc);
    }
    g_flops += 5.0 * N * N; return s;
  }
  void grad(double *g, const double *x) override {
    for (int64_t i = 0; i < n; i++) g[i] = 0;
    for (int64_t i = 0; i < N; i++) for (int64_t j = i + 1; j < N; j++) {
      double a = x[3*i]-x[3*j], b = x[3*i+1]-x[3*j+1], cc = x[3*i+2]-x[3*j+2];
      double r2 = a*a + b*b + cc*cc, r = std::sqrt(r2), ir3 = 1.0 / (r2 * r);
      g[3*i] -= a*ir3; g[3*i+1] -= b*ir3; g[3*i+2] -= cc*ir3;
      g[3*j] += a*ir3; g[3*j+1] += b*ir3; g[3*j+2] += cc*ir3;
    }
    g_flops += 11.0 * N * N;
  }
  void c(double *cv, const double *x) override {
    for (int64_t i = 0; i < N; i++) cv[i] = x[3*i]*x[3*i] + x[3*i+1]*x[3*i+1] + x[3*i+2]*x[3*i+2] - 1.0;
  }
  void jac(double *J, int64_t ld, double *cv, const double *x) override {
    for (int64_t j = 0; j < n; j++) for (int64_t i = 0; i < m; i++) J[i + j * ld] = 0;
    for (int64_t i = 0; i < N; i++) for (int k = 0; k < 3; k++) J[i + (3*i+k) * ld] = 2 * x[3*i+k];
    c(cv, x);
  }
  void hess(double *dest, const double *v, const double *x, const double *lam, const double *) override {
    for (int64_t i = 0; i < n; i++) dest[i] = 0;
    for (int64_t i = 0; i < N; i++) for (int64_t j = i + 1; j < N; j++) {
      double a = x[3*i]-x[3*j], b = x[3*i+1]-x[3*j+1], cc = x[3*i+2]-x[3*j+2];
      double wa = v[3*i]-v[3*j], wb = v[3*i+1]-v[3*j+1], wc = v[3*i+2]-v[3*j+2];
      double r2 = a*a + b*b + cc*cc, r = std::sqrt(r2), ir3 = 1.0 / (r2 * r), ir5 = ir3 / r2;
      double rw = 3.0 * (a*wa + b*wb + cc*wc) * ir5;
      double ha = rw*a - wa*ir3, hb = rw*b - wb*ir3, hc = rw*cc - wc*ir3;
      dest[3*i] += ha; dest[3*i+1] += hb; dest[3*i+2] += hc;
      dest[3*j] -= ha; dest[3*j+1] -= hb; dest[3*j+2] -= hc;
    }
    for (int64_t i = 0; i < N; i++) for (int k = 0; k < 3; k++) dest[3*i+k] += 2 * lam[i] * v[3*i+k];
    g_flops += 17.0 * N * N;
  }
};

struct DiagQuad : Family { /* c_i = 0.5 sum_j Q_ij x_j^2 + A_i.x - b_i ; f = 0.5 sum w_j (x_j-xt_j)^2 */
  const double *Q, *A, *b, *xt, *w;
  DiagQuad(int64_t n_, int64_t m_, const double *prm) {
    n = n_; m = m_; Q = prm; A = Q + m * n; b = A + m * n; xt = b + m; w = xt + n;
  }
  double f(const double *x) override {
    double s = 0; for (int64_t j = 0; j < n; j++) { double t = x[j] - xt[j]; s += w[j] * t * t; }
    g_flops += 4.0 * n; return 0.5 * s;
  }
  void grad(double *g, const double *x) override { for (int64_t j = 0; j < n; j++) g[j] = w[j] * (x[j] - xt[j]); }
  void c(double *cv, const double *x) override {
    for (int64_t i = 0; i < m; i++) {
      const double *q = Q + i * n, *a = A + i * n; double s = 0;
      for (int64_t j = 0; j < n; j++) s += (0.5 * q[j] * x[j] + a[j]) * x[j];
      cv[i] = s - b[i];
    }
    g_flops += 4.0 * m * n;
  }
  void jac(double *J, int64_t ld, double *cv, const double *x) override {
    for (int64_t i = 0; i < m; i++) {
      const double *q = Q + i * n, *a = A + i * n;
      for (int64_t j = 0; j < n; j++) J[i + j * ld] = q[j] * x[j] + a[j];
    }
    c(cv, x); g_flops += 2.0 * m * n;
  }
  void hess(double *dest, const double *v, const double *, const double *lam, const double *) override {
    for (int64_t j = 0; j < n; j++) dest[j] = 0;
    for (int64_t i = 0; i < m; i++) { const double *q = Q + i * n; for (int64_t j = 0; j < n; j++) dest[j] += lam[i] * q[j]; }
    for (int64_t j = 0; j < n; j++) dest[j] = (w[j] + dest[j]) * v[j];
    g_flops += 2.0 * m * n + 2.0 * n;
  }
};

struct SinSystem : Family { /* test/test_retractions.jl:34-54 ; f = 0.5|x-t|^2 */
  const double *t;
  SinSystem(int64_t n_, int64_t m_, const double *prm) : t(prm) { n = n_; m = m_; }
  double f(const double *x) override {
    double s = 0; for (int64_t j = 0; j < n; j++) { double u = x[j] - t[j]; s += u * u; } return 0.5 * s;
  }
  void grad(double *g, const double *x) override { for (int64_t j = 0; j < n; j++) g[j] = x[j] - t[j]; }
  void c(double *cv, const double *x) override { for (int64_t i = 0; i < m; i++) cv[i] = x[2*i+1] - std::sin(x[2*i]); }
  void jac(double *J, int64_t ld, double *cv, const double *x) override {
    for (int64_t j = 0; j < n; j++) for (int64_t i = 0; i < m; i++) J[i + j * ld] = 0;
    for (int64_t i = 0; i < m; i++) { cv[i] = x[2*i+1] - std::sin(x[2*i]); J[i + (2*i+1) * ld] = 1.0; J[i + (2*i) * ld] = -std::cos(x[2*i]); }
  }
  void hess(double *dest, const double *v, const double *x, const double *lam, const double *) override {
    for (int64_t j = 0; j < n; j++) dest[j] = v[j];
    for (int64_t i = 0; i < m; i++) dest[2*i] += lam[i] * std::sin(x[2*i]) * v[2*i];
  }
};

struct BoxQuad : Family { /* f = |x-t|^2, optional c = a.x - b */
  const double *t, *a; double b;
  BoxQuad(int64_t n_, int64_t m_, const double *prm) { n = n_; m = m_; t = prm; a = prm + n; b = prm[2 * n]; }
  double f(const double *x) override { double s = 0; for (int64_t j = 0; j < n; j++) { double u = x[j] - t[j]; s += u * u; } return s; }
  void grad(double *g, const double *x) override { for (int64_t j = 0; j < n; j++) g[j] = 2 * (x[j] - t[j]); }
  void c(double *cv, const double *x) override { if (m) cv[0] = dot(a, x, n) - b; }
  void jac(double *J, int64_t ld, double *cv, const double *x) override {
    if (!m) return; for (int64_t j = 0; j < n; j++) J[j * ld] = a[j]; cv[0] = dot(a, x, n) - b;
  }
  void hess(double *dest, const double *v, const double *, const double *, const double *) override {
    for (int64_t j = 0; j < n; j++) dest[j] = 2 * v[j];
  }
};

static std::unique_ptr<Family> make_family(int family, const double *prm, int64_t n, int64_t m, int64_t p) {
  std::unique_ptr<Family> F;
  switch (family) {
    case ORC_FAM_ROSENBROCK: if (n != 2 || m || p) return F; F.reset(new Rosenbrock()); break;
    case ORC_FAM_README_EQ: if (m != 1 || p) return F; F.reset(new ReadmeEq(n)); break;
    case ORC_FAM_README_INEQ: if (m || p != 1) return F; F.reset(new ReadmeIneq(n, prm)); break;
    case ORC_FAM_THOMSON: if (n % 3 || m != n / 3 || p) return F; F.reset(new Thomson(n)); break;
    case ORC_FAM_DIAGQUAD: if (p) return F; F.reset(new DiagQuad(n, m, prm)); break;
    case ORC_FAM_SIN: if (p || 2 * m > n) return F; F.reset(new SinSystem(n, m, prm)); break;
    case ORC_FAM_BOXQUAD: if (p || m > 1) return F; F.reset(new BoxQuad(n, m, prm)); break;
    default: break;
  }
  return F;
}

/* The explicit-derivative problem the core driver sees (optimize.jl:119). */
struct Problem {
  int64_t n = 0, m = 0;
  virtual ~Problem() {}
  virtual double f(const double *x) = 0;
  virtual void grad(double *g, const double *x) = 0;
  virtual void c(double *cval, const double *x) = 0;
  virtual void jac(double *Jc, double *cval, const double *x) = 0; /* m x n col-major, also writes cval */
  virtual void hess_lag_vec(double *dest, const double *src, const double *x, const double *lam) = 0;
};

/* optimize.jl:88-104: equality + bounds, no general inequalities */
struct PlainProblem : Problem {
  Family &F;
  PlainProblem(Family &F_) : F(F_) { n = F.n; m = F.m; }
  double f(const double *x) override { if (g_stats) g_stats->f_evals++; return F.f(x); }
  void grad(double *g, const double *x) override { F.grad(g, x); }
  void c(double *cv, const double *x) override { F.c(cv, x); }
  void jac(double *J, double *cv, const double *x) override { F.jac(J, m, cv, x); }
  void hess_lag_vec(double *dest, const double *src, const double *x, const double *lam) override {
    F.hess(dest, src, x, lam, nullptr);
  }
};

/* optimize.jl:13-71: slack-variable wrapper. x_aux=[x;s], c_aux=[c(x); d(x)-s] */
struct SlackProblem : Problem {
  Family &F; int64_t n0, m0, p;
  SlackProblem(Family &F_) : F(F_) { n0 = F.n; m0 = F.m; p = F.p; n = n0 + p; m = m0 + p; }
  double f(const double *x) override { if (g_stats) g_stats->f_evals++; return F.f(x); }              /* :38-40 */
  void grad(double *g, const double *x) override { F.grad(g, x); for (int64_t k = 0; k < p; k++) g[n0 + k] = 0; }
  void c(double *cv, const double *x) override {                                                       /* :42-51 */
    if (m0 > 0) F.c(cv, x);
    F.d(cv + m0, x);
    for (int64_t k = 0; k < p; k++) cv[m0 + k] -= x[n0 + k];
  }
  void jac(double *J, double *cv, const double *x) override {
    for (int64_t j = 0; j < n; j++) for (int64_t i = 0; i < m; i++) J[i + j * m] = 0;
    if (m0 > 0) F.jac(J, m, cv, x);
    F.jacd(J + m0, m, cv + m0, x);
    for (int64_t k = 0; k < p; k++) { J[(m0 + k) + (n0 + k) * m] = -1.0; cv[m0 + k] -= x[n0 + k]; }
  }
  void hess_lag_vec(double *dest, const double *src, const double *x, const double *lam) override {
    F.hess(dest, src, x, lam, lam + m0);
    for (int64_t k = 0; k < p; k++) dest[n0 + k] = 0;
  }
};

/* ------------------------------------------------------------------ bound embedding: src/inequality_helper.jl */
struct IneqData { /* :1-8 */
  vec q, r, s, t; std::vector<char> isline, isparabola; int64_t n = 0;
  IneqData() {}
  IneqData(const double *xl, const double *xu, int64_t n_) : q(n_), r(n_), s(n_), t(n_), isline(n_, 0), isparabola(n_, 0), n(n_) {
    for (int64_t i = 0; i < n; i++) { /* :54-82 */
      bool linf = std::isinf(xl[i]), uinf = std::isinf(xu[i]);
      if (linf && uinf) { q[i] = 0; r[i] = 0; s[i] = 0; t[i] = 0; isline[i] = 1; }
      else if (!linf && uinf) { q[i] = 0; r[i] = xl[i]; s[i] = -1.0; t[i] = xl[i]; isparabola[i] = 1; }
      else if (linf && !uinf) { q[i] = 0; r[i] = xu[i]; s[i] = 1.0; t[i] = xu[i]; isparabola[i] = 1; }
      else { q[i] = 1.0; r[i] = (xu[i] + xl[i]) / 2; s[i] = 1.0; t[i] = (xu[i] - xl[i]) * (xu[i] - xl[i]) / 4; }
    }
  }
};

struct IneqDecomp { /* :10-19 (U, Sigma, Vt are views of the driver's arrays) */
  double *U = nullptr, *Sig = nullptr, *Vt = nullptr;
  vec Dx, Dy, S; double *Jct = nullptr; int64_t rank = 0, n = 0, m = 0;
};

static void generate_initial_y(double *xaug, const IneqData &id) { /* :92-109 */
  int64_t n = id.n; double *x = xaug, *y = xaug + n;
  for (int64_t i = 0; i < n; i++) {
    if (id.isline[i]) y[i] = x[i];
    else if (id.isparabola[i]) y[i] = std::sqrt(std::max(-(x[i] - id.t[i]) / id.s[i], 0.0)) + id.r[i];
    else y[i] = std::sqrt(std::max(id.t[i] - (x[i] - id.r[i]) * (x[i] - id.r[i]), 0.0)) + id.r[i];
  }
}
static void calculate_h(double *cvalaug, const double *xaug, const IneqData &id) { /* :112-122 */
  int64_t n = id.n; const double *x = xaug, *y = xaug + n;
  for (int64_t i = 0; i < n; i++) {
    double dx = x[i] - id.r[i], dy = y[i] - id.r[i];
    cvalaug[i] = id.q[i] * (dx * dx) + (1.0 - id.q[i] * id.q[i]) * x[i] + id.s[i] * (dy * dy) - (1.0 - id.s[i] * id.s[i]) * y[i] - id.t[i];
  }
  g_flops += 12.0 * n;
}
static void inequality_gradient(IneqDecomp &de, const double *xaug, const IneqData &id) { /* :125-141 */
  int64_t n = id.n;
  for (int64_t i = 0; i < n; i++) {
    double dx = 2.0 * id.q[i] * (xaug[i] - id.r[i]) + (id.q[i] == 0.0 ? 1.0 : 0.0);
    double dy = 2.0 * id.s[i] * (xaug[n + i] - id.r[i]) - (id.s[i] == 0.0 ? 1.0 : 0.0);
    double S = std::sqrt(dx * dx + dy * dy);
    de.S[i] = S; de.Dx[i] = dx / S; de.Dy[i] = dy / S;
  }
  g_flops += 12.0 * n;
}
/* Q*v : dest(2n) = U[:,1:rank] v[n+1:n+rank] + [Dx;Dy] v[1:n]   (:161-176) ; 5-arg form (:179-194) */
static void ineq_project_mul(double *dest, const IneqDecomp &de, const double *v, double a, double b) {
  int64_t n = de.n, N = 2 * n;
  gemv('N', N, de.rank, a, de.U, N, v + n, b, dest);
  for (int64_t i = 0; i < n; i++) { dest[i] += a * de.Dx[i] * v[i]; dest[n + i] += a * de.Dy[i] * v[i]; }
  g_flops += 6.0 * n;
}
/* Q'*w : dest[1:n] = Dx.w_x + Dy.w_y ; dest[n+1:n+rank] = U' w  (:197-212) */
static void ineq_project_mulT(double *dest, const IneqDecomp &de, const double *w) {
  int64_t n = de.n, N = 2 * n;
  for (int64_t i = 0; i < n; i++) dest[i] = de.Dx[i] * w[i] + de.Dy[i] * w[n + i];
  gemv('T', N, de.rank, 1.0, de.U, N, w, 0.0, dest + n);
  g_flops += 3.0 * n;
}
/* bigA*v (2n) : (:215-251) dest = a*bigA*v + b*dest */
static void ineq_bigA_mul(double *dest, const IneqDecomp &de, const double *v, double a, double b) {
  int64_t n = de.n, m = de.m;
  gemv('N', n, m, a, de.Jct, n, v + n, b, dest);
  for (int64_t i = 0; i < n; i++) {
    dest[i] += a * de.Dx[i] * de.S[i] * v[i];
    dest[n + i] = ((b == 0.0) ? 0.0 : b * dest[n + i]) + a * de.Dy[i] * de.S[i] * v[i];
  }
  g_flops += 8.0 * n;
}
/* bigA'*w (n+m) : (:254-271) */
static void ineq_bigA_mulT(double *dest, const IneqDecomp &de, const double *w) {
  int64_t n = de.n, m = de.m;
  for (int64_t i = 0; i < n; i++) dest[i] = de.S[i] * de.Dx[i] * w[i] + de.S[i] * de.Dy[i] * w[n + i];
  gemv('T', n, m, 1.0, de.Jct, n, w, 0.0, dest + n);
  g_flops += 6.0 * n;
}
/* calculate_lambda_kkt! (:286-308) */
static void calculate_lambda_kkt(double *lam, double *lamy, double *QtF, const IneqDecomp &de) {
  int64_t n = de.n, m = de.m, rank = de.rank;
  for (int64_t j = 0; j < rank; j++) QtF[n + j] /= de.Sig[j];
  for (int64_t j = rank; j < m; j++) QtF[n + j] = 0.0;
  gemv('T', m, m, 1.0, de.Vt, m, QtF + n, 0.0, lam);  /* lam = Vt' * QtF[n+1:n+m] */
  gemv('N', n, m, 1.0, de.Jct, n, lam, 0.0, lamy);
  for (int64_t i = 0; i < n; i++) { lamy[i] *= -1.0 * de.Dx[i] / de.S[i]; lamy[i] += QtF[i] / de.S[i]; }
  g_flops += 4.0 * n;
}
/* augmented_hess_lag_vec! (:144-158) */
static void augmented_hess_lag_vec(double *dest, const double *src, Problem &P, const double *x, const double *lam,
                                   const double *lamy, const IneqData &id) {
  int64_t n = id.n;
  P.hess_lag_vec(dest, src, x, lam);
  for (int64_t i = 0; i < n; i++) {
    dest[i] += 2 * lamy[i] * id.q[i] * src[i];
    dest[n + i] = 2 * lamy[i] * id.s[i] * src[n + i];
  }
  g_flops += 7.0 * n;
}
/* y_retract! -- src/retractions.jl:451-500 */
static void y_retract(double *xnewaug, const double *xaug, const IneqData &id) {
  int64_t n = id.n; double *xnew = xnewaug, *ynew = xnewaug + n; const double *x = xaug, *y = xaug + n;
  for (int64_t i = 0; i < n; i++) {
    if (id.isline[i]) { xnew[i] = ynew[i]; }
    else if (id.isparabola[i]) {
      double s = id.s[i], r = id.r[i];
      double g1 = -s, g2 = -2 * (y[i] - r);
      double ng = std::sqrt(g1 * g1 + g2 * g2);
      double ux = x[i] - xnew[i] + g1 / ng, uy = y[i] - ynew[i] + g2 / ng;
      double a = s * uy * uy, b = ux + 2 * s * (ynew[i] - r) * uy, c = xnew[i] + s * (ynew[i] - r) * (ynew[i] - r) - r;
      double a1 = -b / (2 * a), a2 = std::sqrt(b * b - 4 * a * c) / (2 * a);
      double gam = std::min(a1 + a2, a1 - a2);
      xnew[i] += gam * ux; ynew[i] += gam * uy;
    } else {
      double c = id.r[i], rho = std::sqrt(id.t[i]);
      double dist = std::sqrt((xnew[i] - c) * (xnew[i] - c) + (ynew[i] - c) * (ynew[i] - c));
      ynew[i] = c + rho * (ynew[i] - c) / dist;
      xnew[i] = c + rho * (xnew[i] - c) / dist;
    }
  }
  g_flops += 30.0 * n;
}

/* ------------------------------------------------------------------ projcg! -- src/projcg.jl:40-121 */
struct Projector { /* the "U" argument: either view(U,:,1:rank) or InequalityDecompProject */
  virtual ~Projector() {}
  virtual void mul(double *dest, const double *v, double a, double b) = 0; /* dest = a U v + b dest */
  virtual void mulT(double *dest, const double *w) = 0;                    /* dest = U' w */
};
struct DenseProjector : Projector {
  const double *U; int64_t N, rank;
  DenseProjector(const double *U_, int64_t N_, int64_t r_) : U(U_), N(N_), rank(r_) {}
  void mul(double *dest, const double *v, double a, double b) override { gemv('N', N, rank, a, U, N, v, b, dest); }
  void mulT(double *dest, const double *w) override { gemv('T', N, rank, 1.0, U, N, w, 0.0, dest); }
};
struct IneqProjector : Projector {
  const IneqDecomp &de;
  IneqProjector(const IneqDecomp &d) : de(d) {}
  void mul(double *dest, const double *v, double a, double b) override { ineq_project_mul(dest, de, v, a, b); }
  void mulT(double *dest, const double *w) override { ineq_project_mulT(dest, de, w); }
};
typedef std::function<void(double *, const double *)> LinOp;

struct ProjCGWork { vec r, g, d, rp, gp, Ad, Utr; ProjCGWork(int64_t n, int64_t m) : r(n), g(n), d(n), rp(n), gp(n), Ad(n), Utr(m) {} };

static void projcg(double *x, double *lam, const LinOp &A, Projector &U, const double *b, const double *c, int64_t n,
                   int64_t m, double tol, int64_t maxit, ProjCGWork &w, int64_t *iters, double *nr_out) {
  double *r = w.r.data(), *g = w.g.data(), *d = w.d.data(), *rp = w.rp.data(), *gp = w.gp.data(), *Ad = w.Ad.data(),
         *Utr = w.Utr.data();
  U.mul(x, c, 1.0, 0.0);                                   /* :55 x = U c */
  A(r, x); for (int64_t i = 0; i < n; i++) r[i] = r[i] - b[i]; /* :56-57 r = A x - b */
  std::copy(r, r + n, g);
  U.mulT(Utr, r); U.mul(g, Utr, -1.0, 1.0);                /* :59-60 */
  std::copy(g, g + n, r);
  for (int64_t i = 0; i < n; i++) d[i] = -1.0 * g[i];
  int64_t i = 0; double nr = INF;
  int64_t lim = std::min(maxit, n + m);
  while (i < lim) {                                        /* :71 */
    i++;
    A(Ad, d);
    double dAd = dot(d, Ad, n);
    if (dAd <= 0) {                                        /* :77-82 */
      double nd = nrm2(d, n);
      for (int64_t k = 0; k < n; k++) x[k] = d[k] / nd;
      for (int64_t k = 0; k < m; k++) lam[k] = QNAN;
      if (g_stats) { g_stats->projcg_iters += i; g_stats->projcg_negcurv++; }
      *iters = i; *nr_out = INF; return;
    }
    double rg = dot(r, g, n);
    if (rg <= 0) break;                                    /* :87-89 */
    double alpha = rg / dAd;
    for (int64_t k = 0; k < n; k++) { x[k] += alpha * d[k]; rp[k] = r[k] + alpha * Ad[k]; }
    std::copy(rp, rp + n, gp);
    U.mulT(Utr, rp); U.mul(gp, Utr, -1.0, 1.0);            /* :95-97 */
    double beta = dot(rp, gp, n) / rg;
    for (int64_t k = 0; k < n; k++) { d[k] = beta * d[k] - gp[k]; g[k] = gp[k]; r[k] = gp[k]; }
    g_flops += 7.0 * n;
    nr = nrm2(g, n);
    if (nr < tol) break;                                   /* :107-111 */
  }
  A(r, x); for (int64_t k = 0; k < n; k++) r[k] = b[k] - r[k]; /* :115-116 */
  U.mulT(lam, r);                                          /* :118 */
  if (g_stats) g_stats->projcg_iters += i;
  *iters = i; *nr_out = nr;
}

/* ------------------------------------------------------------------ retractions -- src/retractions.jl */
/* "fulljac" of pcg!/ProjPenalty: either J (m x n col-major) or idecomp' (bigA') */
struct FullJac {
  bool ineq; const double *J; int64_t m, n; const IneqDecomp *de;
  int64_t rows() const { return ineq ? de->n + de->m : m; }
  int64_t cols() const { return ineq ? 2 * de->n : n; }
  void mul(double *dest, const double *v) const { /* dest = fulljac * v */
    if (ineq) ineq_bigA_mulT(dest, *de, v); else gemv('N', m, n, 1.0, J, m, v, 0.0, dest);
  }
  void mulT(double *dest, const double *w, double a, double b) const { /* dest = a fulljac' w + b dest */
    if (ineq) ineq_bigA_mul(dest, *de, w, a, b); else gemv('T', m, n, a, J, m, w, b, dest);
  }
};

/* pcg! :179-246 with no_precondition (:259-263) */
static int pcg(double mu, const FullJac &J, double *x, double *r, double *p, double *z, double *tmp_m, double tol,
               int64_t maxiter, int64_t *iters) {
  int64_t n = J.cols();
  double norm_res = INF, rho = 1.0;
  std::fill(p, p + n, 0.0);
  int64_t i = 0;
  while (norm_res > tol && i < maxiter) {
    std::copy(r, r + n, z);                       /* M!(z, r) */
    double rho_prev = rho; rho = dot(z, r, n);
    double beta = rho / rho_prev;
    for (int64_t k = 0; k < n; k++) p[k] = z[k] + beta * p[k];
    std::copy(p, p + n, z);
    J.mul(tmp_m, p);
    J.mulT(z, tmp_m, 1.0, mu);                    /* z = J'(J p) + mu p */
    double alpha = rho / dot(p, z, n);
    axpy(alpha, p, x, n); axpy(-alpha, z, r, n);
    norm_res = nrm2(r, n);
    i++;
  }
  *iters = i;
  return (i == maxiter) ? 1 : 0;
}

struct NRWork { vec D, tmp_m, tmp_m2, dc; NRWork(int64_t m) : D(m * m), tmp_m(m), tmp_m2(m), dc(m) {} };
struct PPWork {
  vec J, tmp_m, r, p, z, dx, g, cvalaug;
  PPWork(int64_t m, int64_t n, int64_t m_ineq, int64_t n_ineq)
      : J(m * n), tmp_m(m_ineq), r(n_ineq), p(n_ineq), z(n_ineq), dx(n_ineq), g(n_ineq), cvalaug(m_ineq, 0.0) {}
};

struct Retractor {
  enum Kind { EUCLIDEAN, YRETRACT, NR, PP } kind = EUCLIDEAN;
  Problem *P = nullptr; int64_t n = 0, m = 0, N = 0; /* n = #x variables, N = n or 2n */
  bool ineq = false; const IneqData *idata = nullptr; IneqDecomp *idecomp = nullptr;
  const double *U = nullptr, *Sig = nullptr, *Vt = nullptr;
  double tol = 1e-6, mu0 = 1e-2; int64_t maxiter = 100, maxiter_pcg = 100;
  NRWork *nrw = nullptr; PPWork *ppw = nullptr;
};

/* retract!(::NR) :75-177 */
static void retract_nr(double *cval, double *xnew, const double *xtilde, const double *x, Retractor &R, int *flag,
                       int64_t *it1, int64_t *it2) {
  int64_t m = R.m, N = R.N;
  double *D = R.nrw->D.data(), *tmp_m = R.nrw->tmp_m.data(), *tmp_m2 = R.nrw->tmp_m2.data(), *dc = R.nrw->dc.data();
  std::copy(xtilde, xtilde + N, xnew);
  if (R.ineq) y_retract(xnew, x, *R.idata);
  R.P->c(cval, xnew);
  for (int64_t j = 0; j < m; j++) for (int64_t k = 0; k < m; k++) D[k + j * m] = R.Vt[k + j * m] / R.Sig[k]; /* :126-130 */
  int64_t i = 0;
  while (i < R.maxiter) {
    if (nrminf(cval, m) < R.tol) break;
    gemv('N', m, m, -1.0, D, m, cval, 0.0, tmp_m);      /* :140 */
    gemv('N', N, m, 1.0, R.U, N, tmp_m, 1.0, xnew);     /* :141 */
    if (R.ineq) y_retract(xnew, x, *R.idata);
    R.P->c(tmp_m2, xnew);
    for (int64_t k = 0; k < m; k++) { dc[k] = tmp_m2[k] - cval[k]; cval[k] = tmp_m2[k]; }
    gemv('T', m, m, 1.0, D, m, tmp_m, 0.0, tmp_m2);     /* :156 D' dx */
    gemv('N', m, m, -1.0, D, m, dc, 1.0, tmp_m);        /* :157 tmp_m = dx - D dc */
    double alpha = 1.0 / dot(tmp_m2, dc, m);
    for (int64_t j = 0; j < m; j++) for (int64_t k = 0; k < m; k++) D[k + j * m] += alpha * tmp_m[k] * tmp_m2[j]; /* ger! :160 */
    g_flops += 2.0 * m * m;
    i++;
  }
  *flag = (i == R.maxiter) ? 1 : 0; *it1 = i; *it2 = 0;
}

/* retract!(::ProjPenalty) :265-441 */
static void retract_pp(double *cval, double *xnew, const double *xtilde, const double *x, Retractor &R, int *flag_out,
                       int64_t *it1, int64_t *it2) {
  (void)x;
  PPWork &w = *R.ppw;
  double *J = w.J.data(), *tmp_m = w.tmp_m.data(), *r = w.r.data(), *p = w.p.data(), *z = w.z.data(), *dx = w.dx.data(),
         *g = w.g.data(), *cvalaug = w.cvalaug.data();
  int64_t n = R.n, m = R.m, N = R.N, Maug = (int64_t)w.cvalaug.size();
  IneqDecomp *de = R.idecomp;
  FullJac fj; fj.ineq = R.ineq; fj.J = J; fj.m = m; fj.n = n; fj.de = de;
  int flag = 0;
  std::copy(xtilde, xtilde + N, xnew);
  double mu = R.mu0;
  int64_t i = 0, pcg_iter_count = 0;
  while (i < R.maxiter) {
    R.P->jac(J, cval, xnew);                                 /* :340 */
    double curtol = nrminf(cval, m);
    if (R.ineq) {
      inequality_gradient(*de, xnew, *R.idata);              /* :344 */
      for (int64_t a = 0; a < m; a++) for (int64_t b = 0; b < n; b++) de->Jct[b + a * n] = J[a + b * m]; /* :347 */
      calculate_h(cvalaug, xnew, *R.idata);                  /* :350 */
      curtol = std::max(curtol, nrminf(cvalaug, Maug));      /* :352 (includes the stale tail of cvalaug) */
    }
    for (int64_t a = 0; a < m; a++) cvalaug[Maug - m + a] = cval[a]; /* :356 */
    if (curtol < R.tol) break;                               /* :359-361 */
    for (int64_t k = 0; k < N; k++) g[k] = xnew[k] - xtilde[k];
    double prev_obj_val = dot(cvalaug, cvalaug, Maug) + mu * dot(g, g, N); /* :366 */
    fj.mulT(g, cvalaug, 1.0, mu);                            /* :369 */
    std::fill(dx, dx + N, 0.0);
    std::copy(g, g + N, r);
    int64_t pcg_i = 0;
    int pcg_flag = pcg(mu, fj, dx, r, p, z, tmp_m, R.tol, R.maxiter_pcg, &pcg_i); /* :375 */
    pcg_iter_count += pcg_i;
    if (pcg_flag > 0) { flag = 2; break; }                   /* :377-381 */
    std::copy(xnew, xnew + N, p);                            /* :384 */
    double ar_dot = -dot(g, dx, N);
    double alpha = 1.0;
    for (int64_t k = 0; k < N; k++) { xnew[k] -= alpha * dx[k]; g[k] = xnew[k] - xtilde[k]; }
    double dist2 = dot(g, g, N);
    R.P->c(cval, xnew);                                      /* :392 */
    if (R.ineq) calculate_h(cvalaug, xnew, *R.idata);
    for (int64_t a = 0; a < m; a++) cvalaug[Maug - m + a] = cval[a];
    int armijo_count = 0;
    while (dot(cvalaug, cvalaug, Maug) + mu * dist2 > prev_obj_val + 1e-4 * alpha * ar_dot) { /* :403 */
      alpha /= 2;
      for (int64_t k = 0; k < N; k++) { xnew[k] = p[k] - alpha * dx[k]; g[k] = xnew[k] - xtilde[k]; }
      dist2 = dot(g, g, N);
      R.P->c(cvalaug, xnew);                                 /* :410 c! written into cvalaug[1:m] ... */
      if (R.ineq) calculate_h(cvalaug, xnew, *R.idata);      /* :413-415 */
      for (int64_t a = 0; a < m; a++) cvalaug[Maug - m + a] = cval[a]; /* :417 ... and overwritten by the stale cval */
      armijo_count++;
      if (g_stats) g_stats->pp_backtracks++;
      if (armijo_count == 100) { flag = 3; break; }
    }
    i++;
    mu = std::min(mu * 0.1, nrm2(cvalaug, Maug));            /* :431 */
  }
  if (i == R.maxiter) flag = 1;
  *flag_out = flag; *it1 = i; *it2 = pcg_iter_count;
}

static void retract(double *cval, double *xnew, const double *xtilde, const double *x, Retractor &R, int *flag,
                    int64_t *it1, int64_t *it2) {
  switch (R.kind) {
    case Retractor::EUCLIDEAN: std::copy(xtilde, xtilde + R.N, xnew); *flag = 0; *it1 = 0; *it2 = 0; break; /* :61-65 */
    case Retractor::YRETRACT: std::copy(xtilde, xtilde + R.N, xnew); y_retract(xnew, x, *R.idata); *flag = 0; *it1 = 0; *it2 = 0; break; /* :67-72 */
    case Retractor::NR: retract_nr(cval, xnew, xtilde, x, R, flag, it1, it2); break;
    case Retractor::PP: retract_pp(cval, xnew, xtilde, x, R, flag, it1, it2); break;
  }
  if (g_stats) { g_stats->retract_outer += *it1; g_stats->retract_pcg += *it2; }
}

/* ------------------------------------------------------------------ linesearch -- src/linesearch.jl */
struct LsOut { int flag; int64_t it1, it2; double newf, f_diff, step_diff, alpha; };

/* armijo! :32-89 */
static LsOut armijo(double *xnew, const double *x, int64_t n, int64_t N, const double *d, const double *g, Problem &P,
                    double fval, Retractor &R, double *cval, const orc_params &prm, double *xtilde) {
  LsOut o; o.f_diff = INF; o.step_diff = INF; o.alpha = prm.alpha; o.flag = 0; o.it1 = 0; o.it2 = 0; o.newf = 0.0;
  double ar_dot = dot(d, g, N);
  double *step = xtilde;
  while (o.step_diff > prm.eps_x) {
    for (int64_t k = 0; k < N; k++) xtilde[k] = x[k] + o.alpha * d[k];
    int64_t i1, i2; int flag;
    retract(cval, xnew, xtilde, x, R, &flag, &i1, &i2);
    o.flag = flag; o.it1 += i1; o.it2 += i2;
    if (g_stats) g_stats->armijo_trials++;
    if (flag > 0) { o.alpha *= prm.s; continue; }            /* :57-60 */
    for (int64_t k = 0; k < N; k++) step[k] = xnew[k] - x[k];
    o.newf = P.f(xnew);
    o.step_diff = nrm2(step, n);                             /* :66 first n entries only */
    o.f_diff = std::fabs(o.newf - fval);
    if (prm.disable_linesearch) break;
    if ((o.newf - fval) <= prm.sigma * o.alpha * ar_dot) break; /* :75 */
    o.alpha *= prm.s;
    if (o.alpha < 1e-100) { o.flag = 99; break; }            /* :82-85 */
  }
  return o;
}

/* exact_linesearch! :107-339 */
static LsOut exact_linesearch(double *xnew, const double *x, int64_t n, int64_t N, const double *d, Problem &P, double fval,
                              Retractor &R, double *cval, const orc_params &prm, vec *tmp) {
  const double phi1 = (3 - std::sqrt(5.0)) / 2, phi2 = (std::sqrt(5.0) - 1) / 2, phi3 = (std::sqrt(5.0) + 1) / 2;
  double Delta = prm.alpha;
  LsOut o; o.flag = 0; o.it1 = 0; o.it2 = 0; o.newf = 0.0;
  double f_a = 0, f_b = 0, f_c = 0, f_d = 0, a_a = 0, a_b = 0, a_c = 0, a_d = 0;
  double *x_a = tmp[0].data(), *x_b = tmp[1].data(), *x_c = tmp[2].data(), *x_d = tmp[3].data(), *swp;
  double *step = tmp[0].data();
  bool do_shrinking = true;
  int flag = 0; int64_t i1, i2;
  auto RET = [&](double *pt) { retract(cval, xnew, pt, x, R, &flag, &i1, &i2); o.it1 += i1; o.it2 += i2; std::copy(xnew, xnew + N, pt); };
  std::copy(x, x + N, x_d); f_d = fval;
  while (true) {                                             /* growing :150-189 */
    swp = x_b; x_b = x_c; x_c = x_d; x_d = swp;
    f_b = f_c; f_c = f_d; a_b = a_c; a_c = a_d;
    for (int64_t k = 0; k < N; k++) x_d[k] = x[k] + (a_d + Delta) * d[k];
    RET(x_d);
    a_d += Delta;
    if (flag > 0 || a_d > 1.0) { f_d = INF; break; }
    f_d = P.f(x_d);
    if (f_d > f_c) break;
    do_shrinking = false;
    Delta *= phi3;
  }
  if (do_shrinking) {                                        /* :192-239 */
    f_b = fval; a_b = 0.0; std::copy(x, x + N, x_b);
    f_c = INF; a_c = Delta;
    swp = x_d; x_d = x_c; x_c = swp;
    while (true) {
      swp = x_d; x_d = x_c; x_c = swp;
      f_d = f_c; a_d = a_c;
      for (int64_t k = 0; k < N; k++) x_c[k] = x[k] + (phi1 * a_c) * d[k];
      RET(x_c);
      a_c *= phi1;
      if (flag > 0 || a_c > 1.0) f_c = INF; else f_c = P.f(x_c);
      if (f_c <= fval || a_c < 1e-100) break;
    }
  }
  f_a = f_b; f_b = f_c; a_a = a_b; a_b = a_c;                /* :242-251 */
  swp = x_a; x_a = x_b; x_b = x_c; x_c = swp;
  a_c = a_a + phi2 * (a_d - a_a);
  for (int64_t k = 0; k < N; k++) x_c[k] = x[k] + a_c * d[k];
  RET(x_c);
  if (flag > 0 || a_c > 1.0) f_c = INF; else f_c = P.f(x_c);
  double nd = nrm2(d, N);
  while ((a_c - a_b) > 1e-6 * nd) {                          /* :270-322 */
    if (f_b < f_c || std::isinf(f_c)) {
      swp = x_d; x_d = x_c; x_c = x_b; x_b = swp;
      f_d = f_c; f_c = f_b; a_d = a_c; a_c = a_b;
      a_b = a_a + phi1 * (a_d - a_a);
      for (int64_t k = 0; k < N; k++) x_b[k] = x[k] + a_b * d[k];
      RET(x_b);
      f_b = P.f(x_b);
    } else {
      swp = x_a; x_a = x_b; x_b = x_c; x_c = swp;
      f_a = f_b; f_b = f_c; a_a = a_b; a_b = a_c;
      a_c = a_a + phi2 * (a_d - a_a);
      for (int64_t k = 0; k < N; k++) x_c[k] = x[k] + a_c * d[k];
      RET(x_c);
      if (flag > 0 || a_c > 1.0) f_c = INF; else f_c = P.f(x_c);
    }
  }
  (void)f_a; (void)f_d;
  if (f_b < f_c) { std::copy(x_b, x_b + N, xnew); o.newf = f_b; o.alpha = a_b; }
  else { std::copy(x_c, x_c + N, xnew); o.newf = f_c; o.alpha = a_c; }
  for (int64_t k = 0; k < N; k++) step[k] = xnew[k] - x[k];
  o.step_diff = nrm2(step, n);
  o.f_diff = std::fabs(o.newf - fval);
  o.flag = flag;
  return o;
}

/* ------------------------------------------------------------------ core driver -- src/optimize.jl:119-443 */
static int core(Problem &P, const double *x0, const double *xl, const double *xu, const orc_params &prm, double *x_out,
                double *obj_hist, int64_t obj_cap, int64_t *obj_len, double *lambda, orc_term *term) {
  int64_t n = P.n, m = P.m;
  if (prm.beta > 0 && !t_noise) return -10; /* stochastic perturbation (optimize.jl:264-273): needs the noise rows (orc_set_noise) */
  bool ineq;
  if (!xl && !xu) ineq = false;                               /* :151 */
  else {
    bool alll = true, allu = true;
    for (int64_t i = 0; i < n; i++) { if (!(xl[i] == -INF)) alll = false; if (!(xu[i] == INF)) allu = false; }
    ineq = !(alll && allu);
  }
  IneqData idata;
  vec PJct, lamy;
  if (ineq) {
    for (int64_t i = 0; i < n; i++) if (xl[i] > xu[i]) return -2; /* :160-162 */
    PJct.resize(2 * n * m); lamy.assign(n, 0.0);
    idata = IneqData(xl, xu, n);
  }
  int64_t N = ineq ? 2 * n : n, M = ineq ? m + n : m;           /* :172-173 */
  vec x(N), xtilde(N), xnew(N), Jc(m * n), Jct(n * m), g(N, 0.0), d(N), tmp_m(M), cval(m, 0.0), lam(m, 0.0);
  std::copy(x0, x0 + n, x.begin());
  if (ineq) generate_initial_y(x.data(), idata);
  vec U(N * m), Sig(m), Vt(m * m); SvdWork svdw;
  vec newton_d(N), newton_dl(M), newton_b2(M, 0.0);
  ProjCGWork cgw(N, M);
  double prev_grad_norm = 0.0, grad_norm = INF;
  IneqDecomp de; de.U = U.data(); de.Sig = Sig.data(); de.Vt = Vt.data(); de.Jct = Jct.data(); de.rank = m; de.n = n; de.m = m;
  if (ineq) { de.Dx.resize(n); de.Dy.resize(n); de.S.resize(n); }
  IneqProjector ineqproject(de);
  LinOp newton_map;
  if (ineq) newton_map = [&](double *dest, const double *src) { augmented_hess_lag_vec(dest, src, P, x.data(), lam.data(), lamy.data(), idata); };
  else newton_map = [&](double *dest, const double *src) { P.hess_lag_vec(dest, src, x.data(), lam.data()); };
  NRWork nrw(m); PPWork ppw(m, n, M, N);
  Retractor R; R.P = &P; R.n = n; R.m = m; R.N = N; R.ineq = ineq; R.idata = &idata; R.idecomp = &de;
  R.U = U.data(); R.Sig = Sig.data(); R.Vt = Vt.data(); R.tol = prm.eps_c; R.mu0 = prm.mu0;
  R.maxiter = prm.maxiter_retract; R.maxiter_pcg = prm.maxiter_pcg; R.nrw = &nrw; R.ppw = &ppw;
  vec exact_tmp[4]; if (prm.linesearch == 1) for (auto &v : exact_tmp) v.resize(N);

  int64_t i = 0; double f_diff = INF, step_diff = INF, kkt_diff = INF;
  double fval = P.f(x.data());
  int64_t nobj = 0;
  if (nobj < obj_cap) obj_hist[nobj] = fval; nobj++;
  if (m > 0) P.c(cval.data(), x.data());
  int term_cond = 0;
  while (true) {
    P.grad(g.data(), x.data());                                 /* :259 (g[n+1:] stays 0) */
    for (int64_t k = 0; k < N; k++) d[k] = -1.0 * g[k];
    if (prm.beta > 0 && i < g_noise_T) {                        /* :264-273 with the supplied randn! rows */
      const double coef = prm.t_beta > 0 ? prm.beta * std::max(1.0 - (double)i / (double)prm.t_beta, 0.0) : prm.beta;
      const double *nz = t_noise + i * N;
      for (int64_t k = 0; k < N; k++) d[k] += coef * nz[k];
    }
    if (ineq) inequality_gradient(de, x.data(), idata);         /* :277 */
    int64_t rank = m;
    if (m > 0) {
      P.jac(Jc.data(), cval.data(), x.data());                  /* :283 */
      for (int64_t a = 0; a < m; a++) for (int64_t b = 0; b < n; b++) Jct[b + a * n] = Jc[a + b * m]; /* :284 */
      if (ineq) {
        for (int64_t a = 0; a < m; a++) for (int64_t b = 0; b < n; b++) {      /* :288-289 */
          PJct[b + a * 2 * n] = (1.0 - de.Dx[b] * de.Dx[b]) * Jct[b + a * n];
          PJct[n + b + a * 2 * n] = -1.0 * de.Dy[b] * de.Dx[b] * Jct[b + a * n];
        }
        ksvd(PJct.data(), (int)(2 * n), (int)m, U.data(), Sig.data(), Vt.data(), svdw);
      } else {
        ksvd(Jct.data(), (int)n, (int)m, U.data(), Sig.data(), Vt.data(), svdw); /* destroys Jct, as the reference */
      }
      for (int64_t j = 0; j < m; j++) if (Sig[j] < prm.eps_rank) { rank = j; break; } /* :297-302 */
      if (!ineq) {
        gemv('T', N, rank, 1.0, U.data(), N, d.data(), 0.0, tmp_m.data());  /* :306 */
        gemv('N', N, rank, -1.0, U.data(), N, tmp_m.data(), 1.0, d.data()); /* :307 */
      }
    }
    if (ineq) {
      de.rank = rank;
      /* NOTE: in the ineq branch Jct was not destroyed (SVD ran on PJct) */
      ineq_project_mulT(tmp_m.data(), de, d.data());            /* :316 */
      ineq_project_mul(d.data(), de, tmp_m.data(), -1.0, 1.0);  /* :317 */
    }
    kkt_diff = nrminf(d.data(), N);                             /* :320 */
    if (ineq) calculate_lambda_kkt(lam.data(), lamy.data(), tmp_m.data(), de); /* :332 */
    else if (m > 0) {
      for (int64_t j = 0; j < rank; j++) tmp_m[j] /= Sig[j];
      for (int64_t j = rank; j < m; j++) tmp_m[j] = 0.0;
      gemv('T', m, m, 1.0, Vt.data(), m, tmp_m.data(), 0.0, lam.data()); /* :342 */
    }
    if (f_diff <= prm.eps_f) { term_cond = 0; break; }          /* :347-359 */
    else if (step_diff <= prm.eps_x) { term_cond = 1; break; }
    else if (i >= prm.maxiter) { term_cond = 3; break; }
    else if (kkt_diff <= prm.eps_kkt) { term_cond = 2; break; }

    if (prm.do_newton) {                                        /* :364-390 */
      int64_t clen = ineq ? n + rank : rank;
      grad_norm = nrm2(d.data(), N);
      double tol = prm.tn_kappa * std::min(1.0, grad_norm / prev_grad_norm) * grad_norm;
      prev_grad_norm = grad_norm;
      DenseProjector dp(U.data(), N, rank);
      Projector &Q = ineq ? (Projector &)ineqproject : (Projector &)dp;
      int64_t tn_iter; double tn_res;
      projcg(newton_d.data(), newton_dl.data(), newton_map, Q, d.data(), newton_b2.data(), N, clen, tol, prm.tn_maxiter,
             cgw, &tn_iter, &tn_res);
      if (dot(newton_d.data(), d.data(), N) > 0.0) { d = newton_d; if (g_stats) g_stats->newton_accepted++; }
    }
    if (m > 0) R.kind = (rank == m && !prm.do_project_retract) ? Retractor::NR : Retractor::PP; /* :396-412 */
    else R.kind = ineq ? Retractor::YRETRACT : Retractor::EUCLIDEAN;
    /* the non-ineq SVD destroyed Jct; nothing downstream reads it (PP recomputes its own J) */
    LsOut o;
    if (prm.linesearch == 0 || prm.disable_linesearch)
      o = armijo(xnew.data(), x.data(), n, N, d.data(), g.data(), P, fval, R, cval.data(), prm, xtilde.data());
    else
      o = exact_linesearch(xnew.data(), x.data(), n, N, d.data(), P, fval, R, cval.data(), prm, exact_tmp);
    f_diff = o.f_diff; step_diff = o.step_diff;
    x = xnew; fval = o.newf;                                    /* :424-426 */
    if (nobj < obj_cap) obj_hist[nobj] = fval; nobj++;
    i++;
  }
  std::copy(x.begin(), x.begin() + n, x_out);
  for (int64_t j = 0; j < m; j++) lambda[j] = lam[j];
  *obj_len = nobj;
  term->condition = term_cond; term->f_diff = f_diff; term->step_diff = step_diff; term->kkt_diff = kkt_diff; term->iter = i;
  return 0;
}

/* ------------------------------------------------------------------ exported API */
extern "C" void orc_default_params(orc_params *p) { /* src/LFPSQP.jl:57-81 */
  memset(p, 0, sizeof(*p));
  p->alpha = 1.0; p->beta = 0.0; p->t_beta = 0; p->s = 0.5; p->sigma = 1e-4; p->eps_c = 1e-6; p->eps_f = 1e-6;
  p->eps_x = 0.0; p->eps_kkt = 1e-6; p->eps_rank = 1e-10; p->maxiter = 10000; p->maxiter_retract = 100;
  p->maxiter_pcg = 100; p->mu0 = 1e-2; p->disable_linesearch = 0; p->do_project_retract = 1; p->disp = 0;
  p->linesearch = 0; p->do_newton = 1; p->tn_maxiter = 10000; p->tn_kappa = 0.5; p->callback_period = 100;
}

extern "C" int orc_optimize(int family, const double *fam_params, int64_t n, int64_t m, int64_t p, const double *x0,
                            const double *xl, const double *xu, const orc_params *prm, double *x_out, double *obj_hist,
                            int64_t obj_cap, int64_t *obj_len, double *lambda, orc_term *term, orc_stats *stats) {
  auto F = make_family(family, fam_params, n, m, p);
  if (!F) return -1;
  orc_stats local; memset(&local, 0, sizeof(local));
  g_stats = &local; g_flops = 0.0;
  int rc;
  if (p == 0) {                                                 /* optimize.jl:15-17 -> :88-104 */
    PlainProblem P(*F);
    rc = core(P, x0, xl, xu, *prm, x_out, obj_hist, obj_cap, obj_len, lambda, term);
  } else {                                                      /* optimize.jl:13-71 via :83-85 (dl=-Inf, du=0) */
    SlackProblem P(*F);
    vec x0a(n + p), xla(n + p), xua(n + p), xo(n + p);
    std::copy(x0, x0 + n, x0a.begin());
    F->d(x0a.data() + n, x0);                                   /* :28 */
    for (int64_t i = 0; i < n; i++) { xla[i] = xl ? xl[i] : -INF; xua[i] = xu ? xu[i] : INF; }
    for (int64_t k = 0; k < p; k++) { xla[n + k] = -INF; xua[n + k] = 0.0; }
    rc = core(P, x0a.data(), xla.data(), xua.data(), *prm, xo.data(), obj_hist, obj_cap, obj_len, lambda, term);
    if (rc == 0) std::copy(xo.begin(), xo.begin() + n, x_out);  /* :68 */
  }
  local.flops = g_flops;
  if (stats) *stats = local;
  g_stats = nullptr;
  return rc;
}

extern "C" int orc_optimize_batched(int family, const double *fam_params, int64_t fam_stride, int64_t n, int64_t m,
                                    int64_t p, int64_t B, const double *x0, const double *xl, const double *xu,
                                    const orc_params *prm, double *x_out, double *obj_hist, int64_t H, int64_t *obj_len,
                                    double *lambda, orc_term *term, orc_stats *stats, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<int64_t> next(0); std::atomic<int> err(0);
  auto worker = [&]() {
    for (;;) {
      int64_t k0 = next.fetch_add(64);
      if (k0 >= B) break;
      for (int64_t k = k0; k < std::min(B, k0 + 64); k++) {
        t_noise = g_noise ? g_noise + k * g_noise_stride : nullptr;
        int rc = orc_optimize(family, fam_params ? fam_params + k * fam_stride : nullptr, n, m, p, x0 + k * n, xl, xu, prm,
                              x_out + k * n, obj_hist + k * H, H, obj_len + k, lambda + k * (m + p), term + k,
                              stats ? stats + k : nullptr);
        if (rc) err = rc;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(worker);
  worker();
  for (auto &t : th) t.join();
  return err;
}

extern "C" int orc_projcg_dense(int64_t n, int64_t mU, const double *A, const double *U, const double *b, const double *c,
                                double tol, int64_t maxit, double *x, double *lam, int64_t *iters, double *nr) {
  ProjCGWork w(n, mU);
  LinOp op = [&](double *dest, const double *src) { gemv('N', n, n, 1.0, A, n, src, 0.0, dest); };
  DenseProjector P(U, n, mU);
  if (maxit < 0) maxit = n + mU;
  projcg(x, lam, op, P, b, c, n, mU, tol, maxit, w, iters, nr);
  return 0;
}

/* projcg! with a DIAGONAL operator A = diag(hd) (the Lagrangian Hessian of the DIAGQUAD / C5 family is diagonal: the
 * reference's hess_lag_vec! closure would be an elementwise product) and a dense orthonormal U: the reference's projected
 * CG at the BASELINE C5 shape without materialising an n x n matrix.  Used by bench.py's large-n cpu_baseline. */
extern "C" int orc_projcg_diag(int64_t n, int64_t mU, const double *hd, const double *U, const double *b, const double *c,
                               double tol, int64_t maxit, double *x, double *lam, int64_t *iters, double *nr) {
  ProjCGWork w(n, mU);
  LinOp op = [&](double *dest, const double *src) { for (int64_t i = 0; i < n; i++) dest[i] = hd[i] * src[i]; g_flops += (double)n; };
  DenseProjector P(U, n, mU);
  if (maxit < 0) maxit = n + mU;
  projcg(x, lam, op, P, b, c, n, mU, tol, maxit, w, iters, nr);
  return 0;
}

extern "C" int orc_pcg_dense(int64_t m, int64_t n, double mu, const double *J, double *x, double *r, double tol,
                             int64_t maxiter, int64_t *iters) {
  vec p(n), z(n), tmp(m);
  FullJac fj; fj.ineq = false; fj.J = J; fj.m = m; fj.n = n; fj.de = nullptr;
  return pcg(mu, fj, x, r, p.data(), z.data(), tmp.data(), tol, maxiter, iters);
}

extern "C" int orc_retract(int family, const double *fam_params, int64_t n, int64_t m, int method, const double *xbase,
                           const double *xtilde, double tol, int64_t maxiter, int64_t maxiter_pcg, double mu0,
                           double *xnew, double *cval, int64_t *iters, int64_t *pcg_iters) {
  auto F = make_family(family, fam_params, n, m, 0);
  if (!F) return -1;
  PlainProblem P(*F);
  vec Jc(m * n), Jct(n * m), U(n * m), Sig(m), Vt(m * m); SvdWork sw;
  P.jac(Jc.data(), cval, xbase);
  for (int64_t a = 0; a < m; a++) for (int64_t b = 0; b < n; b++) Jct[b + a * n] = Jc[a + b * m];
  ksvd(Jct.data(), (int)n, (int)m, U.data(), Sig.data(), Vt.data(), sw);
  NRWork nrw(m); PPWork ppw(m, n, m, n); IneqData id; IneqDecomp de;
  Retractor R; R.P = &P; R.n = n; R.m = m; R.N = n; R.ineq = false; R.idata = &id; R.idecomp = &de;
  R.U = U.data(); R.Sig = Sig.data(); R.Vt = Vt.data(); R.tol = tol; R.mu0 = mu0; R.maxiter = maxiter;
  R.maxiter_pcg = maxiter_pcg; R.nrw = &nrw; R.ppw = &ppw;
  R.kind = method == 0 ? Retractor::NR : Retractor::PP;
  int flag; int64_t i1, i2;
  retract(cval, xnew, xtilde, xbase, R, &flag, &i1, &i2);
  *iters = i1; *pcg_iters = i2;
  return flag;
}

extern "C" int orc_linesearch_euclid(int family, const double *fam_params, int64_t n, const double *x, const double *d,
                                     int which, const orc_params *prm, double *xnew, double *newf, double *f_diff,
                                     double *step_diff, double *alpha) {
  auto F = make_family(family, fam_params, n, 0, 0);
  if (!F) return -1;
  PlainProblem P(*F);
  vec g(n), xt(n), cval(1); P.grad(g.data(), x);
  double fval = P.f(x);
  Retractor R; R.kind = Retractor::EUCLIDEAN; R.N = n; R.n = n; R.P = &P;
  LsOut o;
  if (which == 0) o = armijo(xnew, x, n, n, d, g.data(), P, fval, R, cval.data(), *prm, xt.data());
  else { vec tmp[4]; for (auto &v : tmp) v.resize(n); o = exact_linesearch(xnew, x, n, n, d, P, fval, R, cval.data(), *prm, tmp); }
  *newf = o.newf; *f_diff = o.f_diff; *step_diff = o.step_diff; *alpha = o.alpha;
  return o.flag;
}

extern "C" void orc_ineq_data(int64_t n, const double *xl, const double *xu, double *q, double *r, double *s, double *t,
                              int32_t *isline, int32_t *isparabola) {
  IneqData id(xl, xu, n);
  for (int64_t i = 0; i < n; i++) { q[i] = id.q[i]; r[i] = id.r[i]; s[i] = id.s[i]; t[i] = id.t[i]; isline[i] = id.isline[i]; isparabola[i] = id.isparabola[i]; }
}
extern "C" void orc_ineq_initial_y(int64_t n, const double *xl, const double *xu, double *xaug) {
  IneqData id(xl, xu, n); generate_initial_y(xaug, id);
}
extern "C" void orc_ineq_h(int64_t n, const double *xl, const double *xu, const double *xaug, double *h) {
  IneqData id(xl, xu, n); calculate_h(h, xaug, id);
}
extern "C" void orc_ineq_gradient(int64_t n, const double *xl, const double *xu, const double *xaug, double *Dx, double *Dy,
                                  double *S) {
  IneqData id(xl, xu, n); IneqDecomp de; de.n = n; de.Dx.resize(n); de.Dy.resize(n); de.S.resize(n);
  inequality_gradient(de, xaug, id);
  std::copy(de.Dx.begin(), de.Dx.end(), Dx); std::copy(de.Dy.begin(), de.Dy.end(), Dy); std::copy(de.S.begin(), de.S.end(), S);
}
extern "C" void orc_y_retract(int64_t n, const double *xl, const double *xu, const double *xaug, double *xnewaug) {
  IneqData id(xl, xu, n); y_retract(xnewaug, xaug, id);
}
extern "C" void orc_ineq_mul(int op, int64_t n, int64_t m, int64_t rank, const double *Dx, const double *Dy,
                             const double *S, const double *Jct, const double *U, const double *in, double *out) {
  IneqDecomp de; de.n = n; de.m = m; de.rank = rank; de.Dx.assign(Dx, Dx + n); de.Dy.assign(Dy, Dy + n); de.S.assign(S, S + n);
  de.Jct = (double *)Jct; de.U = (double *)U;
  switch (op) {
    case 0: ineq_project_mul(out, de, in, 1.0, 0.0); break;
    case 1: ineq_project_mulT(out, de, in); break;
    case 2: ineq_bigA_mul(out, de, in, 1.0, 0.0); break;
    case 3: ineq_bigA_mulT(out, de, in); break;
  }
}
extern "C" void orc_ineq_lambda(int64_t n, int64_t m, const double *Dx, const double *Dy, const double *S, const double *Jct,
                                const double *d, double *lambda, double *lambda_y) {
  /* builds PJct, its thin SVD, Q'd and then calculate_lambda_kkt!, as optimize.jl:288-291,316,332 */
  vec PJ(2 * n * m), U(2 * n * m), Sig(m), Vt(m * m), QtF(n + m); SvdWork sw;
  for (int64_t a = 0; a < m; a++) for (int64_t b = 0; b < n; b++) {
    PJ[b + a * 2 * n] = (1.0 - Dx[b] * Dx[b]) * Jct[b + a * n];
    PJ[n + b + a * 2 * n] = -1.0 * Dy[b] * Dx[b] * Jct[b + a * n];
  }
  ksvd(PJ.data(), (int)(2 * n), (int)m, U.data(), Sig.data(), Vt.data(), sw);
  IneqDecomp de; de.n = n; de.m = m; de.rank = m; de.Dx.assign(Dx, Dx + n); de.Dy.assign(Dy, Dy + n); de.S.assign(S, S + n);
  de.Jct = (double *)Jct; de.U = U.data(); de.Sig = Sig.data(); de.Vt = Vt.data();
  ineq_project_mulT(QtF.data(), de, d);
  calculate_lambda_kkt(lambda, lambda_y, QtF.data(), de);
}

[[maybe_unused]] static std::unique_ptr<Problem> make_problem(Family &F) {
  if (F.p == 0) return std::unique_ptr<Problem>(new PlainProblem(F));
  return std::unique_ptr<Problem>(new SlackProblem(F));
}
extern "C" double orc_family_f(int family, const double *prm, int64_t n, int64_t m, int64_t p, const double *x) {
  auto F = make_family(family, prm, n, m, p); return F ? F->f(x) : QNAN;
}
extern "C" void orc_family_grad(int family, const double *prm, int64_t n, int64_t m, int64_t p, const double *x, double *g) {
  auto F = make_family(family, prm, n, m, p); if (F) F->grad(g, x);
}
extern "C" void orc_family_jac(int family, const double *prm, int64_t n, int64_t m, int64_t p, const double *x, double *Jc,
                               double *cval) {
  auto F = make_family(family, prm, n, m, p); if (!F) return;
  int64_t M = m + p;
  for (int64_t k = 0; k < M * n; k++) Jc[k] = 0;
  if (m) F->jac(Jc, M, cval, x);
  if (p) F->jacd(Jc + m, M, cval + m, x);
}
extern "C" void orc_family_hess(int family, const double *prm, int64_t n, int64_t m, int64_t p, const double *x,
                                const double *lam, const double *src, double *dest) {
  auto F = make_family(family, prm, n, m, p); if (F) F->hess(dest, src, x, lam, lam + m);
}
