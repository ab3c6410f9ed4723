/*
 * oracle/lfpsqp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C++) of the reference algorithm LFPSQP.jl for the hot path
 * named in BASELINE.json.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (lfpsqp.jl_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * Parity pinning: the reference is Julia and Julia is absent from this image,
 * so the oracle cannot be diffed against the reference executable.  It is
 * pinned against the only end-to-end known answer the reference holds
 * (README.md:31-37, Rosenbrock) and the property tests of test/*.jl
 * re-expressed in tests/test_oracle_*.py.  At the LAPACK boundary
 * (dgesvd, la_helper.jl:22-28) parity is UNPINNED: the reference's OpenBLAS
 * version is not fixed by any Manifest.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef LFPSQP_ORACLE_H
#define LFPSQP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/LFPSQP.jl:57-81 (LFPSQPParams) as a POD. */
typedef struct {
  double alpha, beta;
  int64_t t_beta;
  double s, sigma, eps_c, eps_f, eps_x, eps_kkt, eps_rank;
  int64_t maxiter, maxiter_retract, maxiter_pcg;
  double mu0;
  int32_t disable_linesearch, do_project_retract, disp, linesearch, do_newton;
  int32_t _pad;
  int64_t tn_maxiter;
  double tn_kappa;
  int64_t callback_period;
} orc_params;

/* src/LFPSQP.jl:45-51 (TerminationInfo). condition: f_tol=0,x_tol=1,kkt_tol=2,max_iter=3,armijo_error=4 */
typedef struct {
  int32_t condition;
  int32_t _pad;
  double f_diff, step_diff, kkt_diff;
  int64_t iter;
} orc_term;

/* per-solve instrumentation (not part of the reference's return value) */
typedef struct {
  int64_t projcg_iters, projcg_negcurv, armijo_trials, retract_outer, retract_pcg, pp_backtracks;
  int64_t newton_accepted, svd_calls, f_evals;
  double flops; /* FP64 flops in vector/matrix helpers, SVD counted as m(m+1)N + m^3/3 (Gram+Cholesky equivalent) */
} orc_stats;

enum {
  ORC_FAM_ROSENBROCK = 0, /* README.md:18-22; n=2 */
  ORC_FAM_README_EQ = 1,  /* README.md:41-54; f=x.x, c=x1-0.75 */
  ORC_FAM_README_INEQ = 2,/* README.md:57-76; f=coeff.x, d=x.x-1; params=coeff[n] */
  ORC_FAM_THOMSON = 3,    /* n=3N, f=sum 1/|xi-xj|, c_i=|x_i|^2-1 */
  ORC_FAM_DIAGQUAD = 4,   /* params = [Q(m*n row-major), A(m*n row-major), b(m), xt(n), w(n)] */
  ORC_FAM_SIN = 5,        /* test/test_retractions.jl:34-54 constraints; f=0.5|x-t|^2; params=t[n] */
  ORC_FAM_BOXQUAD = 6     /* f=|x-t|^2, optional c=a.x-b (m in {0,1}); params=[t(n), a(n), b] */
};

void orc_default_params(orc_params *p);
/* noise rows for beta > 0 (optimize.jl:264-273): row i (N working entries) of instance k at noise + k*stride + i*N; NULL clears */
void orc_set_noise(const double *noise, int64_t T, int64_t stride_per_instance);
/* dlopen the LAPACK provider (scipy's bundled OpenBLAS) and bind dgesvd. returns 0 on success */
int orc_set_lapack(const char *libpath);

/* optimize(f, c!, d!, x0, xl, xu, m, p, param) -- src/optimize.jl:83-85 and the methods it forwards to.
 * xl/xu may be NULL (= nothing).  obj_hist receives at most obj_cap values, *obj_len the true count.
 * lambda receives m+p values (untruncated, optimize.jl:67-70).  returns 0, or <0 on argument errors. */
int orc_optimize(int family, const double *fam_params, int64_t n, int64_t m, int64_t p,
                 const double *x0, const double *xl, const double *xu, const orc_params *prm,
                 double *x_out, double *obj_hist, int64_t obj_cap, int64_t *obj_len,
                 double *lambda, orc_term *term, orc_stats *stats);

/* B independent instances over nthreads host threads (one instance per thread at a time).
 * x0 is n x B column-major (instance k at x0 + k*n); fam_params stride fam_stride doubles per instance (0 = shared).
 * xl/xu shared across the batch (NULL ok).  obj_hist is H x B. */
int orc_optimize_batched(int family, const double *fam_params, int64_t fam_stride, int64_t n, int64_t m, int64_t p,
                         int64_t B, const double *x0, const double *xl, const double *xu, const orc_params *prm,
                         double *x_out, double *obj_hist, int64_t H, int64_t *obj_len, double *lambda,
                         orc_term *term, orc_stats *stats, int nthreads);

/* ---- unit-level entry points (mirror the reference functions one to one) ---- */

/* projcg! (src/projcg.jl:40-121) with dense symmetric A (n x n col-major) and orthonormal U (n x mU col-major). */
int orc_projcg_dense(int64_t n, int64_t mU, const double *A, const double *U, const double *b, const double *c,
                     double tol, int64_t maxit, double *x, double *lam, int64_t *iters, double *nr);

/* same with a diagonal operator A = diag(hd) (n doubles) */
int orc_projcg_diag(int64_t n, int64_t mU, const double *hd, const double *U, const double *b, const double *c,
                    double tol, int64_t maxit, double *x, double *lam, int64_t *iters, double *nr);

/* pcg! (src/retractions.jl:179-246) with dense J (m x n col-major), no preconditioner. x in/out, r in/out. */
int orc_pcg_dense(int64_t m, int64_t n, double mu, const double *J, double *x, double *r, double tol, int64_t maxiter,
                  int64_t *iters);

/* retract! for NR / ProjPenalty (src/retractions.jl:75-177, :265-441) on a family's equality constraints,
 * factorisation taken at xbase (thin SVD of J(xbase)').  method: 0 = NR, 1 = ProjPenalty. */
int orc_retract(int family, const double *fam_params, int64_t n, int64_t m, int method, const double *xbase,
                const double *xtilde, double tol, int64_t maxiter, int64_t maxiter_pcg, double mu0, double *xnew,
                double *cval, int64_t *iters, int64_t *pcg_iters);

/* armijo! / exact_linesearch! (src/linesearch.jl:32-89, :107-339) with Euclidean retraction on a family objective.
 * which: 0 armijo, 1 exact. */
int orc_linesearch_euclid(int family, const double *fam_params, int64_t n, const double *x, const double *d, int which,
                          const orc_params *prm, double *xnew, double *newf, double *f_diff, double *step_diff,
                          double *alpha);

/* bound embedding (src/inequality_helper.jl). */
void orc_ineq_data(int64_t n, const double *xl, const double *xu, double *q, double *r, double *s, double *t,
                   int32_t *isline, int32_t *isparabola);                         /* :39-85 */
void orc_ineq_initial_y(int64_t n, const double *xl, const double *xu, double *xaug); /* :92-109 */
void orc_ineq_h(int64_t n, const double *xl, const double *xu, const double *xaug, double *h); /* :112-122 */
void orc_ineq_gradient(int64_t n, const double *xl, const double *xu, const double *xaug, double *Dx, double *Dy,
                       double *S);                                                /* :125-141 */
void orc_y_retract(int64_t n, const double *xl, const double *xu, const double *xaug, double *xnewaug); /* retractions.jl:451-500 */
/* op: 0 Q*v, 1 Q'*w, 2 bigA*v, 3 bigA'*w  (inequality_helper.jl:161-271). Jct n x m col-major; U 2n x m; rank<=m */
void orc_ineq_mul(int op, int64_t n, int64_t m, int64_t rank, const double *Dx, const double *Dy, const double *S,
                  const double *Jct, const double *U, const double *in, double *out);
/* calculate_lambda_kkt! (:286-308): given d (2n) and SVD of PJct computes lambda (m) and lambda_y (n). */
void orc_ineq_lambda(int64_t n, int64_t m, const double *Dx, const double *Dy, const double *S, const double *Jct,
                     const double *d, double *lambda, double *lambda_y);

/* family callbacks exposed for cross-checking the device kernels' callbacks */
double orc_family_f(int family, const double *fam_params, int64_t n, int64_t m, int64_t p, const double *x);
void orc_family_grad(int family, const double *fam_params, int64_t n, int64_t m, int64_t p, const double *x, double *g);
void orc_family_jac(int family, const double *fam_params, int64_t n, int64_t m, int64_t p, const double *x,
                    double *Jc /* (m+p) x n col-major: c rows then d rows */, double *cval);
void orc_family_hess(int family, const double *fam_params, int64_t n, int64_t m, int64_t p, const double *x,
                     const double *lam /* m+p */, const double *src, double *dest);

#ifdef __cplusplus
}
#endif
#endif
