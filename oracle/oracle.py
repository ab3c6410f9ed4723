"""oracle/oracle.py -- ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The oracle is a C++ restatement of the reference (see lfpsqp_oracle.cpp, which cites the reference
file:line for every function).  Parity at the LAPACK boundary is unpinned (no Manifest in the reference).
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblfpsqp_oracle.so")

F_TOL, X_TOL, KKT_TOL, MAX_ITER, ARMIJO_ERROR = range(5)
FAM = dict(rosenbrock=0, readme_eq=1, readme_ineq=2, thomson=3, diagquad=4, sin=5, boxquad=6)


class Params(C.Structure):  # src/LFPSQP.jl:57-81
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("t_beta", C.c_int64), ("s", C.c_double),
                ("sigma", C.c_double), ("eps_c", C.c_double), ("eps_f", C.c_double), ("eps_x", C.c_double),
                ("eps_kkt", C.c_double), ("eps_rank", C.c_double), ("maxiter", C.c_int64),
                ("maxiter_retract", C.c_int64), ("maxiter_pcg", C.c_int64), ("mu0", C.c_double),
                ("disable_linesearch", C.c_int32), ("do_project_retract", C.c_int32), ("disp", C.c_int32),
                ("linesearch", C.c_int32), ("do_newton", C.c_int32), ("_pad", C.c_int32),
                ("tn_maxiter", C.c_int64), ("tn_kappa", C.c_double), ("callback_period", C.c_int64)]


class Term(C.Structure):  # src/LFPSQP.jl:45-51
    _fields_ = [("condition", C.c_int32), ("_pad", C.c_int32), ("f_diff", C.c_double), ("step_diff", C.c_double),
                ("kkt_diff", C.c_double), ("iter", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("projcg_iters", "projcg_negcurv", "armijo_trials", "retract_outer",
                                          "retract_pcg", "pp_backtracks", "newton_accepted", "svd_calls",
                                          "f_evals")] + [("flops", C.c_double)]


TERM_DTYPE = np.dtype([("condition", "<i4"), ("status", "<i4"), ("f_diff", "<f8"), ("step_diff", "<f8"),
                       ("kkt_diff", "<f8"), ("iter", "<i8")])
STATS_DTYPE = np.dtype([(k, "<i8") for k in ("projcg_iters", "projcg_negcurv", "armijo_trials", "retract_outer",
                                              "retract_pcg", "pp_backtracks", "newton_accepted", "svd_calls",
                                              "f_evals")] + [("flops", "<f8")])

_lib = None
_libs = {}


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("lfpsqp_oracle.cpp", "lfpsqp_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _LIB


def _find_openblas():
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so*"))
    if not cands:
        raise RuntimeError("oracle: scipy's bundled OpenBLAS not found (needed for dgesvd)")
    return os.path.abspath(cands[0])


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            build(force=True)
        L = C.CDLL(path)
        L.orc_family_f.restype = C.c_double
        rc = L.orc_set_lapack(_find_openblas().encode())
        if rc != 0:
            raise RuntimeError("oracle: could not bind dgesvd")
        _libs[path] = L
    return _libs[path]


def lib():
    global _lib
    if _lib is None:
        _lib = _load(_LIB)
    return _lib


class variant:
    """Context manager: run the oracle built with FMA contraction (rounding-sensitivity probe, see Makefile)."""

    def __init__(self, name):
        self.path = os.path.join(_HERE, "liblfpsqp_oracle_%s.so" % name)

    def __enter__(self):
        global _lib
        self.prev = lib()
        _lib = _load(self.path)
        return self

    def __exit__(self, *a):
        global _lib
        _lib = self.prev


_noise_keep = None


def set_noise(noise):
    """noise: None, or array (B, T, N) / (T, N) of the randn! rows for beta > 0 (optimize.jl:264-273)."""
    global _noise_keep
    if noise is None:
        _noise_keep = None
        for L in list(_libs.values()) or [lib()]:
            L.orc_set_noise(None, C.c_int64(0), C.c_int64(0))
        return
    a = np.ascontiguousarray(noise, dtype=np.float64)
    if a.ndim == 2:
        a = a[None]
    _noise_keep = a
    lib()
    for L in _libs.values():
        L.orc_set_noise(_dp(a), C.c_int64(a.shape[1]), C.c_int64(a.shape[1] * a.shape[2]))


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def optimize(family, n, m, p, x0, xl=None, xu=None, fam_params=None, params=None, obj_cap=20000):
    """Returns (x, obj_values, lambda_kkt, term(dict), stats(dict)) like optimize.jl:442 (+stats)."""
    L = lib()
    fam = FAM[family] if isinstance(family, str) else family
    prm = params or default_params()
    x0 = _f64(x0); xl = _f64(xl); xu = _f64(xu); fp = _f64(fam_params)
    x = np.empty(n); obj = np.empty(obj_cap); lam = np.zeros(m + p)
    nobj = C.c_int64(0); t = Term(); st = Stats()
    rc = L.orc_optimize(fam, _dp(fp), C.c_int64(n), C.c_int64(m), C.c_int64(p), _dp(x0), _dp(xl), _dp(xu),
                        C.byref(prm), _dp(x), _dp(obj), C.c_int64(obj_cap), C.byref(nobj), _dp(lam), C.byref(t),
                        C.byref(st))
    if rc != 0:
        raise ValueError("oracle optimize failed rc=%d" % rc)
    term = dict(condition=t.condition, f_diff=t.f_diff, step_diff=t.step_diff, kkt_diff=t.kkt_diff, iter=t.iter)
    stats = {k: getattr(st, k) for k, _ in Stats._fields_}
    return x, obj[:min(nobj.value, obj_cap)].copy(), lam, term, stats


def optimize_batched(family, n, m, p, x0, xl=None, xu=None, fam_params=None, fam_stride=0, params=None, H=64,
                     nthreads=1):
    """x0: (B, n) C-contiguous (= n x B column-major).  Returns x (B,n), obj (B,H), obj_len (B,), lam (B,m+p),
    term (B,) structured, stats (B,) structured."""
    L = lib()
    fam = FAM[family] if isinstance(family, str) else family
    prm = params or default_params()
    x0 = _f64(x0); B = x0.shape[0]
    xl = _f64(xl); xu = _f64(xu); fp = _f64(fam_params)
    x = np.empty((B, n)); obj = np.full((B, H), np.nan); olen = np.zeros(B, dtype=np.int64)
    lam = np.zeros((B, max(m + p, 0))); term = np.zeros(B, dtype=TERM_DTYPE); stats = np.zeros(B, dtype=STATS_DTYPE)
    rc = L.orc_optimize_batched(fam, _dp(fp), C.c_int64(fam_stride), C.c_int64(n), C.c_int64(m), C.c_int64(p),
                                C.c_int64(B), _dp(x0), _dp(xl), _dp(xu), C.byref(prm), _dp(x), _dp(obj),
                                C.c_int64(H), olen.ctypes.data_as(C.c_void_p), _dp(lam),
                                term.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("oracle optimize_batched failed rc=%d" % rc)
    return x, obj, olen, lam, term, stats


def projcg_dense(A, U, b, c, tol=1e-6, maxit=-1):
    """projcg! (src/projcg.jl:40-121) with dense A (n x n symmetric) and orthonormal U (n x m)."""
    L = lib()
    n, mU = U.shape
    A = np.asfortranarray(A, dtype=np.float64); U = np.asfortranarray(U, dtype=np.float64)
    b = _f64(b); c = _f64(c)
    x = np.zeros(n); lam = np.zeros(mU); it = C.c_int64(0); nr = C.c_double(0)
    L.orc_projcg_dense(C.c_int64(n), C.c_int64(mU), _dp(A), _dp(U), _dp(b), _dp(c), C.c_double(tol), C.c_int64(maxit),
                       _dp(x), _dp(lam), C.byref(it), C.byref(nr))
    return x, lam, it.value, nr.value


def projcg_diag(hd, U, b, c, tol=1e-6, maxit=-1):
    """projcg! with A = diag(hd) and orthonormal U (n x m, Fortran order): returns x, lam, iters, nr."""
    L = lib()
    n, mU = U.shape
    if not U.flags.f_contiguous:
        U = np.asfortranarray(U, dtype=np.float64)
    hd = _f64(hd); b = _f64(b); c = _f64(c)
    x = np.zeros(n); lam = np.zeros(mU); it = C.c_int64(0); nr = C.c_double(0)
    L.orc_projcg_diag(C.c_int64(n), C.c_int64(mU), _dp(hd), _dp(U), _dp(b), _dp(c), C.c_double(tol), C.c_int64(maxit),
                      _dp(x), _dp(lam), C.byref(it), C.byref(nr))
    return x, lam, it.value, nr.value


def pcg_dense(J, mu, b, tol=1e-6, maxiter=100):
    """pcg! (src/retractions.jl:179-246): returns x, r, flag, iters."""
    L = lib()
    m, n = J.shape
    J = np.asfortranarray(J, dtype=np.float64)
    x = np.zeros(n); r = np.array(b, dtype=np.float64, copy=True); it = C.c_int64(0)
    flag = L.orc_pcg_dense(C.c_int64(m), C.c_int64(n), C.c_double(mu), _dp(J), _dp(x), _dp(r), C.c_double(tol),
                           C.c_int64(maxiter), C.byref(it))
    return x, r, flag, it.value


def retract(family, n, m, method, xbase, xtilde, tol, maxiter=100, maxiter_pcg=100, mu0=1e-2, fam_params=None):
    L = lib()
    fam = FAM[family] if isinstance(family, str) else family
    xbase = _f64(xbase); xtilde = _f64(xtilde); fp = _f64(fam_params)
    xnew = np.zeros(n); cval = np.zeros(m); it = C.c_int64(0); pit = C.c_int64(0)
    flag = L.orc_retract(fam, _dp(fp), C.c_int64(n), C.c_int64(m), C.c_int(0 if method == "nr" else 1), _dp(xbase),
                         _dp(xtilde), C.c_double(tol), C.c_int64(maxiter), C.c_int64(maxiter_pcg), C.c_double(mu0),
                         _dp(xnew), _dp(cval), C.byref(it), C.byref(pit))
    return flag, xnew, cval, it.value, pit.value


def linesearch_euclid(family, n, x, d, which="armijo", fam_params=None, params=None):
    L = lib()
    fam = FAM[family] if isinstance(family, str) else family
    prm = params or default_params()
    x = _f64(x); d = _f64(d); fp = _f64(fam_params)
    xnew = np.zeros(n); newf = C.c_double(); fd = C.c_double(); sd = C.c_double(); al = C.c_double()
    flag = L.orc_linesearch_euclid(fam, _dp(fp), C.c_int64(n), _dp(x), _dp(d), C.c_int(0 if which == "armijo" else 1),
                                   C.byref(prm), _dp(xnew), C.byref(newf), C.byref(fd), C.byref(sd), C.byref(al))
    return flag, xnew, newf.value, fd.value, sd.value, al.value


# ---- bound embedding (src/inequality_helper.jl) ----
def ineq_data(xl, xu):
    L = lib(); xl = _f64(xl); xu = _f64(xu); n = len(xl)
    q, r, s, t = (np.zeros(n) for _ in range(4)); il = np.zeros(n, dtype=np.int32); ip = np.zeros(n, dtype=np.int32)
    L.orc_ineq_data(C.c_int64(n), _dp(xl), _dp(xu), _dp(q), _dp(r), _dp(s), _dp(t), il.ctypes.data_as(C.c_void_p),
                    ip.ctypes.data_as(C.c_void_p))
    return q, r, s, t, il.astype(bool), ip.astype(bool)


def ineq_initial_y(xl, xu, x):
    L = lib(); xl = _f64(xl); xu = _f64(xu); n = len(xl)
    xaug = np.zeros(2 * n); xaug[:n] = x
    L.orc_ineq_initial_y(C.c_int64(n), _dp(xl), _dp(xu), _dp(xaug))
    return xaug


def ineq_h(xl, xu, xaug):
    L = lib(); xl = _f64(xl); xu = _f64(xu); n = len(xl); xaug = _f64(xaug); h = np.zeros(n)
    L.orc_ineq_h(C.c_int64(n), _dp(xl), _dp(xu), _dp(xaug), _dp(h))
    return h


def ineq_gradient(xl, xu, xaug):
    L = lib(); xl = _f64(xl); xu = _f64(xu); n = len(xl); xaug = _f64(xaug)
    Dx, Dy, S = (np.zeros(n) for _ in range(3))
    L.orc_ineq_gradient(C.c_int64(n), _dp(xl), _dp(xu), _dp(xaug), _dp(Dx), _dp(Dy), _dp(S))
    return Dx, Dy, S


def y_retract(xl, xu, xaug, xnewaug):
    L = lib(); xl = _f64(xl); xu = _f64(xu); n = len(xl); xaug = _f64(xaug)
    out = np.array(xnewaug, dtype=np.float64, copy=True)
    L.orc_y_retract(C.c_int64(n), _dp(xl), _dp(xu), _dp(xaug), _dp(out))
    return out


def ineq_mul(op, Dx, Dy, S, Jct, U, rank, v):
    """op: 'Q', 'Qt', 'A', 'At' (inequality_helper.jl:161-271)."""
    L = lib(); n, m = Jct.shape
    Jct = np.asfortranarray(Jct, dtype=np.float64); U = np.asfortranarray(U, dtype=np.float64); v = _f64(v)
    code = dict(Q=0, Qt=1, A=2, At=3)[op]
    out = np.zeros({0: 2 * n, 1: n + rank, 2: 2 * n, 3: n + m}[code])
    L.orc_ineq_mul(C.c_int(code), C.c_int64(n), C.c_int64(m), C.c_int64(rank), _dp(_f64(Dx)), _dp(_f64(Dy)),
                   _dp(_f64(S)), _dp(Jct), _dp(U), _dp(v), _dp(out))
    return out


def ineq_lambda(Dx, Dy, S, Jct, d):
    L = lib(); n, m = Jct.shape
    Jct = np.asfortranarray(Jct, dtype=np.float64); d = _f64(d)
    lam = np.zeros(m); lamy = np.zeros(n)
    L.orc_ineq_lambda(C.c_int64(n), C.c_int64(m), _dp(_f64(Dx)), _dp(_f64(Dy)), _dp(_f64(S)), _dp(Jct), _dp(d),
                      _dp(lam), _dp(lamy))
    return lam, lamy


# ---- family callbacks (for cross-checking the device callbacks) ----
def family_f(family, n, m, p, x, fam_params=None):
    L = lib(); fam = FAM[family] if isinstance(family, str) else family
    return L.orc_family_f(fam, _dp(_f64(fam_params)), C.c_int64(n), C.c_int64(m), C.c_int64(p), _dp(_f64(x)))


def family_grad(family, n, m, p, x, fam_params=None):
    L = lib(); fam = FAM[family] if isinstance(family, str) else family
    g = np.zeros(n)
    L.orc_family_grad(fam, _dp(_f64(fam_params)), C.c_int64(n), C.c_int64(m), C.c_int64(p), _dp(_f64(x)), _dp(g))
    return g


def family_jac(family, n, m, p, x, fam_params=None):
    L = lib(); fam = FAM[family] if isinstance(family, str) else family
    J = np.zeros((m + p, n), order="F"); cval = np.zeros(m + p)
    L.orc_family_jac(fam, _dp(_f64(fam_params)), C.c_int64(n), C.c_int64(m), C.c_int64(p), _dp(_f64(x)), _dp(J),
                     _dp(cval))
    return J, cval


def family_hess(family, n, m, p, x, lam, v, fam_params=None):
    L = lib(); fam = FAM[family] if isinstance(family, str) else family
    out = np.zeros(n)
    L.orc_family_hess(fam, _dp(_f64(fam_params)), C.c_int64(n), C.c_int64(m), C.c_int64(p), _dp(_f64(x)),
                      _dp(_f64(lam)), _dp(_f64(v)), _dp(out))
    return out
