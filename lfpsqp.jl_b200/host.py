"""Generic problems: the explicit-derivative core `optimize(f, grad!, c!, jac!, hess_lag_vec!, x0, xl, xu, m, param)`
(src/optimize.jl:119-443) with HOST callbacks over lfpsqp_solve_host, and the slack wrapper for `d!`
(src/optimize.jl:13-71) on top of it.

The callbacks keep the reference's in-place conventions (src/autodiff_generators.jl:7-9, :40-42, :80-104):
    f(x) -> float            grad(g, x)            c(cval, x)
    jac(Jc, cval, x)         Jc is an (m, n) array (Fortran order, i.e. the reference's column-major Jc); also fills cval
    hess_lag_vec(dest, src, x, lam)
x is the first n entries of the working vector.  The linear algebra of the hot path (Gram + Cholesky, projections,
projcg, retractions, line search, bound embedding) runs on the device; only the callbacks run on the host.
The reference builds grad!/jac!/hess_lag_vec! by automatic differentiation (src/autodiff_generators.jl); that stays
host-side business of the caller (Julia keeps using ForwardDiff/ReverseDiff; Python callers pass derivatives)."""
import ctypes as C

import numpy as np

from . import _lib
from .api import LFPSQPParams, TerminationCondition, TerminationInfo

_D = C.POINTER(C.c_double)
_F = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, C.c_int64, _D)
_G = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, _D, C.c_int64)
_CC = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, _D, C.c_int64, C.c_int64)
_J = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, _D, _D, C.c_int64, C.c_int64)
_H = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, _D, _D, _D, C.c_int64, C.c_int64)
_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, _D, C.c_int64)
_RN = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, C.c_int64)


class HostCallbacks(C.Structure):  # lfpsqp_host_callbacks (include/lfpsqp_b200.h)
    _fields_ = [("user", C.c_void_p), ("f", _F), ("grad", _G), ("c", _CC), ("jac", _J), ("hess_lag_vec", _H),
                ("callback", _CB), ("randn", _RN)]


def _arr(p, *shape, order="C"):
    a = np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),))
    return a.reshape(shape, order=order)


def optimize_explicit(f, grad, c, jac, hess_lag_vec, x0, xl, xu, m, param=None, ctx=None, history=20000,
                      return_stats=False, randn=None):
    """optimize(f, grad!, c!, jac!, hess_lag_vec!, x0, xl, xu, m, param) -> (x, obj_values, λ_kkt, term_info)."""
    param = param or LFPSQPParams()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    m = int(m)
    if (xl is None) != (xu is None):
        raise _lib.LFPSQPError("xl and xu must both be given or both be nothing")
    if xl is not None:
        xl = np.ascontiguousarray(xl, dtype=np.float64); xu = np.ascontiguousarray(xu, dtype=np.float64)
        if not (len(xl) == len(xu) == n):
            raise _lib.LFPSQPError("xl, xu, and x0 must all be the same length")          # optimize.jl:144-148
    err = []

    def guard(fn):
        def wrapped(*a):
            try:
                fn(*a)
                return 0
            except BaseException as e:  # exceptions cannot unwind through C: store, abort the solve, re-raise below
                err.append(e)
                return 1
        return wrapped

    @guard
    def c_f(_, x, nn, out):
        out[0] = float(f(_arr(x, nn)))

    @guard
    def c_grad(_, g, x, nn):
        grad(_arr(g, nn), _arr(x, nn))

    @guard
    def c_c(_, cv, x, nn, mm):
        c(_arr(cv, mm), _arr(x, nn))

    @guard
    def c_jac(_, Jc, cv, x, nn, mm):
        jac(_arr(Jc, mm, nn, order="F"), _arr(cv, mm), _arr(x, nn))

    @guard
    def c_hess(_, dest, src, x, lam, nn, mm):
        hess_lag_vec(_arr(dest, nn), _arr(src, nn), _arr(x, nn), _arr(lam, mm) if mm > 0 else np.zeros(0))

    user_cb = param.callback

    @guard
    def c_cb(_, it, xa, na):
        user_cb(int(it), _arr(xa, na))

    @guard
    def c_rn(_, buf, na):
        _arr(buf, na)[:] = randn(int(na))

    cbs = HostCallbacks()
    cbs.user = None
    cbs.f = _F(c_f); cbs.grad = _G(c_grad); cbs.hess_lag_vec = _H(c_hess)
    if m > 0:
        if c is None or jac is None:
            raise _lib.LFPSQPError("c! and jac! are required when m > 0")
        cbs.c = _CC(c_c); cbs.jac = _J(c_jac)
    if user_cb is not None:
        cbs.callback = _CB(c_cb)
    if randn is not None:
        cbs.randn = _RN(c_rn)
    keep = [getattr(cbs, k) for k, _ in HostCallbacks._fields_ if k != "user"]   # keep the thunks alive during the call
    ctx = ctx or _lib.default_context()
    saved_cb, param.callback = param.callback, None   # to_c() refuses callbacks: this path carries it in the struct
    try:
        cp = param.to_c()
    finally:
        param.callback = saved_cb
    H = int(history)
    x = np.empty(n); obj = np.full(H, np.nan); olen = np.zeros(1, dtype=np.int64)
    lam = np.zeros(max(m, 1)); term = np.zeros(1, dtype=_lib.TERM_DTYPE); stats = np.zeros(1, dtype=_lib.STATS_DTYPE)
    rc = ctx.lib.lfpsqp_solve_host(ctx.h, C.byref(cbs), n, m, _lib.ptr(x0), _lib.ptr(xl), _lib.ptr(xu),
                                   C.cast(C.pointer(cp), C.c_void_p), _lib.ptr(x), _lib.ptr(obj), H, _lib.ptr(olen),
                                   _lib.ptr(lam), _lib.ptr(term), _lib.ptr(stats))
    del keep
    if err:
        raise err[0]
    ctx.check(rc)
    t = term[0]
    info = TerminationInfo(TerminationCondition(int(t["condition"])), float(t["f_diff"]), float(t["step_diff"]),
                           float(t["kkt_diff"]), int(t["iter"]))
    if int(t["iter"]) == param.maxiter:
        import warnings
        warnings.warn("Maximum # of outer iterations reached")                            # optimize.jl:438-440
    res = (x, obj[:min(int(olen[0]), H)].copy(), lam[:m], info)
    if return_stats:
        return res + ({k: int(stats[0][k]) for k in _lib.STATS_FIELDS}, int(t["status"]))
    return res


def slack_callbacks(f, grad, hess_vec, c, jac, chess_vec, d, djac, dhess_vec, dl, du, x0, xl, xu, m, p):
    """The augmented problem of the slack wrapper (src/optimize.jl:23-51) with explicit derivatives in place of the
    reference's AD: returns (f_aux, grad_aux, c_aux, jac_aux, hess_aux, x0_aux, xl_aux, xu_aux) for
    x_aux = [x ; s], s0 = d(x0), bounds [xl ; dl] <= x_aux <= [xu ; du], constraints [c(x) ; d(x) - s]."""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    dl = np.asarray(dl, dtype=np.float64); du = np.asarray(du, dtype=np.float64)
    if not (len(dl) == len(du) == p):
        raise _lib.LFPSQPError("Bound vectors dl and du must be of size p")               # optimize.jl:19-21
    x0a = np.empty(n + p); x0a[:n] = x0; d(x0a[n:], x0)                                  # :26-28
    xl = -np.inf * np.ones(n) if xl is None else np.asarray(xl, float)
    xu = np.inf * np.ones(n) if xu is None else np.asarray(xu, float)
    xla = np.concatenate([xl, dl]); xua = np.concatenate([xu, du])                       # :30-36

    def f_aux(x):
        return f(x[:n])

    def grad_aux(g, x):
        grad(g[:n], x[:n]); g[n:] = 0.0

    def c_aux(cv, x):                                                                    # :42-51
        if m > 0:
            c(cv[:m], x[:n])
        d(cv[m:], x[:n]); cv[m:] -= x[n:]

    def jac_aux(Jc, cv, x):
        Jc[:, :] = 0.0
        if m > 0:
            Jm = np.zeros((m, n), order="F"); jac(Jm, cv[:m], x[:n]); Jc[:m, :n] = Jm
        Jd = np.zeros((p, n), order="F"); djac(Jd, cv[m:], x[:n]); Jc[m:, :n] = Jd
        Jc[m:, n:] = -np.eye(p)
        cv[m:] -= x[n:]

    def hess_aux(dest, src, x, lam):
        dest[:] = 0.0
        t = np.zeros(n)
        hess_vec(t, src[:n], x[:n]); dest[:n] += t
        if m > 0:
            t[:] = 0.0; chess_vec(t, src[:n], x[:n], lam[:m]); dest[:n] += t
        t[:] = 0.0; dhess_vec(t, src[:n], x[:n], lam[m:]); dest[:n] += t

    return f_aux, grad_aux, c_aux, jac_aux, hess_aux, x0a, xla, xua


def optimize_slack(f, grad, hess_vec, c, jac, chess_vec, d, djac, dhess_vec, dl, du, x0, xl, xu, m, p, param=None, **kw):
    """The slack wrapper optimize(f, c!, d!, dl, du, x0, xl, xu, m, p, param) (src/optimize.jl:13-71) with explicit
    derivatives in place of the reference's AD (see slack_callbacks).
        hess_vec(dest, src, x): dest = Hess f(x) src ;  chess_vec / dhess_vec(dest, src, x, lam): dest = sum_i lam_i Hess c_i src
        c(cval, x), jac(Jc, cval, x) / d(dval, x), djac(Jd, dval, x): as the reference (may be None when m == 0)"""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    if d is None or p == 0:                                                               # optimize.jl:15-17
        def hl(dest, src, x, lam):
            hess_vec(dest, src, x)
            if m > 0:
                t = np.zeros(n); chess_vec(t, src, x, lam); dest += t
        return optimize_explicit(f, grad, c, jac, hl, x0, xl, xu, m, param, **kw)
    f_aux, grad_aux, c_aux, jac_aux, hess_aux, x0a, xla, xua = slack_callbacks(f, grad, hess_vec, c, jac, chess_vec, d, djac,
                                                                                 dhess_vec, dl, du, x0, xl, xu, m, p)
    out = optimize_explicit(f_aux, grad_aux, c_aux, jac_aux, hess_aux, x0a, xla, xua, m + p, param, **kw)
    return (out[0][:n],) + tuple(out[1:])                                                # :67-70 (lambda untruncated)
