"""Large-n mode (one instance, dense m x n Jacobian, optionally column-sharded over GPUs) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from .api import LFPSQPParams, TerminationCondition, TerminationInfo
from .families import DIAGQUAD, THOMSON


class LargeProblem:
    """Binds one (sharded) problem to a Context: lfpsqp_large_setup, then solve / factor / project / projcg."""

    def __init__(self, fam, ctx=None, col0=0, n_loc=None, n_global=None, params_dev_ptr=None):
        self.ctx = ctx or _lib.default_context()
        self.fam = fam
        self.m = fam.m
        self.n_loc = fam.n if n_loc is None else n_loc
        self.n = fam.n if n_global is None else n_global
        if params_dev_ptr is not None:
            pp, on_dev = C.c_void_p(params_dev_ptr), 1
        else:
            pp, on_dev = _lib.ptr(fam.params), 0
        self.ctx.check(self.ctx.lib.lfpsqp_large_setup(self.ctx.h, fam.id, self.n, self.m, col0, self.n_loc, pp, on_dev))

    def solve(self, x0, param=None, history=4096, return_stats=False):
        param = param or LFPSQPParams()
        cp = param.to_c()
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.empty(self.n_loc); obj = np.full(history, np.nan); olen = np.zeros(1, dtype=np.int64)
        lam = np.zeros(max(self.m, 1)); term = np.zeros(1, dtype=_lib.TERM_DTYPE); stats = np.zeros(1, dtype=_lib.STATS_DTYPE)
        self.ctx.check(self.ctx.lib.lfpsqp_large_solve(self.ctx.h, _lib.ptr(x0), C.cast(C.pointer(cp), C.c_void_p), _lib.ptr(x),
                                                       _lib.ptr(obj), history, _lib.ptr(olen), _lib.ptr(lam), _lib.ptr(term),
                                                       _lib.ptr(stats)))
        t = term[0]
        info = TerminationInfo(TerminationCondition(int(t["condition"])), float(t["f_diff"]), float(t["step_diff"]),
                               float(t["kkt_diff"]), int(t["iter"]))
        res = (x, obj[:min(int(olen[0]), history)].copy(), lam[:self.m], info)
        if return_stats:
            return res + ({k: int(stats[0][k]) for k in _lib.STATS_FIELDS}, int(t["status"]))
        return res

    def factor(self, x, want=("G", "L", "Linv")):
        m = self.m
        x = np.ascontiguousarray(x, dtype=np.float64)
        G = np.zeros((m, m)) if "G" in want else None
        L = np.zeros((m, m)) if "L" in want else None
        Li = np.zeros((m, m)) if "Linv" in want else None
        rd = C.c_int(0); ms = C.c_double(0)
        self.ctx.check(self.ctx.lib.lfpsqp_large_factor(self.ctx.h, _lib.ptr(x), _lib.ptr(G), _lib.ptr(L), _lib.ptr(Li),
                                                        C.cast(C.pointer(rd), C.c_void_p), C.cast(C.pointer(ms), C.c_void_p)))
        return dict(G=G, L=L, Linv=Li, rank_deficient=rd.value, gram_ms=ms.value)

    def project(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v); lam = np.zeros(max(self.m, 1))
        self.ctx.check(self.ctx.lib.lfpsqp_large_project(self.ctx.h, _lib.ptr(v), _lib.ptr(out), _lib.ptr(lam)))
        return out, lam[:self.m]

    def projcg(self, x, lam=None, tol=0.0, maxit=10, chunk=0, want_solution=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
        sol = np.empty(self.n_loc) if want_solution else None
        it = C.c_int64(0); nr = C.c_double(0); st = C.c_int(0); ms = C.c_double(0)
        self.ctx.check(self.ctx.lib.lfpsqp_large_projcg(self.ctx.h, _lib.ptr(x), _lib.ptr(lam), tol, maxit, chunk, _lib.ptr(sol),
                                                        C.cast(C.pointer(it), C.c_void_p), C.cast(C.pointer(nr), C.c_void_p),
                                                        C.cast(C.pointer(st), C.c_void_p), C.cast(C.pointer(ms), C.c_void_p)))
        return dict(sol=sol, iters=it.value, nr=nr.value, status=st.value, ms=ms.value)


def make_diagquad(n, m, seed=0, cond=1e4, dtype=np.float64):
    """BASELINE config C5 (definition pinned in SURVEY.md 8d / DESIGN.md): Q, A ~ N(0,1)/sqrt(n), x0 ~ N(0,1),
    b such that c(x0) = 0, f = 1/2 (x-xt)' diag(w) (x-xt) with w log-uniform in [1, cond]."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    Q = rng.standard_normal((m, n)) / np.sqrt(n); A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.standard_normal(n); xt = rng.standard_normal(n)
    w = np.exp(rng.uniform(0.0, np.log(cond), n))
    b = 0.5 * Q @ (x0 * x0) + A @ x0
    return Q, A, b, xt, w, x0
