"""Probe (GPU box): large-n mode with finite bounds and the host-callback family vs the CPU oracle, with the
oracle-vs-oracle(+FMA) sensitivity next to it.  Prints one line per case; used to calibrate tests/test_gpu_large_bounds.py."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import importlib
L = importlib.import_module("lfpsqp.jl_b200")
from oracle import oracle as O

L.default_context(0)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def bounds_for(x0, seed, kind="mixed"):
    rng = np.random.default_rng(seed)
    n = x0.size
    xl = x0 - rng.uniform(0.05, 1.0, n); xu = x0 + rng.uniform(0.05, 1.0, n)
    if kind == "mixed":
        k = rng.integers(0, 4, n)
        xl[k == 0] = -np.inf; xu[k == 0] = np.inf       # line
        xu[k == 1] = np.inf                              # lower only
        xl[k == 2] = -np.inf                             # upper only
    return xl, xu


def report(tag, got, orc, orc2=None):
    x, obj, lam, info, st, status = got
    ox, oobj, olam, ot, ost = orc
    s = "%-34s status=%d cond %d/%d iter %d/%d xerr %.2e ferr %.2e lamerr %.2e" % (
        tag, status, int(info.condition), ot["condition"], info.iter, ot["iter"], rel(x, ox),
        abs(obj[-1] - oobj[-1]) / max(abs(oobj[-1]), 1e-300), rel(lam, olam) if lam.size else 0.0)
    s += " | cg %d/%d rout %d/%d rpcg %d/%d bt %d/%d" % (st["projcg_iters"], ost["projcg_iters"], st["retract_outer"], ost["retract_outer"],
                                                         st["retract_pcg"], ost["retract_pcg"], st["pp_backtracks"], ost["pp_backtracks"])
    if orc2 is not None:
        s += " || oracle-fma: iter %d xerr %.2e" % (orc2[3]["iter"], rel(orc2[0], ox))
    print(s, flush=True)


for (n, m, seed) in [(64, 4, 5), (256, 16, 2), (512, 32, 1), (1000, 130, 3), (300, 0, 4)]:
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    for kind in ("mixed", "box"):
        xl, xu = bounds_for(x0, seed + 10, kind)
        for nr in (0, 1):
            if nr and m == 0:
                continue
            # the reference algorithm itself does not terminate on these (NR fails at every alpha: linesearch.jl:57-60
            # has no lower bound on alpha; checked with the oracle on the CPU) -- skip them
            if (n, kind, nr) in ((512, "mixed", 1), (1000, "mixed", 1), (1000, "box", 0), (1000, "mixed", 0)):
                continue
            par = L.LFPSQPParams(do_project_retract=not nr)
            P = L.LargeProblem(fam)
            P.set_bounds(xl, xu)
            t0 = time.time()
            got = P.solve(x0, par, return_stats=True)
            t1 = time.time()
            op = O.default_params(do_project_retract=0 if nr else 1)
            orc = O.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, params=op)
            with O.variant("fma"):
                orc2 = O.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, params=op)
            report("diagquad n=%d m=%d %s %s (%.2fs)" % (n, m, kind, "NR" if nr else "PP", t1 - t0), got, orc, orc2)

# ---------------- host callbacks: the same diagquad problem through numpy callbacks
for (n, m, seed, bounded) in [(64, 4, 5, False), (64, 4, 5, True), (301, 17, 2, False), (301, 17, 2, True)]:
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)

    def f(x): return 0.5 * np.sum(w * (x - xt) ** 2)
    def grad(g, x): g[:] = w * (x - xt)
    def c(cv, x): cv[:] = 0.5 * Q @ (x * x) + A @ x - b
    def jac(Jc, cv, x): Jc[:, :] = Q * x[None, :] + A; c(cv, x)
    def hlv(dest, src, x, lam): dest[:] = (w + Q.T @ lam) * src
    xl, xu = bounds_for(x0, seed + 10) if bounded else (None, None)
    got = L.optimize(f, grad, c, jac, hlv, x0, xl, xu, m, L.LFPSQPParams(), return_stats=True)
    orc = O.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params)
    report("HOST diagquad n=%d m=%d %s" % (n, m, "bounded" if bounded else "free"), got, orc)

# README inequality example (README.md:57-76) as a GENERIC problem: slack wrapper + host callbacks
n = 50
coeff = np.random.default_rng(0).standard_normal(n)
def f(x): return float(coeff @ x)
def grad(g, x): g[:] = coeff
def hv(dest, src, x): dest[:] = 0.0
def d(dv, x): dv[0] = x @ x - 1.0
def djac(Jd, dv, x): Jd[0, :] = 2.0 * x; d(dv, x)
def dhv(dest, src, x, lam): dest[:] = 2.0 * lam[0] * src
got = L.optimize_slack(f, grad, hv, None, None, None, d, djac, dhv, [-np.inf], [0.0], np.zeros(n), None, None, 0, 1, return_stats=True)
orc = O.optimize("readme_ineq", n, 0, 1, np.zeros(n), xl=-np.inf * np.ones(n), xu=np.inf * np.ones(n), fam_params=coeff)
report("HOST README inequality (slack)", got, orc)
print("x* err vs -coeff/|coeff|: %.2e" % rel(got[0], -coeff / np.linalg.norm(coeff)))
