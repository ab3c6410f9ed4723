"""GPU box: persistent fused projcg (large_fused.cu) vs the multi-kernel loop -- same results, time per iteration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, importlib
L = importlib.import_module("lfpsqp.jl_b200")
L.default_context(0)

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

def run(n, m, seed, fused, K=None, tol=1e-9, reps=1):
    os.environ["LFPSQP_FUSED_PROJCG"] = "1" if fused else "0"
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=1e3)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
    P.factor(x0, want=())
    lam = np.random.default_rng(1).standard_normal(m) * 0.1
    out = []
    for _ in range(reps):
        if K is None: r = P.projcg(x0, lam=lam, tol=tol, maxit=10000)
        else: r = P.projcg(x0, lam=lam, tol=0.0, maxit=K, chunk=16)
        out.append(r)
    return out[-1], P.ctx.last_launches

for (n, m) in [(2048, 96), (4096, 200), (1000, 130), (770, 1), (20000, 512)]:
    a, la = run(n, m, 3, False); b, lb = run(n, m, 3, True)
    print("n=%d m=%d tol: iters %d/%d status %d/%d nr %.3e/%.3e sol rel diff %.2e launches %d/%d" % (
        n, m, a["iters"], b["iters"], a["status"], b["status"], a["nr"], b["nr"], rel(b["sol"], a["sol"]), la, lb), flush=True)
    a, la = run(n, m, 3, False, K=7); b, lb = run(n, m, 3, True, K=7)
    print("          K=7: iters %d/%d status %d/%d sol rel diff %.2e" % (a["iters"], b["iters"], a["status"], b["status"], rel(b["sol"], a["sol"])), flush=True)
# negative curvature exit
os.environ["LFPSQP_FUSED_PROJCG"] = "1"
n, m = 2048, 96
Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=11, cond=1e3)
for fused in (0, 1):
    os.environ["LFPSQP_FUSED_PROJCG"] = str(fused)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w)); P.factor(x0, want=())
    lam2 = -50.0 * np.abs(np.random.default_rng(2).standard_normal(m))
    r = P.projcg(x0, lam=lam2, tol=1e-20, maxit=10000)
    print("negcurv fused=%d status %d iters %d |x| %.15f" % (fused, r["status"], r["iters"], np.linalg.norm(r["sol"])), flush=True)
# full solves agree
for fused in (0, 1):
    os.environ["LFPSQP_FUSED_PROJCG"] = str(fused)
    Q, A, b, xt, w, x0 = L.make_diagquad(4096, 200, seed=1, cond=100.0)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
    t0 = time.time(); x, obj, lam, info, st, status = P.solve(x0, L.LFPSQPParams(), return_stats=True); t1 = time.time()
    print("solve fused=%d iter %d cond %d f %.15e cg %d  %.3fs" % (fused, info.iter, int(info.condition), obj[-1], st["projcg_iters"], t1 - t0), flush=True)
# pcg! (retractions.jl:179-246): fused one-launch kernel vs the multi-kernel loop
for (n, m) in [(2048, 96), (1000, 130), (20000, 512)]:
    res = []
    for fused in (0, 1):
        os.environ["LFPSQP_FUSED_PROJCG"] = str(fused)
        Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=3, cond=1e3)
        P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
        rhs = np.random.default_rng(5).standard_normal(n)
        res.append(P.pcg(x0, 1e-2, rhs, tol=1e-8, maxiter=100) + (P.ctx.last_launches,))
    (xa, ra, fa, ia, la), (xb, rb, fb, ib, lb) = res
    print("pcg n=%d m=%d: iters %d/%d flag %d/%d x rel diff %.2e |r| %.2e/%.2e launches %d/%d" % (
        n, m, ia, ib, fa, fb, rel(xb, xa), np.linalg.norm(ra), np.linalg.norm(rb), la, lb), flush=True)
# C5 timing
if len(sys.argv) > 1:
    import torch
    n, m, K = 65536, 2048, 64
    for fused in (0, 1):
        a, la = run(n, m, 0, bool(fused), K=K, reps=4)
        print("C5 fused=%d: %.1f us/iteration (%.0f it/s), launches %d, roofline frac vs 6650 GB/s %.3f" % (
            fused, 1e3 * a["ms"] / K, K / a["ms"] * 1e3, la, (16.0 * m * n + 8.0 * m * m + 104.0 * n) / (a["ms"] / K * 1e-3) / 6650e9), flush=True)
    # pcg at C5: wall time of the unit-level call minus its setup (jac + H2D) measured with maxiter = 0
    for fused in (0, 1):
        os.environ["LFPSQP_FUSED_PROJCG"] = str(fused)
        Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=0, cond=1e3)
        P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
        rhs = np.random.default_rng(5).standard_normal(n)
        def timed(mi):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); out = P.pcg(x0, 1e-2, rhs, tol=0.0, maxiter=mi); best = min(best, time.perf_counter() - t0)
            return best, out
        t0_, _ = timed(0); t1_, out = timed(100)
        per = (t1_ - t0_) / 100
        print("C5 pcg fused=%d: %.1f us/iteration (%.0f it/s), iters %d, roofline frac vs 6650 GB/s %.3f" % (
            fused, per * 1e6, 1 / per, out[3], (16.0 * m * n + 88.0 * n) / per / 6650e9), flush=True)
