"""Small batched + large-n run for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
rng = np.random.default_rng(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode in ("all", "smem"):
    os.environ["LFPSQP_BATCHED_KERNEL"] = "smem"
    co = rng.standard_normal((8, 20)); inf = np.inf * np.ones(20)
    fam = L.families.readme_inequality(co)
    print(L.optimize_batched(fam.f, None, fam.d, np.zeros((8, 20)), -inf, inf, 0, 1)[4]["iter"])
    t = rng.standard_normal((8, 12)); fam = L.families.sin_system(12, 3, t)
    print(L.optimize_batched(fam.f, fam.c, np.zeros((8, 12)), 3, L.LFPSQPParams(do_project_retract=False))[4]["iter"])
    print(L.optimize_batched(fam.f, fam.c, np.zeros((8, 12)), 3, L.LFPSQPParams(linesearch=L.exact))[4]["iter"])
    x0 = rng.standard_normal((4, 6, 3)); x0 /= np.linalg.norm(x0, axis=2, keepdims=True)
    fam = L.families.thomson(6)
    print(L.optimize_batched(fam.f, fam.c, x0.reshape(4, 18), 6)[4]["iter"])
    del os.environ["LFPSQP_BATCHED_KERNEL"]
if mode in ("all", "reg"):
    co = rng.standard_normal((8, 50)); inf = np.inf * np.ones(50)
    fam = L.families.readme_inequality(co)
    print(L.optimize_batched(fam.f, None, fam.d, np.zeros((8, 50)), -inf, inf, 0, 1)[4]["iter"])
    print(L.optimize_batched(L.families.rosenbrock().f, rng.uniform(-2, 2, (64, 2)))[4]["iter"][:8])
if mode in ("all", "large"):
    Q, A, b, xt, w, x0 = L.make_diagquad(300, 70, seed=2, cond=50.0)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
    print(P.solve(x0, L.LFPSQPParams(maxiter=3))[3])
    print(P.solve(x0, L.LFPSQPParams(maxiter=2, do_project_retract=False))[3])
if mode in ("all", "new"):
    # round-1 additions: bounds in large-n mode, host callbacks, fused projcg / pcg kernels (cooperative launch)
    Q, A, b, xt, w, x0 = L.make_diagquad(64, 4, seed=5, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    P = L.LargeProblem(fam)
    P.set_bounds(x0 - 0.5, x0 + 0.7)
    print(P.solve(x0, L.LFPSQPParams(maxiter=3))[3])
    print(P.solve(x0, L.LFPSQPParams(maxiter=2, do_project_retract=False))[3])
    P.set_bounds(None, None)
    print(P.solve(x0, L.LFPSQPParams(maxiter=3))[3])            # fused projcg + fused pcg
    P.factor(x0, want=())
    print(P.projcg(x0, lam=np.zeros(4), tol=1e-8, maxit=50)["iters"], P.pcg(x0, 1e-2, np.ones(64), tol=1e-8, maxiter=50)[3])
    def f(x): return 0.5 * np.sum(w * (x - xt) ** 2)
    def grad(g, x): g[:] = w * (x - xt)
    def c(cv, x): cv[:] = 0.5 * Q @ (x * x) + A @ x - b
    def jac(Jc, cv, x): Jc[:, :] = Q * x[None, :] + A; c(cv, x)
    def hlv(dest, src, x, lam): dest[:] = (w + Q.T @ lam) * src
    print(L.optimize(f, grad, c, jac, hlv, x0, x0 - 0.5, x0 + 0.7, 4, L.LFPSQPParams(maxiter=3))[3])
    Q, A, b, xt, w, x0 = L.make_diagquad(65, 3, seed=6, cond=50.0)   # odd n through the callbacks
    print(L.optimize(f, grad, c, jac, hlv, x0, None, None, 3, L.LFPSQPParams(maxiter=2))[3])
if mode in ("all", "r2"):
    # round-2 additions: lane-group register kernel (LW = 8 / 16, group-masked shuffles, shared-memory stash), caller-supplied
    # noise rows, multi-device context, rank-deficient large-n path (Jacobi eigen-solver), explicit-inverse guard fallback
    co = rng.standard_normal((37, 50)); inf = np.inf * np.ones(50)                # 37: groups of a warp with and without work
    fam = L.families.readme_inequality(co)
    print(L.optimize_batched(fam.f, None, fam.d, np.zeros((37, 50)), -inf, inf, 0, 1)[4]["iter"][:6])
    nz = rng.standard_normal((37, 3, 102))
    print(L.optimize_batched(fam.f, None, fam.d, np.zeros((37, 50)), -inf, inf, 0, 1, L.LFPSQPParams(beta=1e-2, t_beta=3), noise=nz)[4]["iter"][:6])
    co = rng.standard_normal((9, 20)); inf = np.inf * np.ones(20)                 # LW = 16, NPL = 2 instantiation
    fam = L.families.readme_inequality(co)
    print(L.optimize_batched(fam.f, None, fam.d, np.zeros((9, 20)), -inf, inf, 0, 1)[4]["iter"])
    t = 2 * rng.standard_normal((5, 12)); xl = np.r_[-np.inf * np.ones(6), -0.5 * np.ones(6)]; xu = np.r_[np.inf * np.ones(3), 0.4 * np.ones(9)]
    fam = L.families.boxquad(t, a=np.ones(12), b=3.0)
    print(L.optimize_batched(fam.f, fam.c, np.tile(np.clip(np.zeros(12), xl, xu) + 0.25, (5, 1)), xl, xu, 1)[4]["iter"])
    mc = L.MultiContext([0])
    print(L.optimize_batched(L.families.rosenbrock().f, rng.uniform(-2, 2, (33, 2)), ctx=mc)[4]["iter"][:4]); mc.close()
    Q, A, b, xt, w, x0 = L.make_diagquad(96, 12, seed=9, cond=50.0)
    Q[11] = Q[0]; A[11] = A[0]; b[11] = b[0]                                     # duplicated constraint -> pseudo-inverse path
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
    print(P.factor(x0, want=())["rank_deficient"], P.solve(x0, L.LFPSQPParams(maxiter=3))[3])
    os.environ["LFPSQP_EXPLICIT_INVERSE"] = "0"                                  # two triangular phases in the fused projcg
    Q, A, b, xt, w, x0 = L.make_diagquad(96, 12, seed=10, cond=50.0)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w)); P.factor(x0, want=())
    print(P.projcg(x0, lam=np.zeros(12), tol=1e-8, maxit=20)["iters"])
    del os.environ["LFPSQP_EXPLICIT_INVERSE"]
if mode in ("all", "r2f"):
    # factorisation with structure detection: zero-slab Gram (SKIP kernel), block flags, parallel diagonal blocks, chain skip after
    # the read-back, batched / triangle-bounded GEMMs of the recursive inverse with ragged and odd sizes, odd-N split-K workspace
    x0 = rng.standard_normal((150, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
    P = L.LargeProblem(L.families.thomson(150))
    for rep in range(3): fac = P.factor(x0)
    print("thomson-150 factor x3:", fac["rank_deficient"], P.solve(x0, L.LFPSQPParams(maxiter=2))[3])
    for (n, m) in [(1200, 193), (1100, 65), (900, 130)]:
        Q, A, b, xt, w, xq = L.make_diagquad(n, m, seed=n + m, cond=50.0)
        band = np.zeros((m, n), dtype=bool)
        for i in range(m):
            lo = (i * 7) % (n - 200); band[i, lo:lo + 150] = True
        for QQ, AA in ((Q, A), (np.where(band, Q, 0.0), np.where(band, A, 0.0))):
            P = L.LargeProblem(L.families.diagquad(QQ, AA, b, xt, w))
            for rep in range(2): fac = P.factor(xq)
            print(n, m, "factor ok", fac["rank_deficient"])
