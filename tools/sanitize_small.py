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
