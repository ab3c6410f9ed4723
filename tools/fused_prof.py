import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, importlib
L = importlib.import_module("lfpsqp.jl_b200")
L.default_context(0)
n, m, K = 65536, 2048, 64
Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=0, cond=1e3)
P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
P.factor(x0, want=())
lam = np.zeros(m)
for _ in range(4):
    r = P.projcg(x0, lam=lam, tol=0.0, maxit=K, chunk=16, want_solution=False)
    print("%.1f us/iteration" % (1e3 * r["ms"] / K), flush=True)
