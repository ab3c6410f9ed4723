import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
from oracle import oracle as O
n, m = 1000, 130
Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=1, cond=100.0)
fam = L.families.diagquad(Q, A, b, xt, w)
P = L.LargeProblem(fam)
for mi in (1, 2, 3, 4, 5):
    prm = L.LFPSQPParams(do_project_retract=False, maxiter=mi)
    import warnings; warnings.simplefilter("ignore")
    x, obj, lam, info, st, status = P.solve(x0, prm, return_stats=True)
    ox, oobj, olam, ot, ost = O.optimize("diagquad", n, m, 0, x0, fam_params=fam.params, params=O.default_params(do_project_retract=0, maxiter=mi))
    print("maxiter", mi, "gpu obj", obj, "status", status, "trials", st["armijo_trials"], "rout", st["retract_outer"], "flag", st["flag_last"],
          "| orc obj", oobj, "trials", ost["armijo_trials"], "rout", ost["retract_outer"], "xerr", np.linalg.norm(x - ox) / np.linalg.norm(ox))
