import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
from oracle import oracle as O
rng = np.random.default_rng(11)
B, n = 256, 50
co = rng.standard_normal((B, n)); inf = np.inf * np.ones(n)
fam = L.families.readme_inequality(co)
a = L.optimize_batched(fam.f, None, fam.d, np.zeros((B, n)), -inf, inf, 0, 1, return_stats=True)
os.environ["LFPSQP_BATCHED_KERNEL"] = "smem"
b = L.optimize_batched(fam.f, None, fam.d, np.zeros((B, n)), -inf, inf, 0, 1, return_stats=True)
o = O.optimize_batched("readme_ineq", n, 0, 1, np.zeros((B, n)), xl=-inf, xu=inf, fam_params=co, fam_stride=n, nthreads=8)
for k in ("projcg_iters", "projcg_negcurv", "armijo_trials", "retract_outer", "retract_pcg", "pp_backtracks", "f_evals"):
    print(k, "reg==smem %.3f reg==orc %.3f smem==orc %.3f" % ((a[5][k] == b[5][k]).mean(), (a[5][k] == o[5][k]).mean(), (b[5][k] == o[5][k]).mean()),
          "mean reg %.2f smem %.2f orc %.2f" % (a[5][k].mean(), b[5][k].mean(), o[5][k].mean()))
d = np.nonzero(a[5]["retract_outer"] != o[5]["retract_outer"])[0][:3]
for k in d: print(k, a[5][k], b[5][k], o[5][k])
