"""C4 (Thomson N = 4096) whole-driver solve, first outer iterations: wall time and phase breakdown."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lfpsqp.jl_b200 as L
npts = 4096
rng = np.random.Generator(np.random.Philox(key=4))
p0 = rng.standard_normal((npts, 3)); p0 /= np.linalg.norm(p0, axis=1, keepdims=True)
P = L.LargeProblem(L.families.thomson(npts))
for rep in range(2):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        x, obj, lam, info, st, status = P.solve(p0.ravel(), L.LFPSQPParams(maxiter=int(sys.argv[1]) if len(sys.argv) > 1 else 10), return_stats=True)
        wall = time.perf_counter() - t0
    print("C4 solve: %d outer iterations %.1f ms wall | %s | %s" % (info.iter, wall * 1e3, {k: round(float(v), 1) for k, v in P.phase_ms().items()}, st), flush=True)
