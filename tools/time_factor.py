"""Wall time of lfpsqp_large_factor (jac! + Gram + Cholesky + triangular inverse) for Thomson N points: host-timed, synchronised."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.Generator(np.random.Philox(key=4))
x0 = rng.standard_normal((npts, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
P = L.LargeProblem(L.families.thomson(npts))
for rep in range(4):
    t0 = time.perf_counter(); fac = P.factor(x0, want=()); dt = time.perf_counter() - t0
    print("factor wall %.2f ms, gram %.2f ms, launches %d" % (dt * 1e3, fac["gram_ms"], P.ctx.last_launches), flush=True)
