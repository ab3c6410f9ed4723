import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.Generator(np.random.Philox(key=4))
x0 = rng.standard_normal((npts, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
P = L.LargeProblem(L.families.thomson(npts))
fac = P.factor(x0, want=())
print(fac["gram_ms"])
