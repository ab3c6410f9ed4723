"""C2 batched kernel time (ms) for the library selected by LFPSQP_LIB_PATH (occupancy experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, importlib
L = importlib.import_module("lfpsqp.jl_b200")
ctx = L.default_context(0)
B, n = 65536, 50
co = np.random.Generator(np.random.Philox(key=0)).standard_normal((B, n))
inf = np.inf * np.ones(n)
fam = L.families.readme_inequality(co)
ms = []
import time
x0 = np.zeros((B, n))
for _ in range(8):
    t0 = time.perf_counter()
    out = L.optimize_batched(fam.f, None, fam.d, x0, -inf, inf, 0, 1, history=16)
    ms.append((time.perf_counter() - t0) * 1e3)      # end to end through the host-buffer entry point (pipelined chunks)
print(os.environ.get("LFPSQP_LIB_PATH", "default"), "e2e ms", ["%.3f" % v for v in ms], "-> %.2f M instances/s" % (B / min(ms[1:]) / 1e3),
      "iters", int(out[4]["iter"].sum()), flush=True)
