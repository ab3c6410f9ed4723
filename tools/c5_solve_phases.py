"""C5 (n = 65536, m = 2048) whole-driver solve, first outer iterations: wall time and phase breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, lfpsqp.jl_b200 as L
n, m = 65536, 2048
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
Q = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
A = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
x0 = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
xt = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
w = torch.exp(torch.rand(n, dtype=torch.float64, device=dev, generator=g) * np.log(1e2))
b = 0.5 * Q @ (x0 * x0) + A @ x0
blob = torch.cat([Q.reshape(-1), A.reshape(-1), b, xt, w]).contiguous()
del Q, A
P = L.LargeProblem(L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0), params_dev_ptr=blob.data_ptr())
x0h = x0.cpu().numpy()
for rep in range(2):
    t0 = time.perf_counter()
    x, obj, lam, info, st, status = P.solve(x0h, L.LFPSQPParams(maxiter=int(sys.argv[1]) if len(sys.argv) > 1 else 6), return_stats=True)
    wall = time.perf_counter() - t0
    print("C5 solve: %d outer iterations %.1f ms wall | %s | %s | f %.6e -> %.6e" % (info.iter, wall * 1e3, P.phase_ms(), st, obj[0], obj[-1]), flush=True)
