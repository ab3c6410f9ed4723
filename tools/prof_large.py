"""Profiling driver (run under ncu on the GPU box): C5 factorisation once + a few projcg iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lfpsqp.jl_b200 as L

n, m = 65536, 2048
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = L.default_context(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
Q = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
A = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
x0 = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
xt = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
w = torch.exp(torch.rand(n, dtype=torch.float64, device=dev, generator=g) * np.log(1e4))
b = 0.5 * Q @ (x0 * x0) + A @ x0
blob = torch.cat([Q.reshape(-1), A.reshape(-1), b, xt, w]).contiguous()
del Q, A
torch.cuda.synchronize()
P = L.LargeProblem(L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0), ctx, params_dev_ptr=blob.data_ptr())
x0h = x0.cpu().numpy()
fac = P.factor(x0h, want=())
r = P.projcg(x0h, lam=np.zeros(m), tol=0.0, maxit=K, chunk=K, want_solution=False)
print("gram_ms", fac["gram_ms"], "projcg", r)
