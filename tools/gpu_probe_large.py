"""Dev probe (GPU box): large-n mode parity vs numpy/oracle at small sizes + timings at BASELINE sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
from oracle import oracle as O

ctx = L.default_context(0)
which = sys.argv[1:] or ["small", "thomson", "c5"]

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

if "small" in which:
    for (n, m) in [(512, 32), (4096, 200), (1000, 130)]:
        Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=1, cond=100.0)
        fam = L.families.diagquad(Q, A, b, xt, w)
        P = L.LargeProblem(fam, ctx)
        J = Q * x0[None, :] + A
        fac = P.factor(x0)
        G = J @ J.T; Lc = np.linalg.cholesky(G)
        print("n=%d m=%d  G err %.2e  L err %.2e  Linv err %.2e rankdef %d gram_ms %.3f" % (
            n, m, rel(np.tril(fac["G"]), np.tril(G)), rel(np.tril(fac["L"]), Lc), rel(fac["Linv"], np.linalg.inv(Lc)), fac["rank_deficient"], fac["gram_ms"]), flush=True)
        v = np.random.default_rng(3).standard_normal(n)
        pv, lam = P.project(v)
        u = np.linalg.solve(G, J @ v)
        print("   project err %.2e  lam err %.2e  |J pv| %.2e" % (rel(pv, v - J.T @ u), rel(lam, u), np.linalg.norm(J @ pv)))
        # projcg vs oracle-level dense reference: solve with numpy
        r = P.projcg(x0, lam=np.zeros(m), tol=1e-10, maxit=500)
        print("   projcg iters %d nr %.2e status %d ms %.3f" % (r["iters"], r["nr"], r["status"], r["ms"]))
        t0 = time.time()
        x, obj, lamk, info, st, status = P.solve(x0, L.LFPSQPParams(), return_stats=True)
        t1 = time.time()
        ox, oobj, olam, ot, ost = O.optimize("diagquad", n, m, 0, x0, fam_params=fam.params)
        print("   solve: gpu", info.condition.name, info.iter, "orc", ot["condition"], ot["iter"], "x err %.2e f err %.2e lam err %.2e status %d  (%.2f s, launches %d)" % (
            rel(x, ox), abs(obj[-1] - oobj[-1]) / abs(oobj[-1]), rel(lamk, olam), status, t1 - t0, ctx.last_launches))
        print("   gpu stats", st, "\n   orc stats", {k: ost[k] for k in ost if k != "flops"}, flush=True)
        x, obj, lamk, info, st, status = P.solve(x0, L.LFPSQPParams(do_project_retract=False), return_stats=True)
        ox, oobj, olam, ot, ost = O.optimize("diagquad", n, m, 0, x0, fam_params=fam.params, params=O.default_params(do_project_retract=0))
        print("   solve NR: gpu", info.condition.name, info.iter, "orc", ot["condition"], ot["iter"], "x err %.2e f err %.2e" % (rel(x, ox), abs(obj[-1] - oobj[-1]) / abs(oobj[-1])), st["retract_outer"], ost["retract_outer"], flush=True)

if "thomson" in which:
    for npts in (32, 100):
        rng = np.random.default_rng(6)
        x0 = rng.standard_normal((npts, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
        fam = L.families.thomson(npts)
        P = L.LargeProblem(fam, ctx)
        t0 = time.time()
        x, obj, lamk, info, st, status = P.solve(x0, L.LFPSQPParams(), return_stats=True)
        t1 = time.time()
        ox, oobj, olam, ot, ost = O.optimize("thomson", 3 * npts, npts, 0, x0)
        print("thomson %d: gpu" % npts, info.condition.name, info.iter, "orc", ot["condition"], ot["iter"], "x err %.2e f err %.2e lam err %.2e (%.2f s)" % (
            rel(x, ox), abs(obj[-1] - oobj[-1]) / abs(oobj[-1]), rel(lamk, olam), t1 - t0))
        print("   gpu stats", st, "\n   orc stats", {k: ost[k] for k in ost if k != "flops"}, flush=True)

if "c5" in which:
    import torch
    n, m = 65536, 2048
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
    fam = L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0)
    P = L.LargeProblem(fam, ctx, params_dev_ptr=blob.data_ptr())
    x0h = x0.cpu().numpy()
    for rep in range(2):
        t0 = time.time(); fac = P.factor(x0h, want=()); t1 = time.time()
        print("C5 factor: gram %.3f ms (%.2f TFLOP/s), total factor wall %.1f ms rankdef %d" % (
            fac["gram_ms"], m * (m + 1) * n / fac["gram_ms"] / 1e9, (t1 - t0) * 1e3, fac["rank_deficient"]), flush=True)
    for chunk in (1, 4, 16):
        for rep in range(2):
            r = P.projcg(x0h, lam=np.zeros(m), tol=0.0, maxit=48, chunk=chunk, want_solution=False)
            per = r["ms"] / max(r["iters"], 1)
            bytes_it = 16.0 * m * n + 8.0 * m * m + 104.0 * n
            print("C5 projcg chunk=%d: %d iters status %d %.3f ms/iter -> %.1f it/s, %.1f GB/s (%.1f%% of 6543)" % (
                chunk, r["iters"], r["status"], per, 1e3 / per, bytes_it / per / 1e6, bytes_it / per / 1e6 / 65.434), flush=True)
    t0 = time.time()
    x, obj, lamk, info, st, status = P.solve(x0h, L.LFPSQPParams(), return_stats=True)
    t1 = time.time()
    print("C5 full solve:", info, "status", status, "%.2f s" % (t1 - t0), st, "f:", obj[:3], obj[-1], "launches", ctx.last_launches, flush=True)

if "c4" in which:
    npts = 4096
    n, m = 3 * npts, npts
    rng = np.random.Generator(np.random.Philox(key=4))
    x0 = rng.standard_normal((npts, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
    P = L.LargeProblem(L.families.thomson(npts), ctx)
    for rep in range(2):
        t0 = time.time(); fac = P.factor(x0, want=()); t1 = time.time()
        print("C4 factor: gram %.3f ms (%.2f TFLOP/s), total factor wall %.1f ms rankdef %d" % (
            fac["gram_ms"], m * (m + 1.0) * n / fac["gram_ms"] / 1e9, (t1 - t0) * 1e3, fac["rank_deficient"]), flush=True)
    lam0 = np.full(m, 50.0)   # positive multipliers keep the projected Hessian positive: no negative-curvature exit
    for chunk in (4, 16):
        r = P.projcg(x0, lam=lam0, tol=0.0, maxit=48, chunk=chunk, want_solution=False)
        per = r["ms"] / max(r["iters"], 1)
        bytes_it = 16.0 * m * n + 8.0 * m * m + 104.0 * n
        print("C4 projcg chunk=%d: %d iters status %d %.3f ms/iter -> %.1f it/s, %.1f GB/s (%.1f%% of 6543)" % (
            chunk, r["iters"], r["status"], per, 1e3 / per, bytes_it / per / 1e6, bytes_it / per / 1e6 / 65.434), flush=True)
    t0 = time.time()
    x, obj, lamk, info, st, status = P.solve(x0, L.LFPSQPParams(maxiter=20), return_stats=True)
    t1 = time.time()
    print("C4 20 outer iterations:", info, "status", status, "%.2f s" % (t1 - t0), st, "f:", obj[:2], obj[-1], "launches", ctx.last_launches, flush=True)
