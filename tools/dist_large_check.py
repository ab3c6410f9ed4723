"""Multi-GPU check (torchrun, one rank per GPU): column-sharded large-n solve vs the CPU oracle + C5 projcg timing.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_large_check.py [c5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lfpsqp.jl_b200 as L
from lfpsqp.jl_b200 import dist as D

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = L.Context(lr)
D.init_comm(ctx, dist, peer=("nopeer" not in sys.argv))
if rank == 0: print("comm mode", ctx.lib.lfpsqp_comm_mode(ctx.h), flush=True)
dev = torch.device("cuda", lr)

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

# ---- parity: sharded solve == oracle (and == what one GPU computes)
for (n, m, nr) in [(4096, 200, False), (1000, 130, False), (4096, 200, True)]:
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=1, cond=100.0)
    col0, nloc = D.column_range(n, world, rank)
    blob = D.shard_diagquad(Q, A, b, xt, w, col0, nloc)
    fam = L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0, blob)
    P = L.LargeProblem(fam, ctx, col0=col0, n_loc=nloc, n_global=n)
    prm = L.LFPSQPParams(do_project_retract=not nr)
    x, obj, lam, info, st, status = P.solve(x0[col0:col0 + nloc], prm, return_stats=True)
    parts = [None] * world
    dist.all_gather_object(parts, (col0, x))
    if rank == 0:
        from oracle import oracle as O
        xfull = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        ox, oobj, olam, ot, ost = O.optimize("diagquad", n, m, 0, x0, fam_params=np.concatenate([Q.ravel(), A.ravel(), b, xt, w]),
                                             params=O.default_params(do_project_retract=0 if nr else 1))
        ok = (int(info.condition) == ot["condition"] and abs(info.iter - ot["iter"]) <= 1 and rel(xfull, ox) <= 1e-8
              and abs(obj[-1] - oobj[-1]) <= 1e-10 * abs(oobj[-1]))
        print("world=%d n=%d m=%d nr=%s: gpu %s %d orc %d %d  x err %.2e f err %.2e lam err %.2e status %d -> %s" % (
            world, n, m, nr, info.condition.name, info.iter, ot["condition"], ot["iter"], rel(xfull, ox),
            abs(obj[-1] - oobj[-1]) / abs(oobj[-1]), rel(lam, olam), status, "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()

# ---- Thomson column-sharded: whole points per rank, x / v all-gathered per callback evaluation
for npts in (64, 128):
    rng = np.random.default_rng(6)
    x0 = rng.standard_normal((npts, 3)); x0 /= np.linalg.norm(x0, axis=1, keepdims=True); x0 = x0.ravel()
    col0, nloc = D.thomson_point_range(npts, world, rank)
    P = L.LargeProblem(L.families.thomson(npts), ctx, col0=col0, n_loc=nloc, n_global=3 * npts)
    x, obj, lam, info, st, status = P.solve(x0[col0:col0 + nloc], L.LFPSQPParams(), return_stats=True)
    parts = [None] * world
    dist.all_gather_object(parts, (col0, x))
    if rank == 0:
        from oracle import oracle as O
        xfull = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        ox, oobj, olam, ot, ost = O.optimize("thomson", 3 * npts, npts, 0, x0)
        with O.variant("fma"):
            fx = O.optimize("thomson", 3 * npts, npts, 0, x0)[0]
        ok = (int(info.condition) == ot["condition"] and abs(info.iter - ot["iter"]) <= 1 and
              (rel(xfull, ox) <= 1e-8 or rel(xfull, ox) <= 10 * rel(fx, ox)) and abs(obj[-1] - oobj[-1]) <= 1e-9 * abs(oobj[-1]))
        print("world=%d THOMSON N=%d: gpu %s %d orc %d %d  x err %.2e (oracle-fma %.2e) f err %.2e lam err %.2e status %d -> %s" % (
            world, npts, info.condition.name, info.iter, ot["condition"], ot["iter"], rel(xfull, ox), rel(fx, ox),
            abs(obj[-1] - oobj[-1]) / abs(oobj[-1]), rel(lam, olam), status, "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()

if "c4" in sys.argv:   # BASELINE config C4 (Thomson N = 4096) column-sharded: first 10 outer iterations, wall + phases
    import time, warnings
    npts = 4096
    rng = np.random.Generator(np.random.Philox(key=4))
    p0 = rng.standard_normal((npts, 3)); p0 /= np.linalg.norm(p0, axis=1, keepdims=True); p0 = p0.ravel()
    col0, nloc = D.thomson_point_range(npts, world, rank)
    P = L.LargeProblem(L.families.thomson(npts), ctx, col0=col0, n_loc=nloc, n_global=3 * npts)
    for rep in range(2):
        dist.barrier()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            x, obj, lam, info, st, status = P.solve(p0[col0:col0 + nloc], L.LFPSQPParams(maxiter=10), return_stats=True)
            wall = time.perf_counter() - t0
        ph = P.phase_ms()
        if rank == 0:
            print("C4 world=%d: 10 outer iterations %.1f ms wall | factor %.1f projcg %.1f linesearch %.1f | projcg iters %d | f %.6f -> %.6f" % (
                world, wall * 1e3, ph["factor"], ph["projcg"], ph["linesearch"], st["projcg_iters"], obj[0], obj[-1]), flush=True)

# ---- finite bounds (2n embedding) column-sharded: every rank passes its slice of xl, xu
for (n, m, seed, nr) in [(256, 16, 2, False), (256, 16, 2, True), (512, 32, 1, True)]:
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    rng = np.random.default_rng(seed + 10)
    xl = x0 - rng.uniform(0.05, 1.0, n); xu = x0 + rng.uniform(0.05, 1.0, n)
    col0, nloc = D.column_range(n, world, rank)
    fam = L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0, D.shard_diagquad(Q, A, b, xt, w, col0, nloc))
    P = L.LargeProblem(fam, ctx, col0=col0, n_loc=nloc, n_global=n)
    P.set_bounds(xl[col0:col0 + nloc], xu[col0:col0 + nloc])
    x, obj, lam, info, st, status = P.solve(x0[col0:col0 + nloc], L.LFPSQPParams(do_project_retract=not nr), return_stats=True)
    parts = [None] * world
    dist.all_gather_object(parts, (col0, x))
    if rank == 0:
        from oracle import oracle as O
        xfull = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        blob = np.concatenate([Q.ravel(), A.ravel(), b, xt, w])
        op = O.default_params(do_project_retract=0 if nr else 1)
        ox, oobj, olam, ot, ost = O.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=blob, params=op)
        with O.variant("fma"):
            fx = O.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=blob, params=op)[0]
        print("world=%d BOUNDS n=%d m=%d nr=%s: gpu %s %d orc %d %d  x err %.2e (oracle-fma %.2e) f err %.2e status %d" % (
            world, n, m, nr, info.condition.name, info.iter, ot["condition"], ot["iter"], rel(xfull, ox), rel(fx, ox),
            abs(obj[-1] - oobj[-1]) / abs(oobj[-1]), status), flush=True)
    dist.barrier()

if "c5" in sys.argv:
    n, m, K = 65536, 2048, 64
    col0, nloc = D.column_range(n, world, rank)
    g = torch.Generator(device=dev); g.manual_seed(0)   # same stream on every rank: generate the full rows, keep the shard
    def shard_rand():
        full = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
        return full[:, col0:col0 + nloc].contiguous()
    Q = shard_rand(); A = shard_rand()
    x0 = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    xt = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    w = torch.exp(torch.rand(n, dtype=torch.float64, device=dev, generator=g) * np.log(1e4))
    x0l = x0[col0:col0 + nloc]
    bpart = 0.5 * Q @ (x0l * x0l) + A @ x0l
    dist.all_reduce(bpart)
    blob = torch.cat([Q.reshape(-1), A.reshape(-1), bpart, xt[col0:col0 + nloc], w[col0:col0 + nloc]]).contiguous()
    del Q, A
    torch.cuda.synchronize()
    fam = L.families.Family(L.families.DIAGQUAD, "diagquad", n, m, 0)
    P = L.LargeProblem(fam, ctx, col0=col0, n_loc=nloc, n_global=n, params_dev_ptr=blob.data_ptr())
    x0h = x0l.cpu().numpy()
    fac = P.factor(x0h, want=())
    for rep in range(3):
        dist.barrier()
        r = P.projcg(x0h, lam=np.zeros(m), tol=0.0, maxit=K, chunk=16, want_solution=False)
        t = torch.tensor([r["ms"]], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            per = float(t.item()) / K
            print("C5 world=%d: gram %.2f ms (local K=%d), projcg %.3f ms/iter -> %.1f it/s (iters %d status %d)" % (
                world, fac["gram_ms"], nloc, per, 1e3 / per, r["iters"], r["status"]), flush=True)
dist.barrier()
dist.destroy_process_group()
