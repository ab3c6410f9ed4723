"""Dev probe (run on the GPU box): parity summaries vs the oracle + kernel times for the batched configs."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lfpsqp.jl_b200 as L
from oracle import oracle as O
from tests.parity import compare_batch, fmt

ctx = L.default_context(0)
rng = np.random.default_rng(0)

def show(res):
    print(fmt(res), flush=True)

def detail(gpu, orc, k):
    print("  inst", k, "gpu term", gpu[4][k], "orc term", orc[4][k], "x gpu", gpu[0][k][:4], "x orc", orc[0][k][:4])
    if len(gpu) > 5: print("   gpu stats", gpu[5][k], "\n   orc stats", orc[5][k])

# rosenbrock
B = 4096
x0 = rng.uniform(-2, 2, (B, 2)); x0[0] = 0
fam = L.families.rosenbrock()
gpu = L.optimize_batched(fam.f, x0, history=128, return_stats=True)
orc = O.optimize_batched("rosenbrock", 2, 0, 0, x0, H=128, nthreads=8)
show(compare_batch(gpu, orc, 2, "rosenbrock")); detail(gpu, orc, 0)
bad = np.nonzero((gpu[4]["iter"] != orc[4]["iter"]))[0][:3]
for k in bad: detail(gpu, orc, k)

# readme eq
fam = L.families.readme_equality(50)
r = L.optimize(fam.f, fam.c, np.ones(50), 1, return_stats=True)
print("readme_eq:", r[3], r[2], r[4], "x0", r[0][0], "f", r[1])

# readme ineq
B, n = 512, 50
co = rng.standard_normal((B, n)); inf = np.inf * np.ones(n)
fam = L.families.readme_inequality(co)
gpu = L.optimize_batched(fam.f, None, fam.d, np.zeros((B, n)), -inf, inf, 0, 1, return_stats=True)
orc = O.optimize_batched("readme_ineq", n, 0, 1, np.zeros((B, n)), xl=-inf, xu=inf, fam_params=co, fam_stride=n, nthreads=8)
show(compare_batch(gpu, orc, n, "readme_ineq")); detail(gpu, orc, 0)
bad = np.nonzero((gpu[4]["iter"] != orc[4]["iter"]) | (gpu[4]["condition"] != orc[4]["condition"]))[0][:3]
for k in bad: detail(gpu, orc, k)

# boxquad
for m in (0, 1):
    B, n = 256, 12
    xl = np.r_[-np.inf * np.ones(3), -0.5 * np.ones(3), -np.inf * np.ones(3), -0.3 * np.ones(3)]
    xu = np.r_[np.inf * np.ones(6), 0.4 * np.ones(3), 0.6 * np.ones(3)]
    t = 2 * rng.standard_normal((B, n))
    fam = L.families.boxquad(t, a=np.ones(n) if m else None, b=3.0)
    x0 = np.tile(np.clip(np.zeros(n), xl, xu) + (0.25 if m else 0.0), (B, 1))
    gpu = L.optimize_batched(fam.f, fam.c, x0, xl, xu, m, return_stats=True) if m else L.optimize_batched(fam.f, None, x0, xl, xu, 0, return_stats=True)
    orc = O.optimize_batched("boxquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, fam_stride=fam.params.shape[1], nthreads=8)
    show(compare_batch(gpu, orc, n, "boxquad m=%d" % m)); detail(gpu, orc, 0)
    bad = np.nonzero((gpu[4]["iter"] != orc[4]["iter"]) | (gpu[4]["condition"] != orc[4]["condition"]))[0][:3]
    for k in bad: detail(gpu, orc, k)

# sin
for nr in (False, True):
    B, n, m = 128, 24, 6
    t = rng.standard_normal((B, n)); x0 = np.zeros((B, n))
    fam = L.families.sin_system(n, m, t)
    gpu = L.optimize_batched(fam.f, fam.c, x0, m, L.LFPSQPParams(do_project_retract=not nr), return_stats=True)
    orc = O.optimize_batched("sin", n, m, 0, x0, fam_params=t, fam_stride=n, params=O.default_params(do_project_retract=0 if nr else 1), nthreads=8)
    show(compare_batch(gpu, orc, n, "sin nr=%s" % nr)); detail(gpu, orc, 0)
    bad = np.nonzero((gpu[4]["iter"] != orc[4]["iter"]) | (gpu[4]["condition"] != orc[4]["condition"]))[0][:3]
    for k in bad: detail(gpu, orc, k)

# thomson
B, npts = 64, 12
x0 = rng.standard_normal((B, npts, 3)); x0 /= np.linalg.norm(x0, axis=2, keepdims=True); x0 = x0.reshape(B, -1)
fam = L.families.thomson(npts)
gpu = L.optimize_batched(fam.f, fam.c, x0, npts, return_stats=True)
orc = O.optimize_batched("thomson", 3 * npts, npts, 0, x0, nthreads=8)
show(compare_batch(gpu, orc, 3 * npts, "thomson12")); detail(gpu, orc, 0)

# diagquad
B, n, m = 64, 64, 4
Q = rng.standard_normal((m, n)) / np.sqrt(n); A = rng.standard_normal((m, n)) / np.sqrt(n)
xt = rng.standard_normal(n); w = np.exp(rng.uniform(0, np.log(100.0), n)); xs = rng.standard_normal(n)
b = 0.5 * Q @ (xs * xs) + A @ xs
x0 = np.tile(xs, (B, 1))
fam = L.families.diagquad(Q, A, b, xt, w)
gpu = L.optimize_batched(fam.f, fam.c, x0, m, return_stats=True)
orc = O.optimize_batched("diagquad", n, m, 0, x0, fam_params=fam.params, fam_stride=0, nthreads=8)
show(compare_batch(gpu, orc, n, "diagquad")); detail(gpu, orc, 0)

# timings
print("--- timings")
B = 1 << 20
x0 = rng.uniform(-2, 2, (B, 2)); x0[0] = 0
fam = L.families.rosenbrock()
for rep in range(3):
    t0 = time.time(); out = L.optimize_batched(fam.f, x0, history=64); t1 = time.time()
    print("C3 rosenbrock 1M: kernel %.3f ms, e2e %.1f ms -> %.3g inst/s (kernel)" % (ctx.last_kernel_ms, (t1 - t0) * 1e3, B / ctx.last_kernel_ms * 1e3), flush=True)
print("  iters mean", out[4]["iter"].mean(), "max", out[4]["iter"].max())
B, n = 65536, 50
co = rng.standard_normal((B, n))
fam = L.families.readme_inequality(co)
for rep in range(3):
    t0 = time.time(); out = L.optimize_batched(fam.f, None, fam.d, np.zeros((B, n)), -inf50 if False else -np.inf*np.ones(n), np.inf*np.ones(n), 0, 1, history=16); t1 = time.time()
    print("C2 readme_ineq 65536: kernel %.3f ms, e2e %.1f ms -> %.3g inst/s (kernel)" % (ctx.last_kernel_ms, (t1 - t0) * 1e3, B / ctx.last_kernel_ms * 1e3), flush=True)
print("  iters mean", out[4]["iter"].mean(), "max", out[4]["iter"].max(), "cond counts", np.bincount(out[4]["condition"]))
t0 = time.time()
orc = O.optimize_batched("readme_ineq", n, 0, 1, np.zeros((4096, n)), xl=-np.inf*np.ones(n), xu=np.inf*np.ones(n), fam_params=co[:4096], fam_stride=n, nthreads=os.cpu_count())
t1 = time.time()
print("CPU oracle C2: %d threads, %.1f inst/s" % (os.cpu_count(), 4096 / (t1 - t0)))
