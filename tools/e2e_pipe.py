"""End-to-end C2 rate through lfpsqp_solve_batched (pinned host buffers) for several pipeline shapes (env knobs of abi.cu):
   python tools/e2e_pipe.py "8:1" "8:0.5" "16:0.5" ...     (chunks:ramp)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bench, lfpsqp.jl_b200 as L
from lfpsqp.jl_b200 import _lib
B, n, H = bench.B_PER_GPU, bench.N_VARS, bench.HIST
coeff, x0 = bench.make_inputs(0, B)
ctx = L.Context(0)
inf = np.inf * np.ones(n); xl = -inf; xu = inf
prm = L.LFPSQPParams(disp=L.off).to_c(); pprm = C.cast(C.pointer(prm), C.c_void_p)
keep = []
def P(shape, dt):
    t = torch.empty(shape, dtype=dt).pin_memory(); keep.append(t); return t.numpy()
h_coeff = P((B, n), torch.float64); h_coeff[:] = coeff
h_x0 = P((B, n), torch.float64); h_x0[:] = x0
h_x = P((B, n), torch.float64); h_obj = P((B, H), torch.float64); h_len = P((B,), torch.int64); h_lam = P((B, 1), torch.float64)
h_term = P((B * 40,), torch.uint8)
def step():
    ctx.check(ctx.lib.lfpsqp_solve_batched(ctx.h, L.families.README_INEQ, n, 0, 1, B, _lib.ptr(h_coeff), n, _lib.ptr(h_x0), _lib.ptr(xl),
                                           _lib.ptr(xu), pprm, _lib.ptr(h_x), _lib.ptr(h_obj), H, _lib.ptr(h_len), _lib.ptr(h_lam),
                                           _lib.ptr(h_term), None))
ref = None
for spec in sys.argv[1:] or ["8:1"]:
    ch, ramp = spec.split(":")
    if ch == "auto":
        os.environ.pop("LFPSQP_PIPE_CHUNKS", None); os.environ.pop("LFPSQP_PIPE_RAMP", None)
    else:
        os.environ["LFPSQP_PIPE_CHUNKS"] = ch; os.environ["LFPSQP_PIPE_RAMP"] = ramp
    for _ in range(3): step()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): step()
        torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / 10)
    if ref is None: ref = h_x.copy()
    same = np.array_equal(ref, h_x)
    print("chunks %s ramp %s: %.3f ms/step -> %.2f M inst/s  (x identical to the first shape: %s)" % (ch, ramp, best * 1e3, B / best / 1e6, same), flush=True)
