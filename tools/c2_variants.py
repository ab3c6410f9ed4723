"""C2 batched kernel: device-resident kernel time (CUDA events) and full-size parity for every register-kernel variant the
library was built with (make EXTRA=-DLFPSQP_REG_EXPERIMENTS).  Usage: python tools/c2_variants.py [lw codes ...]"""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import lfpsqp.jl_b200 as L
from lfpsqp.jl_b200 import _lib
from oracle import oracle as O
from tests.parity import compare_batch, fmt

codes = sys.argv[1:] or ["32", "0"]
dev = torch.device("cuda", 0)
ctx = L.Context(0)
B, n, H = bench.B_PER_GPU, bench.N_VARS, bench.HIST
coeff, x0 = bench.make_inputs(0, B)
inf = np.inf * np.ones(n)
orc = O.optimize_batched("readme_ineq", n, 0, 1, x0, xl=-inf, xu=inf, fam_params=coeff, fam_stride=n, H=H, nthreads=bench.host_cores())
prm = L.LFPSQPParams(disp=L.off).to_c(); pprm = C.cast(C.pointer(prm), C.c_void_p)
d_coeff = torch.from_numpy(coeff).to(dev); d_x0 = torch.from_numpy(x0).to(dev)
d_x = torch.empty((B, n), dtype=torch.float64, device=dev); d_obj = torch.empty((B, H), dtype=torch.float64, device=dev)
d_len = torch.empty(B, dtype=torch.int64, device=dev); d_lam = torch.empty((B, 1), dtype=torch.float64, device=dev)
d_term = torch.empty(B * 40, dtype=torch.uint8, device=dev)
for code in codes:
    os.environ["LFPSQP_REG_LW"] = code
    ms = []
    for k in range(8):
        d_obj.fill_(float("nan"))
        rc = ctx.lib.lfpsqp_solve_batched_dev(ctx.h, L.families.README_INEQ, n, 0, 1, B, d_coeff.data_ptr(), n, d_x0.data_ptr(), _lib.ptr(-inf),
                                              _lib.ptr(inf), pprm, d_x.data_ptr(), d_obj.data_ptr(), H, d_len.data_ptr(), d_lam.data_ptr(),
                                              d_term.data_ptr(), None)
        ctx.check(rc)
        ms.append(ctx.last_kernel_ms)
    term = np.frombuffer(d_term.cpu().numpy().tobytes(), dtype=_lib.TERM_DTYPE)
    gpu = (d_x.cpu().numpy(), d_obj.cpu().numpy(), d_len.cpu().numpy(), d_lam.cpu().numpy(), term)
    res = compare_batch(gpu, orc, n, "variant %s" % code)
    best = min(ms[2:])
    print(json.dumps({"variant": code, "kernel_ms": best, "Minst_per_s": B / best / 1e3, "all": ms}), flush=True)
    print(fmt(res), flush=True)
