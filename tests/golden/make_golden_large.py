"""Generates tests/golden/large_midsize.npz: the CPU oracle's results (three builds: plain / fma / seq summation) on the
mid-size BASELINE C4 / C5 family problems of tests/test_gpu_parity_fullsize.py.  One SVD-based oracle solve of the
8192 x 512 case takes minutes on the host, so the GPU test compares against these committed vectors instead of
re-running the oracle on the GPU box.  Inputs are regenerated from the seeds by the test itself.

    python tests/golden/make_golden_large.py
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

large = importlib.import_module("lfpsqp.jl_b200.large")
CASES = {  # name -> (family, builder)
    "thomson512": ("thomson", 512),
    "diagquad_8192x512": ("diagquad", (8192, 512, 1, 100.0)),
}


def inputs(name):
    fam, spec = CASES[name]
    if fam == "thomson":
        npts = spec
        rng = np.random.Generator(np.random.Philox(key=4))          # bench.py's C4 seed (SEED + 4)
        p0 = rng.standard_normal((npts, 3)); p0 /= np.linalg.norm(p0, axis=1, keepdims=True)
        return fam, 3 * npts, npts, p0.ravel(), None
    n, m, seed, cond = spec
    Q, A, b, xt, w, x0 = large.make_diagquad(n, m, seed=seed, cond=cond)
    return fam, n, m, x0, np.concatenate([Q.ravel(), A.ravel(), b, xt, w])


if __name__ == "__main__":
    out = {}
    for name in CASES:
        fam, n, m, x0, blob = inputs(name)
        for v in ("base", "fma", "seq"):
            t0 = time.time()
            if v == "base":
                r = O.optimize(fam, n, m, 0, x0, fam_params=blob)
            else:
                with O.variant(v):
                    r = O.optimize(fam, n, m, 0, x0, fam_params=blob)
            x, obj, lam, term, st = r
            print(name, v, "%.1f s" % (time.time() - t0), term, st, flush=True)
            out["%s/%s/x" % (name, v)] = x
            out["%s/%s/obj" % (name, v)] = obj
            out["%s/%s/lam" % (name, v)] = lam
            out["%s/%s/term" % (name, v)] = np.array([term["condition"], term["iter"]], dtype=np.int64)
            out["%s/%s/stats" % (name, v)] = np.array([st[k] for k in ("projcg_iters", "projcg_negcurv", "armijo_trials", "retract_outer",
                                                                        "retract_pcg", "pp_backtracks", "newton_accepted", "svd_calls", "f_evals")], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "large_midsize.npz"), **out)
