"""Parameter coverage and edge cases of the driver (src/LFPSQP.jl:57-81, src/optimize.jl:347-359, :364-390, :415-420)
through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

from tests.parity import compare_batch, fmt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lfpsqp.jl_b200 as L
    L.default_context(0)
    return L


CASES = [
    dict(maxiter=0), dict(maxiter=3), dict(eps_x=1e-3), dict(eps_f=1e-10, eps_kkt=1e-3), dict(do_newton=False, maxiter=200),
    dict(disable_linesearch=True, maxiter=50), dict(alpha=0.5, s=0.25, sigma=1e-2), dict(tn_kappa=0.1, tn_maxiter=3),
    dict(mu0=1e-1, eps_c=1e-8), dict(maxiter_pcg=2), dict(maxiter_retract=2),
]


@pytest.mark.parametrize("kw", CASES, ids=[",".join("%s=%s" % kv for kv in c.items()) for c in CASES])
def test_parameter_coverage_vs_oracle(L, oracle, kw):
    rng = np.random.default_rng(41)
    prm = L.LFPSQPParams(**kw)
    oprm = oracle.default_params(**{k: (int(v) if isinstance(v, bool) else v) for k, v in kw.items()})
    # unconstrained (thread-per-instance kernel)
    B = 256
    x0 = rng.uniform(-2, 2, (B, 2)); x0[0] = 0.0
    gpu = L.optimize_batched(L.families.rosenbrock().f, x0, prm, history=256)
    orc = oracle.optimize_batched("rosenbrock", 2, 0, 0, x0, params=oprm, H=256, nthreads=8)
    res = compare_batch(gpu, orc, 2, "rosenbrock %s" % kw)
    print(fmt(res))
    assert res["cond_frac"] >= 0.99 and res["iter_pm1_frac"] >= 0.99 and res["x_frac"] >= 0.97
    # slack + bounds (register-resident kernel) and equalities (shared-memory kernel)
    B, n = 64, 50
    co = rng.standard_normal((B, n)); inf = np.inf * np.ones(n)
    fam = L.families.readme_inequality(co)
    gpu = L.optimize_batched(fam.f, None, fam.d, np.zeros((B, n)), -inf, inf, 0, 1, prm)
    orc = oracle.optimize_batched("readme_ineq", n, 0, 1, np.zeros((B, n)), xl=-inf, xu=inf, fam_params=co, fam_stride=n,
                                  params=oprm, nthreads=8)
    res = compare_batch(gpu, orc, n, "readme_ineq %s" % kw)
    print(fmt(res))
    assert res["cond_frac"] >= 0.97 and res["iter_pm1_frac"] >= 0.97 and res["x_frac"] >= 0.95
    B, n, m = 32, 24, 6
    t = rng.standard_normal((B, n))
    fam = L.families.sin_system(n, m, t)
    gpu = L.optimize_batched(fam.f, fam.c, np.zeros((B, n)), m, prm)
    orc = oracle.optimize_batched("sin", n, m, 0, np.zeros((B, n)), fam_params=t, fam_stride=n, params=oprm, nthreads=8)
    res = compare_batch(gpu, orc, n, "sin %s" % kw)
    print(fmt(res))
    assert res["cond_frac"] >= 0.96 and res["iter_pm1_frac"] >= 0.96 and res["x_frac"] >= 0.9


def test_edge_cases(L, oracle):
    fam = L.families.rosenbrock()
    # empty batch
    x, obj, olen, lam, term = L.optimize_batched(fam.f, np.zeros((0, 2)))
    assert x.shape == (0, 2) and term.shape == (0,)
    # history shorter than the iteration count: first H values kept, obj_len reports the true length (iters + 1)
    x, obj, olen, lam, term = L.optimize_batched(fam.f, np.zeros((3, 2)), history=4)
    assert np.all(olen == 18) and obj.shape == (3, 4) and np.all(term["iter"] == 17)
    full = L.optimize(fam.f, np.zeros(2))[1]
    assert np.array_equal(obj[0], full[:4])
    # n = 1 with both bounds (circle embedding), start on a bound
    f1 = L.families.boxquad(np.array([[3.0], [-3.0], [0.2]]))
    gpu = L.optimize_batched(f1.f, None, np.array([[1.0], [-1.0], [0.0]]), np.array([-1.0]), np.array([1.0]), 0)
    orc = oracle.optimize_batched("boxquad", 1, 0, 0, np.array([[1.0], [-1.0], [0.0]]), xl=np.array([-1.0]), xu=np.array([1.0]),
                                  fam_params=f1.params, fam_stride=3, nthreads=1)
    assert np.array_equal(gpu[4]["condition"], orc[4]["condition"])
    assert np.allclose(gpu[0], orc[0], atol=1e-7) and np.allclose(gpu[0].ravel(), [1.0, -1.0, 0.2], atol=1e-3)
    # large-n mode through the one-call entry point: finite bounds run the 2n embedding; xl > xu mirrors optimize.jl:160-162
    import ctypes as C
    from lfpsqp.jl_b200 import _lib
    ctx = L.default_context(0)
    Q, A, b, xt, w, x0 = L.make_diagquad(64, 4, seed=1, cond=50.0)
    blob = np.concatenate([Q.ravel(), A.ravel(), b, xt, w])
    prm = L.LFPSQPParams().to_c()
    out = np.zeros(64); obj = np.zeros(64); ol = np.zeros(1, dtype=np.int64); lam = np.zeros(4); term = np.zeros(1, dtype=_lib.TERM_DTYPE)
    lo, hi = x0 - 0.5, x0 + 0.5     # keep the host buffers alive across the calls
    rc = ctx.lib.lfpsqp_solve_large(ctx.h, L.families.DIAGQUAD, 64, 4, _lib.ptr(blob), _lib.ptr(x0), _lib.ptr(lo), _lib.ptr(hi),
                                    C.cast(C.pointer(prm), C.c_void_p), _lib.ptr(out), _lib.ptr(obj), 64, _lib.ptr(ol), _lib.ptr(lam), _lib.ptr(term), None)
    assert rc == 0 and term[0]["status"] == 0 and np.all(out >= lo - 2e-5) and np.all(out <= hi + 2e-5)
    ox = oracle.optimize("diagquad", 64, 4, 0, x0, xl=lo, xu=hi, fam_params=blob)
    with oracle.variant("fma"):   # bound-active problems are rounding-sensitive in the reference algorithm itself
        of = oracle.optimize("diagquad", 64, 4, 0, x0, xl=lo, xu=hi, fam_params=blob)
    sens = np.linalg.norm(of[0] - ox[0]) / np.linalg.norm(ox[0])
    assert term[0]["condition"] == ox[3]["condition"] and abs(int(term[0]["iter"]) - ox[3]["iter"]) <= 1 + abs(of[3]["iter"] - ox[3]["iter"])
    assert np.linalg.norm(out - ox[0]) <= max(1e-8, 10 * sens) * np.linalg.norm(ox[0])
    rc = ctx.lib.lfpsqp_solve_large(ctx.h, L.families.DIAGQUAD, 64, 4, _lib.ptr(blob), _lib.ptr(x0), _lib.ptr(hi), _lib.ptr(lo),
                                    C.cast(C.pointer(prm), C.c_void_p), _lib.ptr(out), _lib.ptr(obj), 64, _lib.ptr(ol), _lib.ptr(lam), _lib.ptr(term), None)
    assert rc == -2
