"""GPU parity tests: large-n mode with finite bounds (the 2n embedding of src/inequality_helper.jl) and the host-callback
family (explicit-derivative core, src/optimize.jl:119) vs the CPU oracle on the same seeded inputs, through the C ABI.

Bound-active problems are rounding-sensitive IN THE REFERENCE ALGORITHM ITSELF (two builds of the oracle, with and
without FMA contraction, already differ by 1e-5 in x on some of these cases), so the bar is the north-star tolerance
(x 1e-8, objective 1e-10, relative) or "within 10x of oracle-vs-oracle(+FMA)", whichever is larger; condition and
iteration count keep the +-1 rule unless the two oracle builds themselves disagree by more."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lfpsqp.jl_b200 as L
    L.default_context(0)
    return L


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def bounds_for(x0, seed, kind):
    rng = np.random.default_rng(seed)
    n = x0.size
    xl = x0 - rng.uniform(0.05, 1.0, n); xu = x0 + rng.uniform(0.05, 1.0, n)
    if kind == "mixed":   # all four InequalityData classes (inequality_helper.jl:54-82): line, lower, upper, circle
        k = rng.integers(0, 4, n)
        xl[k == 0] = -np.inf; xu[k == 0] = np.inf
        xu[k == 1] = np.inf
        xl[k == 2] = -np.inf
    return xl, xu


def check(got, orc, orc_fma):
    x, obj, lam, info, st, status = got
    ox, oobj, olam, ot, ost = orc
    fx, fobj, flam, ft, _ = orc_fma
    sens_x = rel(fx, ox); sens_it = abs(ft["iter"] - ot["iter"])
    sens_f = abs(fobj[-1] - oobj[-1]) / abs(oobj[-1])
    assert status == 0
    assert int(info.condition) == ot["condition"]
    assert abs(info.iter - ot["iter"]) <= max(1, sens_it + 1)
    assert rel(x, ox) <= max(1e-8, 10 * sens_x), (rel(x, ox), sens_x)
    assert abs(obj[-1] - oobj[-1]) <= max(1e-10, 10 * sens_f) * abs(oobj[-1]), (obj[-1], oobj[-1], sens_f)
    assert len(obj) == info.iter + 1


# (n, m, seed, kind, nr): cases on which the reference algorithm terminates (checked with the oracle; with NR and many
# active bounds it can fail at every alpha, and linesearch.jl:57-60 has no lower bound on alpha)
CASES = [(64, 4, 5, "mixed", 0), (64, 4, 5, "mixed", 1), (64, 4, 5, "box", 0), (64, 4, 5, "box", 1),
         (256, 16, 2, "mixed", 0), (256, 16, 2, "mixed", 1), (256, 16, 2, "box", 0), (256, 16, 2, "box", 1),
         (512, 32, 1, "box", 0), (512, 32, 1, "box", 1), (1000, 130, 3, "box", 1),
         (300, 0, 4, "mixed", 0), (300, 0, 4, "box", 0)]


@pytest.mark.parametrize("n,m,seed,kind,nr", CASES)
def test_diagquad_bounds_vs_oracle(L, oracle, n, m, seed, kind, nr):
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    xl, xu = bounds_for(x0, seed + 10, kind)
    P = L.LargeProblem(fam)
    P.set_bounds(xl, xu)
    got = P.solve(x0, L.LFPSQPParams(do_project_retract=not nr), return_stats=True)
    op = oracle.default_params(do_project_retract=0 if nr else 1)
    orc = oracle.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, params=op)
    with oracle.variant("fma"):
        orc_fma = oracle.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, params=op)
    check(got, orc, orc_fma)
    x = got[0]
    # the embedding holds h(x, y) = 0 to eps_c = 1e-6 (retractions.jl:359): on a circle of radius rho that is a bound
    # violation of up to eps_c / (2 rho), rho >= 0.05 here
    assert np.all(x >= xl - 2e-5) and np.all(x <= xu + 2e-5)
    if m > 0:
        assert np.max(np.abs(0.5 * Q @ (x * x) + A @ x - b)) < 1e-5


def test_bounds_all_infinite_is_unbounded_path(L):
    # optimize.jl:151: all xl = -Inf and xu = +Inf => ineq = false, bit-identical to passing nothing
    n, m = 256, 16
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=2, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    P = L.LargeProblem(fam)
    a = P.solve(x0)
    P.set_bounds(-np.inf * np.ones(n), np.inf * np.ones(n))
    b2 = P.solve(x0)
    assert np.array_equal(a[0], b2[0]) and a[3] == b2[3]
    # bounds can be switched on, off and on again on the same bound problem (workspaces are kept)
    xl, xu = bounds_for(x0, 12, "box")
    P.set_bounds(xl, xu); c1 = P.solve(x0)
    P.set_bounds(None, None); c2 = P.solve(x0)
    P.set_bounds(xl, xu); c3 = P.solve(x0)
    assert np.array_equal(c2[0], a[0]) and np.array_equal(c1[0], c3[0]) and not np.array_equal(c1[0], a[0])
    assert np.all(c1[0] >= xl - 2e-5) and np.all(c1[0] <= xu + 2e-5)


def test_bounds_errors(L):
    n, m = 64, 4
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=5, cond=50.0)
    P = L.LargeProblem(L.families.diagquad(Q, A, b, xt, w))
    with pytest.raises(L.LFPSQPError, match="lower bounds cannot be greater"):       # optimize.jl:160-162
        P.set_bounds(np.ones(n), np.zeros(n))
    with pytest.raises(L.LFPSQPError, match="same length"):                          # optimize.jl:144-148
        P.set_bounds(np.zeros(n - 1), np.ones(n - 1))


def test_exact_linesearch_with_bounds(L, oracle):
    n, m = 64, 4
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=5, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    xl, xu = bounds_for(x0, 15, "box")
    P = L.LargeProblem(fam)
    P.set_bounds(xl, xu)
    x, obj, lam, info, st, status = P.solve(x0, L.LFPSQPParams(linesearch=L.exact), return_stats=True)
    op = oracle.default_params(linesearch=1)
    ox, oobj, olam, ot, ost = oracle.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params, params=op)
    # golden-section decisions amplify rounding: same optimum, not the same path
    assert status == 0 and int(info.condition) == ot["condition"]
    assert abs(obj[-1] - oobj[-1]) <= 1e-5 * abs(oobj[-1]) and rel(x, ox) <= 1e-2
    assert np.all(x >= xl - 2e-5) and np.all(x <= xu + 2e-5)


# ------------------------------------------------------------------ host callbacks (generic problems)
def diagquad_callbacks(Q, A, b, xt, w):
    def f(x): return 0.5 * np.sum(w * (x - xt) ** 2)
    def grad(g, x): g[:] = w * (x - xt)
    def c(cv, x): cv[:] = 0.5 * Q @ (x * x) + A @ x - b
    def jac(Jc, cv, x): Jc[:, :] = Q * x[None, :] + A; c(cv, x)
    def hlv(dest, src, x, lam): dest[:] = (w + Q.T @ lam) * src
    return f, grad, c, jac, hlv


@pytest.mark.parametrize("n,m,seed", [(64, 4, 5), (301, 17, 2), (512, 32, 1)])
@pytest.mark.parametrize("nr", [0, 1])
def test_host_callbacks_match_oracle_and_device_family(L, oracle, n, m, seed, nr):
    # optimize(f, grad!, c!, jac!, hess_lag_vec!, x0, xl, xu, m, param) (optimize.jl:119) with numpy callbacks (odd n too)
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    par = L.LFPSQPParams(do_project_retract=not nr)
    x, obj, lam, info, st, status = L.optimize(*diagquad_callbacks(Q, A, b, xt, w), x0, None, None, m, par, return_stats=True)
    ox, oobj, olam, ot, ost = oracle.optimize("diagquad", n, m, 0, x0, fam_params=fam.params,
                                              params=oracle.default_params(do_project_retract=0 if nr else 1))
    assert status == 0 and int(info.condition) == ot["condition"] and abs(info.iter - ot["iter"]) <= 1
    assert rel(x, ox) <= 1e-8 and abs(obj[-1] - oobj[-1]) <= 1e-10 * abs(oobj[-1]) and rel(lam, olam) <= 1e-6
    if n % 2 == 0:   # the device family needs an even column count; same kernels, callbacks on the device instead
        dx, dobj, dlam, dinfo = L.LargeProblem(fam).solve(x0, par)
        assert dinfo.iter == info.iter and rel(x, dx) <= 1e-10


def test_host_callbacks_with_bounds(L, oracle):
    n, m, seed = 64, 4, 5
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=50.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    xl, xu = bounds_for(x0, seed + 10, "mixed")
    x, obj, lam, info = L.optimize(*diagquad_callbacks(Q, A, b, xt, w), x0, xl, xu, m)
    ox, oobj, olam, ot, ost = oracle.optimize("diagquad", n, m, 0, x0, xl=xl, xu=xu, fam_params=fam.params)
    assert int(info.condition) == ot["condition"] and abs(info.iter - ot["iter"]) <= 1
    assert rel(x, ox) <= 1e-8 and abs(obj[-1] - oobj[-1]) <= 1e-10 * abs(oobj[-1])


def test_host_readme_examples(L, oracle):
    # README.md:18-37 (Rosenbrock golden), :41-54 (equality), :57-76 (inequality through the slack wrapper) as GENERIC problems
    def f(x): return (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    def grad(g, x):
        g[0] = -2 * (1 - x[0]) - 400 * x[0] * (x[1] - x[0] ** 2); g[1] = 200 * (x[1] - x[0] ** 2)
    def hlv(dest, src, x, lam):
        h11 = 2 - 400 * (x[1] - x[0] ** 2) + 800 * x[0] ** 2; h12 = -400 * x[0]
        dest[0] = h11 * src[0] + h12 * src[1]; dest[1] = h12 * src[0] + 200 * src[1]
    x, obj, lam, info = L.optimize(f, grad, None, None, hlv, np.zeros(2), None, None, 0)
    assert info.condition == L.TerminationCondition.f_tol and info.iter == 17            # README.md:31-37
    assert info.f_diff == pytest.approx(1.0898882046786806e-7, rel=1e-6)
    assert info.step_diff == pytest.approx(0.0007384068067118611, rel=1e-6)
    assert info.kkt_diff == pytest.approx(4.332627751789361e-5, rel=1e-6)

    n = 50
    def f2(x): return float(x @ x)
    def g2(g, x): g[:] = 2 * x
    def c2(cv, x): cv[0] = x[0] - 0.75
    def j2(Jc, cv, x): Jc[:, :] = 0.0; Jc[0, 0] = 1.0; c2(cv, x)
    def h2(dest, src, x, lam): dest[:] = 2 * src
    x, obj, lam, info = L.optimize(f2, g2, c2, j2, h2, np.ones(n), None, None, 1)
    ox, oobj, olam, ot, _ = oracle.optimize("readme_eq", n, 1, 0, np.ones(n))
    assert int(info.condition) == ot["condition"] and info.iter == ot["iter"] == 1
    assert rel(x, ox) <= 1e-8 and lam[0] == pytest.approx(olam[0], rel=1e-8)

    coeff = np.random.default_rng(0).standard_normal(n)
    def f3(x): return float(coeff @ x)
    def g3(g, x): g[:] = coeff
    def hv3(dest, src, x): dest[:] = 0.0
    def d3(dv, x): dv[0] = x @ x - 1.0
    def dj3(Jd, dv, x): Jd[0, :] = 2.0 * x; d3(dv, x)
    def dh3(dest, src, x, lam): dest[:] = 2.0 * lam[0] * src
    x, obj, lam, info = L.optimize_slack(f3, g3, hv3, None, None, None, d3, dj3, dh3, [-np.inf], [0.0], np.zeros(n), None, None, 0, 1)
    ox, oobj, olam, ot, _ = oracle.optimize("readme_ineq", n, 0, 1, np.zeros(n), xl=-np.inf * np.ones(n),
                                            xu=np.inf * np.ones(n), fam_params=coeff)
    assert int(info.condition) == ot["condition"] and info.iter == ot["iter"]
    assert rel(x, ox) <= 1e-8 and abs(obj[-1] - oobj[-1]) <= 1e-10 * abs(oobj[-1]) and lam.size == 1
    assert rel(x, -coeff / np.linalg.norm(coeff)) < 1e-5                                 # known solution


def test_host_callback_hook_noise_and_errors(L):
    n, m = 64, 4
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=5, cond=50.0)
    cbs = diagquad_callbacks(Q, A, b, xt, w)
    seen = []
    par = L.LFPSQPParams(callback=lambda i, x: seen.append((i, x.copy())), callback_period=2)   # optimize.jl:432-434
    x, obj, lam, info = L.optimize(*cbs, x0, None, None, m, par)
    assert [i for i, _ in seen] == list(range(2, info.iter + 1, 2)) and seen[-1][1].shape == (n,)
    # beta > 0 (optimize.jl:264-273): the noise comes from the caller's RNG through the randn callback => reproducible
    rng1 = np.random.default_rng(7); rng2 = np.random.default_rng(7)
    pb = L.LFPSQPParams(beta=1e-3, t_beta=5, maxiter=50)
    a = L.optimize_explicit(*cbs, x0, None, None, m, pb, randn=lambda k: rng1.standard_normal(k))
    b2 = L.optimize_explicit(*cbs, x0, None, None, m, pb, randn=lambda k: rng2.standard_normal(k))
    assert np.array_equal(a[0], b2[0]) and not np.array_equal(a[0], x)
    with pytest.raises(L.LFPSQPError, match="beta>0"):
        L.optimize(*cbs, x0, None, None, m, pb)
    # an exception inside a callback aborts the solve and is re-raised to the caller
    def bad_grad(g, x): raise ValueError("boom")
    with pytest.raises(ValueError, match="boom"):
        L.optimize(cbs[0], bad_grad, *cbs[2:], x0, None, None, m)
