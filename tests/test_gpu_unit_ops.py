"""The reference's component tests (test/test_retractions.jl, test/test_inequalities.jl) re-run against the DEVICE code
through the unit-level C-ABI exports, next to the same calls on the CPU oracle."""
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


def _ineq_setup(rng, n=16):
    q4 = n // 4
    xl = np.concatenate([-np.inf * np.ones(q4), rng.standard_normal(q4), -np.inf * np.ones(q4), rng.standard_normal(q4)])
    xu = np.concatenate([np.inf * np.ones(2 * q4), rng.standard_normal(q4), xl[3 * q4:] + 0.5 + rng.random(q4)])
    x = np.concatenate([rng.standard_normal(q4), xl[q4:2 * q4] + rng.integers(0, 3, q4), xu[2 * q4:3 * q4] - rng.integers(0, 3, q4),
                        xl[3 * q4:] + rng.random(q4) * (xu[3 * q4:] - xl[3 * q4:])])
    return xl, xu, x


def test_bound_embedding_ops_match_reference_tests(L, oracle):
    # test/test_inequalities.jl:39-52, :80-141, :143-155, :180-200 on the device
    rng = np.random.default_rng(7)
    n, m = 16, 5
    xl, xu, x = _ineq_setup(rng, n)
    q, r, s, t, il, ip = oracle.ineq_data(xl, xu)
    xaug = L.ineq_op("initial_y", xl, xu, x)
    assert np.array_equal(xaug, oracle.ineq_initial_y(xl, xu, x))
    assert np.allclose(L.ineq_op("h", xl, xu, xaug), 0.0, atol=1e-14)                       # :45-51
    xr = xaug + 0.1 * rng.standard_normal(2 * n)
    assert np.allclose(L.ineq_op("h", xl, xu, xr), oracle.ineq_h(xl, xu, xr), rtol=1e-14, atol=1e-15)
    g = L.ineq_op("gradient", xl, xu, xaug)
    Dx, Dy, S = oracle.ineq_gradient(xl, xu, xaug)
    assert np.allclose(g[:n], Dx, atol=1e-15) and np.allclose(g[n:2 * n], Dy, atol=1e-15) and np.allclose(g[2 * n:], S, atol=1e-15)
    Jct = rng.standard_normal((n, m))
    ghx = Dx * S; ghy = Dy * S
    bigA = np.block([[np.diag(ghx), Jct], [np.diag(ghy), np.zeros((n, m))]])
    v = rng.standard_normal(n + m); w = rng.standard_normal(2 * n)
    assert np.allclose(L.ineq_op("bigA", xl, xu, np.r_[xaug, v], J=Jct.T), bigA @ v, atol=1e-13)      # :117-118
    assert np.allclose(L.ineq_op("bigAt", xl, xu, np.r_[xaug, w], J=Jct.T), bigA.T @ w, atol=1e-13)  # :120-121
    # projector Q Q' and multipliers vs the SVD route of the reference (bigQ from svd(PJct), lambda = bigA \\ d)
    PJ = np.vstack([(1 - Dx * Dx)[:, None] * Jct, (-Dy * Dx)[:, None] * Jct])
    U, Sig, Vt = np.linalg.svd(PJ, full_matrices=False)
    bigQ = np.hstack([np.vstack([np.diag(Dx), np.diag(Dy)]), U])
    d = rng.standard_normal(2 * n)
    out = L.ineq_op("project", xl, xu, np.r_[xaug, d], J=Jct.T)
    assert np.allclose(out[:2 * n], d - bigQ @ (bigQ.T @ d), atol=1e-13)
    lam_ref = np.linalg.lstsq(bigA, d, rcond=None)[0]                                        # :154
    assert np.allclose(np.r_[out[3 * n:3 * n + 0], out[2 * n + m:], out[2 * n:2 * n + m]], lam_ref, atol=1e-11)
    olam, olamy = oracle.ineq_lambda(Dx, Dy, S, Jct, d)
    assert np.allclose(out[2 * n:2 * n + m], olam, atol=1e-11) and np.allclose(out[2 * n + m:], olamy, atol=1e-11)
    # y retraction after a tangent step (:180-200)
    dt = d - bigQ @ (bigQ.T @ d)
    xnew = L.ineq_op("y_retract", xl, xu, np.r_[xaug, xaug + dt])
    assert np.allclose(L.ineq_op("h", xl, xu, xnew), 0.0, atol=1e-12)
    assert np.allclose(xnew, oracle.y_retract(xl, xu, xaug, xaug + dt), rtol=1e-12, atol=1e-13)


def _diagquad(L, n, m, seed):
    Q, A, b, xt, w, x0 = L.make_diagquad(n, m, seed=seed, cond=100.0)
    fam = L.families.diagquad(Q, A, b, xt, w)
    return Q, A, b, x0, fam, L.LargeProblem(fam)


def test_pcg_matches_reference_test(L, oracle):
    # test/test_retractions.jl:105-126: flag 0, |r| < tol, |mu x + J'(J x) - b| < tol ; same iteration count as the oracle
    n, m = 1000, 100
    Q, A, b, x0, fam, P = _diagquad(L, n, m, 3)
    J = Q * x0[None, :] + A
    rng = np.random.default_rng(3)
    for mu in (1e-1, 1e-2, 1e-4):
        rhs = rng.standard_normal(n)
        x, r, flag, it = P.pcg(x0, mu, rhs, tol=1e-6, maxiter=200)
        ox, orr, oflag, oit = oracle.pcg_dense(J, mu, rhs, tol=1e-6, maxiter=200)
        assert flag == 0 and np.linalg.norm(r) < 1e-6
        assert np.linalg.norm(mu * x + J.T @ (J @ x) - rhs) < 1e-6
        assert abs(it - oit) <= 1 and rel(x, ox) < 1e-6


@pytest.mark.parametrize("method", ["nr", "pp"])
def test_retractions_match_reference_test(L, oracle, method):
    # test/test_retractions.jl:90-102 / :144-157: flag 0, |c|_inf < tol, cval == c(xnew), step orthogonality / length
    n, m = 1000, 100
    Q, A, b, x0, fam, P = _diagquad(L, n, m, 4)
    J = Q * x0[None, :] + A
    rng = np.random.default_rng(4)
    step = rng.standard_normal(n); step -= J.T @ np.linalg.solve(J @ J.T, J @ step); step *= 2.0 / np.linalg.norm(step)
    xt = x0 + step
    for tol in (1e-6, 1e-8):
        prm = L.LFPSQPParams(ϵ_c=tol, maxiter_retract=200, maxiter_pcg=200)
        flag, xnew, cval, it, pit = P.retract(method, x0, xt, prm)
        oflag, oxnew, ocval, oit, opit = oracle.retract("diagquad", n, m, method, x0, xt, tol, maxiter=200, maxiter_pcg=200,
                                                        fam_params=fam.params)
        c2 = 0.5 * Q @ (xnew * xnew) + A @ xnew - b
        assert flag == 0 == oflag and np.max(np.abs(cval)) < tol
        assert np.allclose(cval, c2, rtol=0, atol=1e-12)          # cval is c! at xnew
        assert abs(it - oit) <= 1 and rel(xnew, oxnew) < 1e-8
        if method == "nr":
            assert abs(step @ (xnew - xt)) < 1e-6
        else:
            assert np.linalg.norm(step) >= np.linalg.norm(xnew - x0) - tol


def test_linesearch_reference_goldens_on_device(L, oracle):
    # test/test_linesearch.jl:14-22 (armijo!: f = x^2, x = -0.23, d = 1 -> alpha = step_diff = 0.25 exactly) and :24-32
    # (exact_linesearch!: alpha ~= 0.23, atol 1e-6) run on the DEVICE through lfpsqp_linesearch; boxquad with t = 0 is f = |x|^2
    fam = L.families.boxquad(np.zeros(1))
    x = np.array([-0.23]); d = np.array([1.0])
    flag, it1, it2, newf, f_diff, step_diff, alpha, xnew = L.linesearch("armijo", fam, x, d)
    assert flag == it1 == it2 == 0
    assert x[0] == -0.23                                   # input not changed
    assert newf == pytest.approx(xnew[0] ** 2, rel=1e-15) and f_diff == pytest.approx(0.23 ** 2 - newf, rel=1e-14)
    assert alpha == 0.25 and step_diff == pytest.approx(0.25, rel=1e-15)
    oflag, oxnew, onewf, ofd, osd, oal = oracle.linesearch_euclid("boxquad", 1, x, d, "armijo", fam_params=np.zeros(3))
    assert (flag, alpha) == (oflag, oal) and xnew[0] == oxnew[0] and newf == onewf
    flag, it1, it2, newf, f_diff, step_diff, alpha, xnew = L.linesearch("exact", fam, x, d)
    assert flag == it1 == it2 == 0
    assert newf == pytest.approx(xnew[0] ** 2, rel=1e-15) and f_diff == pytest.approx(0.23 ** 2 - newf, rel=1e-12)
    assert step_diff == pytest.approx(alpha, rel=1e-12) and abs(alpha - 0.23) < 1e-6
    oflag, oxnew, onewf, ofd, osd, oal = oracle.linesearch_euclid("boxquad", 1, x, d, "exact", fam_params=np.zeros(3))
    assert flag == oflag and alpha == pytest.approx(oal, rel=1e-12) and xnew[0] == pytest.approx(oxnew[0], rel=1e-12)
    # with retraction: the sin system (test/test_retractions.jl:34-54), NR and ProjPenalty, tangent direction
    rng = np.random.default_rng(12)
    n, m = 24, 6
    t = rng.standard_normal(n)
    fam = L.families.sin_system(n, m, t)
    x0 = np.zeros(n)                                       # feasible: x[2i] = sin(x[2i-1]) = 0
    J = np.zeros((m, n))
    for i in range(m):
        J[i, 2 * i + 1] = 1.0; J[i, 2 * i] = -np.cos(x0[2 * i])
    g = x0 - t
    dd = -g - J.T @ np.linalg.solve(J @ J.T, J @ (-g))
    for nr in (False, True):
        prm = L.LFPSQPParams(do_project_retract=not nr)
        flag, it1, it2, newf, f_diff, step_diff, alpha, xnew = L.linesearch("armijo", fam, x0, dd, param=prm)
        cv = np.array([xnew[2 * i + 1] - np.sin(xnew[2 * i]) for i in range(m)])
        assert flag == 0 and np.max(np.abs(cv)) < 1e-6 and it1 >= 1
        assert newf == pytest.approx(0.5 * np.sum((xnew - t) ** 2), rel=1e-13)
        assert newf - 0.5 * np.sum((x0 - t) ** 2) <= 1e-4 * alpha * (dd @ g)                 # Armijo-Goldstein (linesearch.jl:75)
        assert step_diff == pytest.approx(np.linalg.norm(xnew - x0), rel=1e-12)
        assert (it2 == 0) if nr else (it2 >= 1)


def test_augmented_hessian_action_on_device(L):
    # test/test_inequalities.jl:157-177: dest = bigH v, bigH = [H + 2 diag(lam_y.q), 0; 0, 2 diag(lam_y.s)], H = Lagrangian Hessian
    rng = np.random.default_rng(9)
    n, m = 16, 5
    xl, xu, x = _ineq_setup(rng, n)
    q, r, s, t, il, ip = __import__("oracle.oracle", fromlist=["x"]).ineq_data(xl, xu)
    xaug = L.ineq_op("initial_y", xl, xu, x)
    lam = rng.standard_normal(m); lamy = rng.standard_normal(n); v = rng.standard_normal(2 * n)
    fam = L.families.sin_system(n, m, rng.standard_normal(n))
    dest = L.aug_hess_vec(fam, xl, xu, xaug, lam, lamy, v)
    H = np.eye(n)
    for i in range(m):
        H[2 * i, 2 * i] += lam[i] * np.sin(xaug[2 * i])
    bigH = np.block([[H + 2 * np.diag(lamy * q), np.zeros((n, n))], [np.zeros((n, n)), 2 * np.diag(lamy * s)]])
    assert np.allclose(dest, bigH @ v, rtol=0, atol=1e-14)
    fam = L.families.boxquad(rng.standard_normal(n), a=np.ones(n), b=1.0)
    dest = L.aug_hess_vec(fam, xl, xu, xaug, lam[:1], lamy, v)
    bigH = np.block([[2 * np.eye(n) + 2 * np.diag(lamy * q), np.zeros((n, n))], [np.zeros((n, n)), 2 * np.diag(lamy * s)]])
    assert np.allclose(dest, bigH @ v, rtol=0, atol=1e-14)


def test_projcg_general_c_nonzero(L, oracle):
    # test/test_cg.jl:5-28: projcg! with c != 0 (x0 = U c, projcg.jl:55): nr < tol, |U'x - c| small, KKT residual small
    n, m = 1000, 10
    Q, A, b, x0, fam, P = _diagquad(L, n, m, 8)
    J = Q * x0[None, :] + A
    P.factor(x0, want=())
    rng = np.random.default_rng(8)
    lam = 0.1 * rng.standard_normal(m)
    w = fam.params[2 * m * n + m + n:]
    hd = w + Q.T @ lam
    assert np.all(hd > 0)
    Lc = np.linalg.cholesky(J @ J.T)
    U = J.T @ np.linalg.inv(Lc).T                              # orthonormal basis of range(J') (Cholesky-QR)
    bv = rng.standard_normal(n); cv = rng.standard_normal(m)
    bigM = np.block([[np.diag(hd), U], [U.T, np.zeros((m, m))]])
    rhs = np.r_[bv, cv]
    for tol in (1e-6, 1e-9, 1e-12):
        r = P.projcg_general(x0, lam, bv, cv, tol=tol)
        assert r["status"] in (1, 3) and r["nr"] < tol
        assert np.linalg.norm(U.T @ r["sol"] - cv) < 1e-12
        assert np.linalg.norm(bigM @ np.r_[r["sol"], r["lam"]] - rhs) < max(10 * tol, 1e-11)
        ox, olam, oit, onr = oracle.projcg_dense(np.diag(hd), U, bv, cv, tol=tol)
        assert abs(r["iters"] - oit) <= 1 and rel(r["sol"], ox) < 1e-8
    # c = 0 through the same entry point equals the dedicated path's structure: U'x = 0
    r = P.projcg_general(x0, lam, bv, np.zeros(m), tol=1e-10)
    assert np.linalg.norm(U.T @ r["sol"]) < 1e-12
