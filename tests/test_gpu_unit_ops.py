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
