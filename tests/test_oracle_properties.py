"""The reference's component property tests (test/*.jl) re-expressed against the CPU oracle (seeded)."""
import numpy as np
import pytest


def test_projcg_accuracy(oracle):
    # test/test_cg.jl:1-29
    rng = np.random.default_rng(1)
    n, m = 1000, 10
    A = 0.01 * rng.standard_normal((n, n)); A = A @ A.T + 0.5 * np.eye(n)
    b = rng.standard_normal(n); c = rng.standard_normal(m)
    U, _ = np.linalg.qr(rng.standard_normal((n, m)))
    big = np.block([[A, U], [U.T, np.zeros((m, m))]]); rhs = np.concatenate([b, c])
    for e in range(6, 21, 2):
        tol = 10.0 ** (-e)
        x, lam, it, nr = oracle.projcg_dense(A, U, b, c, tol=tol)
        if nr < tol:  # reachable tolerances (below ~1e-15 the reference's own test relies on rg<=0 breaks)
            assert np.linalg.norm(U.T @ x - c) < 1e-13
            assert np.linalg.norm(big @ np.concatenate([x, lam]) - rhs) < max(tol, 1e-12)
    x, lam, it, nr = oracle.projcg_dense(A, U, b, c, tol=1e-10)
    assert nr < 1e-10


def test_projcg_negative_curvature(oracle):
    # test/test_cg.jl:39-54
    rng = np.random.default_rng(2)
    n, m = 300, 10
    S, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam_ = np.concatenate([rng.random(n - 2 * m) + 1, -1 - rng.random(2 * m)])
    A = S @ np.diag(lam_) @ S.T
    U, _ = np.linalg.qr(rng.standard_normal((n, m)))
    b = rng.standard_normal(n)
    x, lam, it, nr = oracle.projcg_dense(A, U, b, np.zeros(m), tol=1e-20)
    assert np.isinf(nr) and np.all(np.isnan(lam))
    assert np.linalg.norm(U.T @ x) < 1e-13 and x @ A @ x <= 0.0


def test_pcg(oracle):
    # test/test_retractions.jl:105-141 (no preconditioner part)
    rng = np.random.default_rng(3)
    m, n = 100, 1000
    J = rng.standard_normal((m, n))
    for mu in (1e-1, 1e-2, 1e-4):
        b = rng.standard_normal(n)
        x, r, flag, it = oracle.pcg_dense(J, mu, b, tol=1e-6, maxiter=200)
        assert flag == 0 and np.linalg.norm(r) < 1e-6
        assert np.linalg.norm(mu * x + J.T @ (J @ x) - b) < 1e-6


def _sin_setup(rng, n=1000, m=100):
    x0 = np.zeros(n)
    J = np.zeros((m, n))
    for i in range(m):
        J[i, 2 * i + 1] = 1.0; J[i, 2 * i] = -np.cos(x0[2 * i])
    U, _, _ = np.linalg.svd(J.T, full_matrices=False)
    step = rng.standard_normal(n); step -= U @ (U.T @ step); step *= 5.0 / np.linalg.norm(step)
    return x0, step


@pytest.mark.parametrize("tol", [1e-6, 1e-8])
def test_nr_retraction(oracle, tol):
    # test/test_retractions.jl:90-102
    rng = np.random.default_rng(4)
    n, m = 1000, 100
    x0, step = _sin_setup(rng, n, m)
    xt = x0 + step; xt_copy = xt.copy()
    flag, xnew, cval, it, _ = oracle.retract("sin", n, m, "nr", x0, xt, tol, maxiter=1000, fam_params=np.zeros(n))
    c2 = np.array([xnew[2 * i + 1] - np.sin(xnew[2 * i]) for i in range(m)])
    assert flag == 0 and np.max(np.abs(cval)) < tol
    assert np.all(cval == c2) and np.all(xt == xt_copy)
    assert abs(step @ (xnew - xt)) < 1e-6


@pytest.mark.parametrize("tol", [1e-6, 1e-8, 1e-10])
def test_pp_retraction(oracle, tol):
    # test/test_retractions.jl:144-157
    rng = np.random.default_rng(5)
    n, m = 1000, 100
    x0, step = _sin_setup(rng, n, m)
    xt = x0 + step; xt_copy = xt.copy()
    flag, xnew, cval, it, pit = oracle.retract("sin", n, m, "pp", x0, xt, tol, maxiter=100, maxiter_pcg=200,
                                               fam_params=np.zeros(n))
    c2 = np.array([xnew[2 * i + 1] - np.sin(xnew[2 * i]) for i in range(m)])
    assert flag == 0 and np.max(np.abs(cval)) < tol
    assert np.all(cval == c2) and np.all(xt == xt_copy)
    assert np.linalg.norm(step) >= np.linalg.norm(xnew - x0) - tol


def _ineq_setup(rng, n=16, m=5):
    q4 = n // 4
    xl = np.concatenate([-np.inf * np.ones(q4), rng.standard_normal(q4), -np.inf * np.ones(q4), rng.standard_normal(q4)])
    xu = np.concatenate([np.inf * np.ones(2 * q4), rng.standard_normal(q4), xl[3 * q4:] + 0.5 + rng.random(q4)])
    x = np.concatenate([rng.standard_normal(q4), xl[q4:2 * q4] + rng.integers(0, 3, q4),
                        xu[2 * q4:3 * q4] - rng.integers(0, 3, q4),
                        xl[3 * q4:] + rng.random(q4) * (xu[3 * q4:] - xl[3 * q4:])])
    return xl, xu, x


def test_inequality_data_and_initial_y(oracle):
    # test/test_inequalities.jl:22-52
    rng = np.random.default_rng(6)
    n = 16; q4 = 4
    xl, xu, x = _ineq_setup(rng)
    q, r, s, t, il, ip = oracle.ineq_data(xl, xu)
    assert np.allclose(q, np.r_[np.zeros(3 * q4), np.ones(q4)])
    assert np.allclose(r, np.r_[np.zeros(q4), xl[q4:2 * q4], xu[2 * q4:3 * q4], xl[3 * q4:] / 2 + xu[3 * q4:] / 2])
    assert np.allclose(s, np.r_[np.zeros(q4), -np.ones(q4), np.ones(2 * q4)])
    assert np.allclose(t, np.r_[np.zeros(q4), xl[q4:2 * q4], xu[2 * q4:3 * q4], (xu[3 * q4:] - xl[3 * q4:]) ** 2 / 4])
    assert np.all(il == np.r_[np.ones(q4, bool), np.zeros(3 * q4, bool)])
    assert np.all(ip == np.r_[np.zeros(q4, bool), np.ones(2 * q4, bool), np.zeros(q4, bool)])
    xaug = oracle.ineq_initial_y(xl, xu, x)
    assert np.allclose(oracle.ineq_h(xl, xu, xaug), 0.0, atol=1e-14)


def test_inequality_operators_lambda_yretract(oracle):
    # test/test_inequalities.jl:54-200
    rng = np.random.default_rng(7)
    n, m = 16, 5
    xl, xu, x = _ineq_setup(rng)
    q, r, s, t, il, ip = oracle.ineq_data(xl, xu)
    xaug = oracle.ineq_initial_y(xl, xu, x)
    Jct = rng.standard_normal((n, m))
    ghx = 2 * q * (x - r) + (1 - q ** 2); ghy = 2 * s * (xaug[n:] - r) - (1 - s ** 2)
    S = np.sqrt(ghx ** 2 + ghy ** 2); Dx = ghx / S; Dy = ghy / S
    oDx, oDy, oS = oracle.ineq_gradient(xl, xu, xaug)
    assert np.allclose(oDx, Dx, atol=1e-15) and np.allclose(oDy, Dy, atol=1e-15) and np.allclose(oS, S, atol=1e-15)
    PJ = np.vstack([(1 - Dx * Dx)[:, None] * Jct, (-Dy * Dx)[:, None] * Jct])
    U, Sig, Vt = np.linalg.svd(PJ, full_matrices=False)
    bigA = np.block([[np.diag(ghx), Jct], [np.diag(ghy), np.zeros((n, m))]])
    bigQ = np.hstack([np.vstack([np.diag(Dx), np.diag(Dy)]), U])
    assert np.allclose(bigQ.T @ bigQ, np.eye(n + m), atol=1e-13)
    v = rng.standard_normal(n + m); w = rng.standard_normal(2 * n)
    assert np.allclose(oracle.ineq_mul("Q", Dx, Dy, S, Jct, U, m, v), bigQ @ v, atol=1e-14)
    assert np.allclose(oracle.ineq_mul("Qt", Dx, Dy, S, Jct, U, m, w), bigQ.T @ w, atol=1e-14)
    assert np.allclose(oracle.ineq_mul("A", Dx, Dy, S, Jct, U, m, v), bigA @ v, atol=1e-14)
    assert np.allclose(oracle.ineq_mul("At", Dx, Dy, S, Jct, U, m, w), bigA.T @ w, atol=1e-14)
    v2 = rng.standard_normal(n + m - 2)   # reduced rank, :122-138
    assert np.allclose(oracle.ineq_mul("Q", Dx, Dy, S, Jct, U, m - 2, v2), bigQ[:, :n + m - 2] @ v2, atol=1e-14)
    assert np.allclose(oracle.ineq_mul("Qt", Dx, Dy, S, Jct, U, m - 2, w), bigQ[:, :n + m - 2].T @ w, atol=1e-14)
    d = rng.standard_normal(2 * n)          # multipliers, :143-155
    lam, lamy = oracle.ineq_lambda(Dx, Dy, S, Jct, d)
    ref = np.linalg.lstsq(bigA, d, rcond=None)[0]
    assert np.allclose(np.r_[lamy, lam], ref, atol=1e-12)
    # y retraction after a tangent step, :180-200
    d = rng.standard_normal(2 * n); d -= bigQ @ (bigQ.T @ d)
    xnew = oracle.y_retract(xl, xu, xaug, xaug + d)
    assert np.allclose(oracle.ineq_h(xl, xu, xnew), 0.0, atol=1e-12)


def test_bound_embedding_end_to_end(oracle):
    # f=|x-t|^2 with 3 free / 3 lower / 3 upper / 3 boxed variables converges to clip(t)
    rng = np.random.default_rng(8)
    n = 12
    xl = np.r_[-np.inf * np.ones(3), -0.5 * np.ones(3), -np.inf * np.ones(3), -0.3 * np.ones(3)]
    xu = np.r_[np.inf * np.ones(6), 0.4 * np.ones(3), 0.6 * np.ones(3)]
    t = 2 * rng.standard_normal(n)
    prm = np.r_[t, np.ones(n), 3.0]
    x0 = np.clip(np.zeros(n), xl, xu)
    x, obj, lam, term, st = oracle.optimize("boxquad", n, 0, 0, x0, xl=xl, xu=xu, fam_params=prm)
    assert term["condition"] == oracle.F_TOL
    assert np.max(np.abs(x - np.clip(t, xl, xu))) < 1e-4
    # plus one linear equality sum(x)=3: KKT multiplier sign convention grad f + J' lam = 0 on free variables
    x, obj, lam, term, st = oracle.optimize("boxquad", n, 1, 0, x0 + 0.25, xl=xl, xu=xu, fam_params=prm)
    assert term["condition"] in (oracle.F_TOL, oracle.KKT_TOL)
    assert abs(x.sum() - 3.0) < 1e-5
    free = slice(0, 3)
    assert np.allclose(2 * (x[free] - t[free]) + lam[0], 0.0, atol=1e-3)
    assert np.all(x >= xl - 1e-6) and np.all(x <= xu + 1e-6)


def test_batched_matches_single(oracle):
    rng = np.random.default_rng(9)
    B = 16
    x0 = rng.uniform(-2, 2, (B, 2)); x0[0] = 0.0
    x, obj, olen, lam, term, st = oracle.optimize_batched("rosenbrock", 2, 0, 0, x0, nthreads=4)
    for k in (0, 5, 15):
        xs, objs, _, ts, _ = oracle.optimize("rosenbrock", 2, 0, 0, x0[k])
        assert np.all(xs == x[k]) and ts["iter"] == term["iter"][k] and olen[k] == len(objs)
    assert term["iter"][0] == 17
