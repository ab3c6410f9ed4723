"""Pins the CPU oracle against every known answer the reference holds for the hot path."""
import numpy as np
import pytest


def test_readme_rosenbrock_golden(oracle):
    # /root/reference/README.md:31-37 -- the only end-to-end known answer in the reference
    x, obj, lam, term, st = oracle.optimize("rosenbrock", 2, 0, 0, [0.0, 0.0])
    assert term["condition"] == oracle.F_TOL
    assert term["iter"] == 17
    assert term["f_diff"] == pytest.approx(1.0898882046786806e-7, rel=1e-9)
    assert term["step_diff"] == pytest.approx(0.0007384068067118611, rel=1e-9)
    assert term["kkt_diff"] == pytest.approx(4.332627751789361e-5, rel=1e-9)
    assert len(obj) == 18 and lam.size == 0
    assert np.allclose(x, [1.0, 1.0], atol=1e-6)


def test_readme_equality_example(oracle):
    # README.md:41-54 (config C1). Values hand-checked in SURVEY.md App. D.
    x, obj, lam, term, st = oracle.optimize("readme_eq", 50, 1, 0, np.ones(50))
    assert term["condition"] == oracle.KKT_TOL and term["iter"] == 1
    assert term["kkt_diff"] == 0.0
    assert term["f_diff"] == pytest.approx(49.437499625000314, rel=1e-13)
    assert term["step_diff"] == pytest.approx(7.0044628541380805, rel=1e-13)
    assert x[0] == pytest.approx(0.75000024999975, rel=1e-12) and np.all(x[1:] == 0.0)
    assert obj[-1] == pytest.approx(0.5625003749996875, rel=1e-12)
    assert lam[0] == pytest.approx(-1.5000005, rel=1e-9)
    assert st["retract_outer"] == 5 and st["retract_pcg"] == 5


def test_readme_inequality_example(oracle):
    # README.md:57-76 (config C2, one instance). x* = -coeff/|coeff|, f* = -|coeff|, lambda = |coeff|/2
    rng = np.random.default_rng(0)
    co = rng.standard_normal(50)
    inf = np.inf * np.ones(50)
    x, obj, lam, term, st = oracle.optimize("readme_ineq", 50, 0, 1, np.zeros(50), xl=-inf, xu=inf, fam_params=co)
    nc = np.linalg.norm(co)
    assert term["condition"] == oracle.F_TOL
    assert np.linalg.norm(x + co / nc) < 5e-6
    assert obj[-1] == pytest.approx(-nc, abs=5e-6)
    assert lam.shape == (1,) and lam[0] == pytest.approx(nc / 2, rel=1e-5)   # untruncated lambda (optimize.jl:67-70)
    assert np.dot(x, x) - 1.0 < 1e-6                                         # feasible to eps_c


def test_armijo_known_answer(oracle):
    # test/test_linesearch.jl:14-22 : f=x^2, x=-0.23, d=1  => alpha = step_diff = 0.25
    # boxquad with t=0 is f=|x|^2
    prm = np.zeros(3)
    flag, xnew, newf, fd, sd, al = oracle.linesearch_euclid("boxquad", 1, [-0.23], [1.0], "armijo", fam_params=prm)
    assert flag == 0 and al == 0.25 and sd == pytest.approx(0.25)
    assert newf == pytest.approx(xnew[0] ** 2) and fd == pytest.approx(0.23 ** 2 - newf)


def test_exact_linesearch_known_answer(oracle):
    # test/test_linesearch.jl:24-32 : alpha ~= 0.23 (atol 1e-6)
    prm = np.zeros(3)
    flag, xnew, newf, fd, sd, al = oracle.linesearch_euclid("boxquad", 1, [-0.23], [1.0], "exact", fam_params=prm)
    assert flag == 0 and abs(al - 0.23) < 1e-6 and sd == pytest.approx(al)
    assert newf == pytest.approx(xnew[0] ** 2) and fd == pytest.approx(0.23 ** 2 - newf)
