"""Registered device problem families.

Julia closures cannot run on the GPU, so the `f`, `c!`, `d!` arguments of `optimize` are handles onto device code
compiled into liblfpsqp_b200.so (family id + parameter blob), implementing the callback contract of
src/autodiff_generators.jl (grad! :7-9, jac! :40-42, hess_lag_vec! :80-104) analytically.

    fam = families.readme_inequality(coeff)
    x, obj_values, lam, term_info = optimize(fam.f, None, fam.d, x0, xl, xu, 0, 1)
"""
import numpy as np

ROSENBROCK, README_EQ, README_INEQ, THOMSON, DIAGQUAD, SIN, BOXQUAD = range(7)


class DeviceCallback:
    """A handle onto one role ('f', 'c', 'd') of a registered device family."""

    def __init__(self, family, role):
        self.family = family
        self.role = role

    def __repr__(self):
        return "<device %s of %r>" % (self.role, self.family)


class Family:
    def __init__(self, fam_id, name, n, m, p, params=None, batched_params=False):
        self.id = fam_id
        self.name = name
        self.n, self.m, self.p = n, m, p
        # params: (P,) shared blob or (B, P) per-instance blobs
        self.params = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        self.batched_params = batched_params
        self.f = DeviceCallback(self, "f")
        self.c = DeviceCallback(self, "c") if m > 0 else None
        self.d = DeviceCallback(self, "d") if p > 0 else None

    def __repr__(self):
        return "Family(%s, n=%d, m=%d, p=%d)" % (self.name, self.n, self.m, self.p)


def rosenbrock():
    """README.md:18-22: f(x) = (1-x1)^2 + 100 (x2 - x1^2)^2."""
    return Family(ROSENBROCK, "rosenbrock", 2, 0, 0)


def readme_equality(n=50):
    """README.md:41-54: f = dot(x,x), c = x[1] - 0.75."""
    return Family(README_EQ, "readme_eq", n, 1, 0)


def readme_inequality(coeff):
    """README.md:57-76: f = dot(coeff, x), d = dot(x,x) - 1 <= 0.  coeff: (n,) or (B, n) for a batch."""
    coeff = np.asarray(coeff, dtype=np.float64)
    return Family(README_INEQ, "readme_ineq", coeff.shape[-1], 0, 1, coeff, batched_params=coeff.ndim == 2)


def thomson(npoints):
    """f = sum_{i<j} 1/|x_i - x_j|, c_i = |x_i|^2 - 1; n = 3*npoints, m = npoints."""
    return Family(THOMSON, "thomson", 3 * npoints, npoints, 0)


def diagquad(Q, A, b, xt, w):
    """c_i = 1/2 sum_j Q_ij x_j^2 + A_i.x - b_i ; f = 1/2 sum_j w_j (x_j - xt_j)^2.  Q, A: (m, n)."""
    Q = np.asarray(Q, dtype=np.float64); A = np.asarray(A, dtype=np.float64)
    m, n = Q.shape
    blob = np.concatenate([Q.ravel(), A.ravel(), np.asarray(b, float).ravel(), np.asarray(xt, float).ravel(),
                           np.asarray(w, float).ravel()])
    return Family(DIAGQUAD, "diagquad", n, m, 0, blob)


def sin_system(n, m, t=None):
    """test/test_retractions.jl:34-54: c_i = x[2i] - sin(x[2i-1]); objective 1/2 |x - t|^2."""
    t = np.zeros(n) if t is None else np.asarray(t, dtype=np.float64)
    return Family(SIN, "sin", n, m, 0, t, batched_params=t.ndim == 2)


def boxquad(t, a=None, b=0.0):
    """f = |x - t|^2 with an optional linear equality a.x = b (used with bounds xl <= x <= xu)."""
    t = np.asarray(t, dtype=np.float64)
    n = t.shape[-1]
    m = 0 if a is None else 1
    av = np.zeros(n) if a is None else np.asarray(a, dtype=np.float64)
    if t.ndim == 2:
        B = t.shape[0]
        blob = np.concatenate([t, np.broadcast_to(av, (B, n)), np.full((B, 1), float(b))], axis=1)
        return Family(BOXQUAD, "boxquad", n, m, 0, blob, batched_params=True)
    return Family(BOXQUAD, "boxquad", n, m, 0, np.concatenate([t, av, [float(b)]]))
