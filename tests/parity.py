"""Shared parity checker: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

north_star tolerances (BASELINE.json): same termination condition, iteration count within +-1, final x within 1e-8
relative, objective within 1e-10 relative, all Float64."""
import numpy as np

X_RTOL = 1e-8
F_RTOL = 1e-10


def compare_batch(gpu, orc, n_user, label=""):
    """gpu/orc: tuples (x, obj, olen, lam, term).  Returns a dict of per-criterion pass fractions + worst errors."""
    gx, gobj, glen, glam, gterm = gpu[:5]
    ox, oobj, olen, olam, oterm = orc[:5]
    B = gx.shape[0]
    cond_ok = gterm["condition"] == oterm["condition"]
    it_diff = np.abs(gterm["iter"].astype(np.int64) - oterm["iter"].astype(np.int64))
    xs = np.maximum(np.linalg.norm(ox, axis=1), 1e-300)
    x_err = np.linalg.norm(gx - ox, axis=1) / xs
    H = min(gobj.shape[1], oobj.shape[1])
    gf = np.array([gobj[k, min(glen[k], H) - 1] for k in range(B)])
    of = np.array([oobj[k, min(olen[k], H) - 1] for k in range(B)])
    # objective: 1e-10 relative, with a floor of a few ulps of the objective's scale along the trajectory (an optimum
    # with f* ~ 0, e.g. Rosenbrock, has no meaningful relative error: the reference's own rounding exceeds it)
    fscale = np.array([np.nanmax(np.abs(oobj[k, :min(olen[k], H)])) for k in range(B)])
    f_err = np.abs(gf - of) / (np.abs(of) + 8 * 2.2e-16 * fscale / F_RTOL + 1e-300)
    lam_err = np.zeros(B)
    if glam.size:
        lam_err = np.linalg.norm(glam - olam, axis=1) / np.maximum(np.linalg.norm(olam, axis=1), 1e-300)
    same_path = cond_ok & (it_diff == 0)
    res = dict(label=label, B=B, cond_frac=cond_ok.mean(), iter_exact_frac=(it_diff == 0).mean(),
               iter_pm1_frac=(it_diff <= 1).mean(), x_frac=(x_err <= X_RTOL).mean(), f_frac=(f_err <= F_RTOL).mean(),
               x_err_max_same_path=float(x_err[same_path].max()) if same_path.any() else float("nan"),
               f_err_max_same_path=float(f_err[same_path].max()) if same_path.any() else float("nan"),
               x_err_max=float(x_err.max()), f_err_max=float(f_err.max()), lam_err_max=float(lam_err.max()),
               status_nonzero=int((gterm["status"] != 0).sum()),
               all_ok_frac=(cond_ok & (it_diff <= 1) & (x_err <= X_RTOL) & (f_err <= F_RTOL)).mean())
    return res


def fmt(res):
    return ("%(label)s: B=%(B)d cond=%(cond_frac).4f iter==%(iter_exact_frac).4f iter+-1=%(iter_pm1_frac).4f "
            "x<=1e-8:%(x_frac).4f f<=1e-10:%(f_frac).4f all=%(all_ok_frac).4f | max x_err %(x_err_max).2e "
            "(same path %(x_err_max_same_path).2e) f_err %(f_err_max).2e lam_err %(lam_err_max).2e status!=0:%(status_nonzero)d"
            % res)
