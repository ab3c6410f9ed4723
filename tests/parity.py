"""Shared parity checker: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

north_star tolerances (BASELINE.json): same termination condition, iteration count within +-1, final x within 1e-8
relative, objective within 1e-10 relative, all Float64.

Two levels:
  compare_batch  per-criterion pass fractions + worst errors (summary of a batch)
  classify       per-INSTANCE verdict.  An instance outside the tolerances is re-judged between independent builds of the
                 ORACLE ITSELF (plain / +FMA contraction / left-to-right summation, oracle/Makefile): if those builds of the
                 same reference algorithm also disagree on it beyond the same tolerances, the instance is tagged
                 "rounding-sensitive" (no implementation can match the reference on it); otherwise it is a FAILURE.
                 record() appends the verdict to the parity record (profiles/parity_r2.json is a copy of one GPU run).
"""
import json
import os

import numpy as np

X_RTOL = 1e-8
F_RTOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _final_obj(obj, olen):
    H = obj.shape[1]
    idx = np.minimum(olen.astype(np.int64), H) - 1
    return obj[np.arange(obj.shape[0]), idx]


def _obj_scale(obj, olen):
    H = obj.shape[1]
    mask = np.arange(H)[None, :] < np.minimum(olen, H)[:, None]
    a = np.where(mask, np.abs(obj), 0.0)
    return np.nanmax(a, axis=1)


def errors(gpu, orc):
    """Per-instance criteria of one result tuple against another: (cond_ok, it_diff, x_err, f_err, lam_err)."""
    gx, gobj, glen, glam, gterm = gpu[:5]
    ox, oobj, olen, olam, oterm = orc[:5]
    cond_ok = gterm["condition"] == oterm["condition"]
    it_diff = np.abs(gterm["iter"].astype(np.int64) - oterm["iter"].astype(np.int64))
    xs = np.maximum(np.linalg.norm(ox, axis=1), 1e-300)
    x_err = np.linalg.norm(gx - ox, axis=1) / xs
    H = min(gobj.shape[1], oobj.shape[1])
    gf = _final_obj(gobj[:, :H], glen)
    of = _final_obj(oobj[:, :H], olen)
    # objective: 1e-10 relative, with a floor of a few ulps of the objective's scale along the trajectory (an optimum
    # with f* ~ 0, e.g. Rosenbrock, has no meaningful relative error: the reference's own rounding exceeds it)
    fscale = _obj_scale(oobj[:, :H], olen)
    f_err = np.abs(gf - of) / (np.abs(of) + 8 * 2.2e-16 * fscale / F_RTOL + 1e-300)
    lam_err = np.zeros(gx.shape[0])
    if glam.size:
        lam_err = np.linalg.norm(glam - olam, axis=1) / np.maximum(np.linalg.norm(olam, axis=1), 1e-300)
    return cond_ok, it_diff, x_err, f_err, lam_err


def within(gpu, orc):
    cond_ok, it_diff, x_err, f_err, _ = errors(gpu, orc)
    return cond_ok & (it_diff <= 1) & (x_err <= X_RTOL) & (f_err <= F_RTOL)


def compare_batch(gpu, orc, n_user, label=""):
    """gpu/orc: tuples (x, obj, olen, lam, term).  Returns a dict of per-criterion pass fractions + worst errors."""
    gterm = gpu[4]
    B = gpu[0].shape[0]
    cond_ok, it_diff, x_err, f_err, lam_err = errors(gpu, orc)
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


def subset(res, idx):
    return tuple(a[idx] for a in res[:5])


def classify(gpu, oracle_runs, label, extra=None, probe=None):
    """Per-instance verdict.  oracle_runs: dict build-name -> result tuple, first entry = the reference build ("base").
    probe(idx) (optional) re-runs the BASE oracle on the instances idx with every input moved by one ulp (up, then down)
    and returns the result tuples: an instance whose result moves beyond the tolerances under a 1-ulp input change is
    ill-conditioned as a map input -> result (condition number > tolerance / 2^-52), i.e. rounding-sensitive by definition.
    Returns a JSON-able record; record["failures"] must be 0 for parity to be green."""
    names = list(oracle_runs)
    base = oracle_runs[names[0]]
    cond_ok, it_diff, x_err, f_err, lam_err = errors(gpu, base)
    ok = cond_ok & (it_diff <= 1) & (x_err <= X_RTOL) & (f_err <= F_RTOL)
    B = ok.size
    # rounding sensitivity of the reference algorithm itself on each instance: any two oracle builds disagree
    sens = np.zeros(B, dtype=bool)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            sens |= ~within(oracle_runs[names[j]], oracle_runs[names[i]])
    # a GPU result that matches ANY build of the oracle within the tolerances is on the reference's path too
    ok_any = ok.copy()
    for nm in names[1:]:
        ok_any |= within(gpu, oracle_runs[nm])
    bad = ~ok
    tagged = bad & (sens | ok_any)
    failed = bad & ~tagged
    probed = 0
    if probe is not None and failed.any():
        idx = np.nonzero(failed)[0]
        bsub = subset(base, idx)
        moved = np.zeros(idx.size, dtype=bool)
        for res in probe(idx):
            moved |= ~within(res, bsub)
        probed = int(idx.size)
        tagged[idx[moved]] = True
        failed = bad & ~tagged
    rec = dict(label=label, instances=int(B), oracle_builds=names,
               cond_frac=float(cond_ok.mean()), iter_exact_frac=float((it_diff == 0).mean()),
               iter_pm1_frac=float((it_diff <= 1).mean()), x_frac=float((x_err <= X_RTOL).mean()),
               f_frac=float((f_err <= F_RTOL).mean()), within_tolerance=int(ok.sum()),
               within_tolerance_frac=float(ok.mean()),
               rounding_sensitive_in_oracle=int(sens.sum()),
               outside_tolerance=int(bad.sum()), tagged_rounding_sensitive=int(tagged.sum()), failures=int(failed.sum()),
               probed_with_one_ulp_input_change=probed,
               x_err_max=float(np.nanmax(x_err)), x_err_max_within=float(x_err[ok].max()) if ok.any() else None,
               f_err_max=float(np.nanmax(f_err)), lam_err_max=float(np.nanmax(lam_err)),
               status_nonzero=int((gpu[4]["status"] != 0).sum()),
               failing_instances=[int(k) for k in np.nonzero(failed)[0][:32]],
               failing_detail=[dict(k=int(k), cond_gpu=int(gpu[4]["condition"][k]), cond_orc=int(base[4]["condition"][k]),
                                    it_gpu=int(gpu[4]["iter"][k]), it_orc=int(base[4]["iter"][k]), x_err=float(x_err[k]),
                                    f_err=float(f_err[k])) for k in np.nonzero(failed)[0][:32]],
               tagged_detail=[dict(k=int(k), it_gpu=int(gpu[4]["iter"][k]), it_orc=int(base[4]["iter"][k]),
                                   x_err=float(x_err[k]), f_err=float(f_err[k])) for k in np.nonzero(tagged)[0][:8]])
    if extra:
        rec.update(extra)
    return rec


def fmt_record(rec):
    return ("%(label)s: %(instances)d instances | within tolerance %(within_tolerance)d (%(within_tolerance_frac).6f) | "
            "cond %(cond_frac).6f iter==%(iter_exact_frac).6f iter+-1 %(iter_pm1_frac).6f x %(x_frac).6f f %(f_frac).6f | "
            "outside %(outside_tolerance)d = rounding-sensitive %(tagged_rounding_sensitive)d + FAILURES %(failures)d | "
            "max x_err %(x_err_max).2e f_err %(f_err_max).2e" % rec)


def record_path():
    """Where the GPU run writes its parity record: gpurun_out/ (merged back from the GPU box) or $LFPSQP_PARITY_OUT."""
    p = os.environ.get("LFPSQP_PARITY_OUT")
    if p:
        return p
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "parity_r2.json")


def record(rec):
    path = record_path()
    try:
        data = json.load(open(path))
    except Exception:
        data = {"criteria": {"condition": "equal", "iterations": "+-1", "x_rel": X_RTOL, "f_rel": F_RTOL,
                             "classification": "outside tolerance -> rounding-sensitive iff two builds of the oracle "
                                               "(plain / fma / seq summation) disagree on that instance beyond the same "
                                               "tolerances, or the GPU matches one of those builds, or the plain oracle's "
                                               "own result moves beyond the tolerances when its input is changed by one "
                                               "ulp; else FAILURE"},
                "records": {}}
    data["records"][rec["label"]] = rec
    data["failures_total"] = int(sum(r.get("failures", 0) for r in data["records"].values()))
    data["parity"] = "green" if data["failures_total"] == 0 else "red"
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(fmt_record(rec) if "within_tolerance" in rec else json.dumps(rec))
    return rec
