"""Generates tests/golden/*.json: known answers for the hot path.

Two kinds of entries:
 * "reference": values printed in the reference's own README (README.md:31-37, Rosenbrock from x0 = [0, 0]) -- the only
   end-to-end known answer the reference holds (Julia is not installed here, so the reference itself cannot be run);
 * "oracle": outputs of the CPU oracle (oracle/lfpsqp_oracle.cpp, -ffp-contract=off build) on small seeded inputs for
   every benchmark family, committed so that both the oracle (CPU test) and the CUDA path (GPU test) are checked
   against fixed vectors and neither can drift silently.
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def diagquad(n, m, seed, cond):
    rng = np.random.Generator(np.random.Philox(key=seed))
    Q = rng.standard_normal((m, n)) / np.sqrt(n); A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.standard_normal(n); xt = rng.standard_normal(n)
    w = np.exp(rng.uniform(0.0, np.log(cond), n))
    b = 0.5 * Q @ (x0 * x0) + A @ x0
    return Q, A, b, xt, w, x0


def entry(name, family, n, m, p, x0, xl=None, xu=None, fam_params=None, **prm):
    x, obj, lam, term, stats = O.optimize(family, n, m, p, x0, xl=xl, xu=xu, fam_params=fam_params,
                                          params=O.default_params(**prm) if prm else None)
    return dict(name=name, family=family, n=n, m=m, p=p, params=prm, x0=list(map(float, x0)),
                xl=None if xl is None else [float(v) for v in xl], xu=None if xu is None else [float(v) for v in xu],
                fam_params=None if fam_params is None else [float(v) for v in np.ravel(fam_params)],
                x=list(map(float, x)), obj_values=list(map(float, obj)), lam=list(map(float, lam)),
                condition=int(term["condition"]), iter=int(term["iter"]), f_diff=float(term["f_diff"]),
                step_diff=float(term["step_diff"]), kkt_diff=float(term["kkt_diff"]))


def main():
    O.build(); O.lib()
    cases = []
    cases.append(entry("rosenbrock_readme", "rosenbrock", 2, 0, 0, np.zeros(2)))
    cases.append(entry("readme_equality", "readme_eq", 50, 1, 0, np.ones(50)))
    co = np.random.default_rng(0).standard_normal(50)
    inf = np.inf * np.ones(50)
    cases.append(entry("readme_inequality", "readme_ineq", 50, 0, 1, np.zeros(50), xl=-inf, xu=inf, fam_params=co))
    rng = np.random.default_rng(6)
    pts = rng.standard_normal((12, 3)); pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    cases.append(entry("thomson_12", "thomson", 36, 12, 0, pts.ravel()))
    Q, A, b, xt, w, x0 = diagquad(64, 4, 5, 50.0)
    blob = np.concatenate([Q.ravel(), A.ravel(), b, xt, w])
    cases.append(entry("diagquad_64_4_pp", "diagquad", 64, 4, 0, x0, fam_params=blob))
    cases.append(entry("diagquad_64_4_nr", "diagquad", 64, 4, 0, x0, fam_params=blob, do_project_retract=0))
    cases.append(entry("diagquad_64_4_box", "diagquad", 64, 4, 0, x0, xl=x0 - 0.4, xu=x0 + 0.6, fam_params=blob))
    t = np.random.default_rng(3).standard_normal(12)
    cases.append(entry("sin_12_3", "sin", 12, 3, 0, np.zeros(12), fam_params=t))
    out = {"reference": {"source": "/root/reference/README.md:31-37 (Rosenbrock, x0 = [0, 0], default parameters)",
                         "condition": "f_tol", "f_diff": 1.0898882046786806e-7, "step_diff": 0.0007384068067118611,
                         "kkt_diff": 4.332627751789361e-5, "iters": 17},
           "oracle": cases}
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
