"""Full-size, per-instance parity sweep at the BASELINE.json sizes.

Every instance of the bench workloads (same Philox seeds as bench.py) is solved on the GPU through the C ABI and on
the CPU by three independent builds of the oracle (plain IEEE / +FMA contraction / left-to-right summation).  An
instance outside the north-star tolerances is either rounding-sensitive IN THE REFERENCE ALGORITHM ITSELF (two oracle
builds disagree on it) or a FAILURE.  Zero failures are required; the per-criterion fractions and the tagged counts
go to gpurun_out/parity_r2.json (committed copy: profiles/parity_r2.json) and are printed by bench.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu
ROOT = parity.ROOT


@pytest.fixture(scope="module")
def L():
    import lfpsqp.jl_b200 as L
    L.default_context(0)
    return L


def _oracle_builds(oracle, run):
    out = {"base": run()}
    for v in ("fma", "seq"):
        with oracle.variant(v):
            out[v] = run()
    return out


def test_c2_all_65536_bench_instances(L, oracle):
    # BASELINE config C2, the exact instance set bench.py times on rank 0 (bench.make_inputs: Philox key SEED)
    import bench
    B, n = bench.B_PER_GPU, bench.N_VARS
    coeff, x0 = bench.make_inputs(0, B)
    inf = np.inf * np.ones(n)
    fam = L.families.readme_inequality(coeff)
    gpu = L.optimize_batched(fam.f, None, fam.d, x0, -inf, inf, 0, 1, history=bench.HIST, return_stats=True)
    runs = _oracle_builds(oracle, lambda: oracle.optimize_batched("readme_ineq", n, 0, 1, x0, xl=-inf, xu=inf, fam_params=coeff,
                                                                 fam_stride=n, H=bench.HIST, nthreads=bench.host_cores()))
    nc = np.linalg.norm(coeff, axis=1)

    def probe(idx):
        for to in (np.inf, -np.inf):
            yield oracle.optimize_batched("readme_ineq", n, 0, 1, x0[idx], xl=-inf, xu=inf, fam_params=np.nextafter(coeff[idx], to),
                                          fam_stride=n, H=bench.HIST, nthreads=bench.host_cores())
    rec = parity.classify(gpu, runs, "C2 README inequality, all %d bench instances" % B, probe=probe,
                          extra={"max_dist_to_known_solution": float(np.max(np.linalg.norm(gpu[0] + coeff / nc[:, None], axis=1))),
                                 "gpu_stats_equal_oracle": {k: float((gpu[5][k] == runs["base"][5][k]).mean())
                                                            for k in ("retract_outer", "retract_pcg", "armijo_trials", "pp_backtracks",
                                                                      "projcg_negcurv", "f_evals")}})
    parity.record(rec)
    assert rec["status_nonzero"] == 0
    assert rec["failures"] == 0, parity.fmt_record(rec)
    assert rec["within_tolerance_frac"] >= 0.999, parity.fmt_record(rec)


def test_c3_all_1048576_bench_instances(L, oracle):
    # BASELINE config C3 (bench.py extras_section: Philox key SEED+3, instance 0 = README start)
    import bench
    B, H = 1 << 20, 64
    rng = np.random.Generator(np.random.Philox(key=bench.SEED + 3))
    x0 = rng.uniform(-2.0, 2.0, (B, 2)); x0[0] = 0.0
    fam = L.families.rosenbrock()
    gpu = L.optimize_batched(fam.f, x0, history=H)
    runs = _oracle_builds(oracle, lambda: oracle.optimize_batched("rosenbrock", 2, 0, 0, x0, H=H, nthreads=bench.host_cores()))
    t0 = gpu[4][0]

    def probe(idx):
        for to in (np.inf, -np.inf):
            yield oracle.optimize_batched("rosenbrock", 2, 0, 0, np.nextafter(x0[idx], to), H=H, nthreads=bench.host_cores())
    rec = parity.classify(gpu, runs, "C3 Rosenbrock, all %d bench instances" % B, probe=probe,
                          extra={"golden_instance0": {"iter": int(t0["iter"]), "condition": int(t0["condition"]), "f_diff": float(t0["f_diff"])}})
    parity.record(rec)
    assert int(t0["iter"]) == 17 and int(t0["condition"]) == 0 and abs(float(t0["f_diff"]) - 1.0898882046786806e-7) < 1e-15
    assert rec["status_nonzero"] == 0
    assert rec["failures"] == 0, parity.fmt_record(rec)
    assert rec["within_tolerance_frac"] >= 0.999, parity.fmt_record(rec)


def _single(x, obj, lam, cond, it, status=0):
    """wrap one large-n solve as a batch of one for parity.classify"""
    term = np.zeros(1, dtype=[("condition", "<i4"), ("status", "<i4"), ("f_diff", "<f8"), ("step_diff", "<f8"), ("kkt_diff", "<f8"), ("iter", "<i8")])
    term["condition"] = cond; term["iter"] = it; term["status"] = status
    return (x[None, :], np.asarray(obj)[None, :], np.array([len(obj)]), np.asarray(lam)[None, :], term)


def _hpad(t, H):
    o = np.full((1, H), np.nan); o[0, :t[1].shape[1]] = t[1][0]
    return (t[0], o, t[2], t[3], t[4])


def _large_vs_oracle(L, label, fam, x0, runs_raw, oracle_stats):
    """runs_raw: build -> (x, obj, lam, condition, iter) of the oracle (live run or committed fixture)."""
    P = L.LargeProblem(fam)
    x, obj, lam, info, st, status = P.solve(x0, L.LFPSQPParams(), return_stats=True, history=20000)
    gpu = _single(x, obj, lam, int(info.condition), info.iter, status)
    runs = {k: _single(*v) for k, v in runs_raw.items()}
    H = max([gpu[1].shape[1]] + [r[1].shape[1] for r in runs.values()])
    gpu = _hpad(gpu, H); runs = {k: _hpad(v, H) for k, v in runs.items()}
    rec = parity.classify(gpu, runs, label, extra={"gpu_iter": int(info.iter), "gpu_condition": info.condition.name,
                                                    "gpu_stats": st, "oracle_stats": oracle_stats, "phase_ms": P.phase_ms()})
    parity.record(rec)
    assert rec["failures"] == 0, parity.fmt_record(rec)
    assert rec["status_nonzero"] == 0
    return rec


def _live(oracle, family, n, m, x0, blob):
    runs, stats = {}, None
    for v in ("base", "fma", "seq"):
        if v == "base":
            r = oracle.optimize(family, n, m, 0, x0, fam_params=blob)
            stats = {k: (float(q) if k == "flops" else int(q)) for k, q in r[4].items()}
        else:
            with oracle.variant(v):
                r = oracle.optimize(family, n, m, 0, x0, fam_params=blob)
        runs[v] = (r[0], r[1], r[2], r[3]["condition"], r[3]["iter"])
    return runs, stats


def _fixture(name):
    """oracle results committed by tests/golden/make_golden_large.py (an SVD-based solve of these sizes takes minutes)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "large_midsize.npz"))
    runs = {}
    for v in ("base", "fma", "seq"):
        t = z["%s/%s/term" % (name, v)]
        runs[v] = (z["%s/%s/x" % (name, v)], z["%s/%s/obj" % (name, v)], z["%s/%s/lam" % (name, v)], int(t[0]), int(t[1]))
    keys = ("projcg_iters", "projcg_negcurv", "armijo_trials", "retract_outer", "retract_pcg", "pp_backtracks", "newton_accepted", "svd_calls", "f_evals")
    return runs, {k: int(q) for k, q in zip(keys, z["%s/base/stats" % name])}


def _golden_inputs(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_large", os.path.join(ROOT, "tests", "golden", "make_golden_large.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod.inputs(name)


def test_c4_thomson_256_full_solve(L, oracle):
    # BASELINE config C4 family at N = 256 points (n = 768, m = 256), seeded as bench.py's C4 (Philox key SEED+4); oracle run live
    import bench
    npts = 256
    rng = np.random.Generator(np.random.Philox(key=bench.SEED + 4))
    p0 = rng.standard_normal((npts, 3)); p0 /= np.linalg.norm(p0, axis=1, keepdims=True)
    runs, ost = _live(oracle, "thomson", 3 * npts, npts, p0.ravel(), None)
    _large_vs_oracle(L, "C4 Thomson N=256 (n=768, m=256) full solve, large-n mode", L.families.thomson(npts), p0.ravel(), runs, ost)


def test_c4_thomson_512_full_solve(L):
    fam, n, m, x0, blob = _golden_inputs("thomson512")
    runs, ost = _fixture("thomson512")
    _large_vs_oracle(L, "C4 Thomson N=512 (n=1536, m=512) full solve, large-n mode", L.families.thomson(512), x0, runs, ost)


def test_c5_family_8192x512_full_solve(L):
    # BASELINE config C5 family (DESIGN.md section 6 definition) at n = 8192, m = 512 against the committed oracle vectors
    fam, n, m, x0, blob = _golden_inputs("diagquad_8192x512")
    runs, ost = _fixture("diagquad_8192x512")
    Q = blob[:m * n].reshape(m, n); A = blob[m * n:2 * m * n].reshape(m, n)
    b = blob[2 * m * n:2 * m * n + m]; xt = blob[2 * m * n + m:2 * m * n + m + n]; w = blob[2 * m * n + m + n:]
    _large_vs_oracle(L, "C5 family n=8192 m=512 full solve, large-n mode", L.families.diagquad(Q, A, b, xt, w), x0, runs, ost)


def test_column_sharded_world2_vs_oracle():
    # the N > 1 large-n path (column shards + in-kernel peer all-reduce + NCCL Gram) against the oracle: self-spawns two
    # ranks when >= 2 GPUs are visible (the driver's 1-GPU box skips; `gpurun --gpus 2` runs it)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dist_large_check.py"), "record"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-4000:]); print(out.stderr[-2000:])
    assert out.returncode == 0
    assert "MISMATCH" not in out.stdout and out.stdout.count(" OK") >= 3
