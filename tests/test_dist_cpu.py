"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: instance / column sharding, the unique-id exchange
and the gather of per-rank results (the device work itself is covered by the -m gpu tests and tools/dist_large_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lfpsqp.jl_b200 import dist as D
    from oracle import oracle as O
    # (1) unique-id exchange: rank 0's bytes reach everybody
    uid = D.exchange_unique_id(lambda: bytes(range(128)), dist, rank)
    assert uid == bytes(range(128))
    # (2) batched mode: shard the instances, solve the shard (oracle stands in for the device), gather, compare
    B, n = 37, 2
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-2, 2, (B, n))
    lo, hi = D.instance_range(B, world, rank)
    xs, obj, olen, lam, term, st = O.optimize_batched("rosenbrock", 2, 0, 0, x0[lo:hi], nthreads=1)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, xs, term["iter"].copy()))
    if rank == 0:
        full = O.optimize_batched("rosenbrock", 2, 0, 0, x0, nthreads=1)
        got = np.concatenate([p[2] for p in sorted(parts)])
        its = np.concatenate([p[3] for p in sorted(parts)])
        assert sorted(p[:2] for p in parts)[0][0] == 0 and sorted(p[:2] for p in parts)[-1][1] == B
        assert np.array_equal(got, full[0]) and np.array_equal(its, full[4]["iter"])
    # (3) large-n mode: column shards tile [0, n) with even widths and the sharded blob reassembles
    nL, m = 70, 3
    Q = rng.standard_normal((m, nL)); A = rng.standard_normal((m, nL)); b = rng.standard_normal(m)
    xt = rng.standard_normal(nL); w = rng.random(nL) + 1
    col0, nloc = D.column_range(nL, world, rank)
    blob = D.shard_diagquad(Q, A, b, xt, w, col0, nloc)
    assert nloc % 2 == 0 and blob.size == 2 * m * nloc + m + 2 * nloc
    shards = [None] * world
    dist.all_gather_object(shards, (col0, nloc, blob))
    if rank == 0:
        shards.sort(key=lambda s: s[0])
        assert shards[0][0] == 0 and sum(s[1] for s in shards) == nL
        Qr = np.concatenate([s[2][:m * s[1]].reshape(m, s[1]) for s in shards], axis=1)
        assert np.array_equal(Qr, Q)
        # the exchange step of the path: partial row sums all-reduce to the full J v
        v = rng.standard_normal(nL)
    # (4) the all-reduce the large-n path relies on, on this backend: t = sum_g J_g v_g
    import torch
    v = np.arange(nL, dtype=np.float64) / nL
    t = torch.from_numpy((Q[:, col0:col0 + nloc] * v[col0:col0 + nloc]) @ np.ones(nloc) + 0.0)
    dist.all_reduce(t)
    assert np.allclose(t.numpy(), (Q * v) @ np.ones(nL), rtol=1e-13)
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_world2_gloo_sharding_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_ranges_cover_and_balance():
    from lfpsqp.jl_b200 import dist as D
    for B in (1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            r = [D.instance_range(B, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    for n in (2, 70, 65536):
        for world in (1, 2, 4, 8):
            c = [D.column_range(n, world, k) for k in range(world)]
            assert c[0][0] == 0 and sum(x[1] for x in c) == n and all(x[1] % 2 == 0 for x in c)
    with pytest.raises(Exception):
        D.column_range(7, 2, 0)
    # Thomson shards whole points (3 coordinates each), an even number of them per rank
    for npts in (2, 64, 4096, 4098):
        for world in (1, 2, 4, 8):
            if npts // 2 < world:
                continue
            c = [D.thomson_point_range(npts, world, k) for k in range(world)]
            assert c[0][0] == 0 and sum(x[1] for x in c) == 3 * npts
            assert all(x[1] % 6 == 0 and x[0] % 6 == 0 for x in c)                       # whole points, even count
            assert all(c[i][0] + c[i][1] == c[i + 1][0] for i in range(world - 1))      # contiguous
            assert max(x[1] for x in c) - min(x[1] for x in c) <= 6
    with pytest.raises(Exception):
        D.thomson_point_range(7, 2, 0)
