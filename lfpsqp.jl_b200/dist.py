"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for rendezvous / the unique-id broadcast.

* batched mode: instances are independent -> contiguous instance ranges per rank, no collective (`instance_range`)
* large-n mode: J and every n-vector are sharded by columns (`column_range`); the library's own NCCL communicator
  (lfpsqp_comm_init) all-reduces only the m x m Gram, m-vectors and packed scalars.
"""
import ctypes as C
import glob
import os

import numpy as np

from . import _lib


def instance_range(B, world, rank):
    """Contiguous, balanced [lo, hi) instance range of `rank` (batched mode)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def column_range(n, world, rank):
    """Contiguous [col0, col0 + n_loc) column shard with an EVEN n_loc on every rank (128-bit loads need even row
    lengths); the remainder pairs go to the first ranks, one each.  Requires n even."""
    if n % 2:
        raise _lib.LFPSQPError("large-n column sharding needs an even n")
    pairs = n // 2
    base, rem = divmod(pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return 2 * lo, 2 * (hi - lo)


def thomson_point_range(npoints, world, rank):
    """Thomson shards whole points with an even count per rank: returns (col0, n_loc) = 3 x the owned point range."""
    pairs = npoints // 2
    if npoints % 2:
        raise _lib.LFPSQPError("column-sharded Thomson needs an even number of points")
    base, rem = divmod(pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return 6 * lo, 6 * (hi - lo)


def find_nccl():
    """Path of the NCCL library torch itself uses (nvidia-nccl wheel), or None to let the loader search."""
    try:
        import nvidia.nccl
        c = glob.glob(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2"))
        if c:
            return c[0]
    except Exception:
        pass
    return None


def exchange_unique_id(make_id, dist, rank):
    """Rank 0 creates the 128-byte id, everybody receives it (works with the gloo and the nccl backend)."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def init_comm(ctx, dist=None, peer=True):
    """Create the library's NCCL communicator for this process group. Returns (rank, world)."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    path = find_nccl()
    cpath = path.encode() if path else None

    def make_id():
        buf = C.create_string_buffer(128)
        rc = ctx.lib.lfpsqp_comm_unique_id(buf, cpath)
        if rc != 0:
            raise _lib.LFPSQPError("lfpsqp_comm_unique_id failed (rc=%d)" % rc)
        return bytes(buf.raw)

    uid = exchange_unique_id(make_id, dist, rank)
    ctx.check(ctx.lib.lfpsqp_comm_init(ctx.h, rank, world, C.create_string_buffer(uid, 128), cpath))
    if peer and 2 <= world <= 8:
        # peer-memory all-reduce for the small messages: exchange CUDA IPC handles (rank order), map, barrier
        hb = C.create_string_buffer(64)
        ok = ctx.lib.lfpsqp_comm_ipc_export(ctx.h, hb) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(hb.raw) if ok else None)
        if all(h is not None for h in handles):
            rc = ctx.lib.lfpsqp_comm_ipc_import(ctx.h, C.create_string_buffer(b"".join(handles), 64 * world))
            oks = [None] * world
            dist.all_gather_object(oks, rc == 0)
            if not all(oks):      # all or nothing: a mixed mode would deadlock the flag protocol
                ctx.lib.lfpsqp_comm_destroy(ctx.h)
                ctx.check(ctx.lib.lfpsqp_comm_init(ctx.h, rank, world, C.create_string_buffer(uid2 := exchange_unique_id(make_id, dist, rank), 128), cpath))
        dist.barrier()
    return rank, world


def shard_diagquad(Q, A, b, xt, w, col0, n_loc):
    """This rank's DIAGQUAD parameter blob [Q_loc, A_loc, b, xt_loc, w_loc]."""
    sl = slice(col0, col0 + n_loc)
    return np.concatenate([np.ascontiguousarray(Q[:, sl]).ravel(), np.ascontiguousarray(A[:, sl]).ravel(),
                           np.asarray(b, float).ravel(), np.asarray(xt, float)[sl], np.asarray(w, float)[sl]])
