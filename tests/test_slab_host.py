"""CPU, world_size 2 over gloo: the host-side logic of the slab decomposition (row partition,
IPC-handle exchange plumbing, exact MAX reduction of residual bit patterns)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from svirl_b200.parallel.slab import partition_rows, max_reduce_u64


def test_partition_rows_covers_grid():
    for Ny in (4, 129, 2048, 65536):
        for w in (1, 2, 3, 8):
            parts = partition_rows(Ny, w)
            assert parts[0][0] == 0 and parts[-1][1] == Ny
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # residuals of three sweeps as bit patterns: rank-local maxima differ
        r = np.array([[3e-2, 4e-4, 9.9e-7], [2e-2, 7e-4, 1.1e-6]][rank], dtype=np.float64)
        bits = r.view(np.uint64).copy()
        max_reduce_u64(bits)
        got = bits.view(np.float64)
        ok = np.array_equal(got, np.array([3e-2, 7e-4, 1.1e-6]))
        # handle exchange plumbing: every rank learns its neighbours' 144-byte blobs and row ranges
        blob = bytes([rank]) * 144
        parts = partition_rows(100, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, (blob, parts[rank]))
        ok = ok and gathered[1 - rank][0] == bytes([1 - rank]) * 144 and gathered[1 - rank][1] == parts[1 - rank]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_max_reduce_and_handle_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_slab_split_plan_keeps_the_interior_at_single_gpu_rounds():
    """svl_slab_split_plan (host arithmetic behind option slab_bnd = 0): how many CTAs the boundary tile rows of a slab
    batch get.  For the weak-scaling workload (2048 x 2048 nodes per GPU, 56 x 24 output tiles, 296 resident CTAs) the
    interior launch must still finish in the 11 rounds a single GPU needs, and the boundary CTAs (tiles + 3 tile times
    of fence) well before it; the call must be safe for degenerate sizes."""
    from svirl_b200 import _lib
    lib = _lib.load()
    plan = lib.svl_slab_split_plan

    def rounds(nb, ni, slots, t):
        return -(-nb // t) + 3, -(-ni // (slots - t))

    ntx, nty, slots = 37, 86, 296
    assert -(-ntx * nty // slots) == 11                       # one GPU: 3182 tiles on 296 CTAs
    for nbrows in (1, 2):                                      # one neighbour (end ranks) / two neighbours
        nb, ni = ntx * nbrows, ntx * (nty - nbrows)
        t = plan(nb, ni, slots)
        rb, ri = rounds(nb, ni, slots, t)
        assert 1 <= t <= nb and ri == 11 and rb <= 0.85 * ri, (nbrows, t, rb, ri)
        # no other split is better
        for u in range(1, min(nb, slots // 2) + 1):
            ub, ui = rounds(nb, ni, slots, u)
            cost = lambda b, i: i if 20 * b <= 17 * i else (20 * b + 16) // 17      # noqa: E731
            assert cost(rb, ri) <= cost(ub, ui)
    # strong scaling of 32768^2 fp64 on 8 GPUs: 586 tiles per tile row, 148 resident CTAs
    t = plan(2 * 586, 170 * 586, 148)
    rb, ri = rounds(2 * 586, 170 * 586, 148, t)
    assert rb <= 0.85 * ri and ri <= 1.02 * (172 * 586 / 148)
    # degenerate sizes
    assert plan(0, 10, 296) >= 1 and plan(5, 0, 296) >= 1 and plan(3, 3, 2) >= 1 and plan(12, 18, 296) <= 12
