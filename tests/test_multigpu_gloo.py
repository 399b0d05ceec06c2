"""
Host-side logic of the multi-GPU kmerize path, exercised on the CPU with the gloo backend at
world_size 2 (and 3): ownership is a partition, the sizes-then-payload all-to-all delivers every key to
its owner exactly once, and counting the received shares reproduces the single-process oracle counts.
"""
import os
import socket

import numpy as np
import pytest

from zotmer_b200 import multigpu


def test_owner_function_is_a_balanced_partition():
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 2 ** 50, 200000, dtype=np.uint64)
    low = np.arange(100000, dtype=np.uint64)          # low-complexity: consecutive small integers
    for n in (1, 2, 3, 4, 8):
        for ks in (keys, low):
            ow = multigpu.owner_of(ks, n)
            assert ow.min() >= 0 and ow.max() < n
            share = np.bincount(ow, minlength=n) / len(ks)
            assert np.all(np.abs(share - 1.0 / n) < 0.02), (n, share)
    # exact 128-bit arithmetic cross-check
    for x in [0, 1, 2 ** 64 - 1, 0x123456789abcdef, 2 ** 63]:
        h = int(multigpu.mix64(np.array([x], np.uint64))[0])
        for n in (2, 3, 8):
            assert int(multigpu.owner_of(np.array([x], np.uint64), n)[0]) == (h * n) >> 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds a different shard of keys drawn from a common pool (lots of duplicates)
        pool = np.random.default_rng(7).integers(0, 2 ** 50, 5000, dtype=np.uint64)
        mine = pool[np.random.default_rng(100 + rank).integers(0, len(pool), 40000)]
        grouped, counts = multigpu.bucket_host(mine, world)
        send = torch.from_numpy(grouped.view(np.int64).copy())
        recv, rc = multigpu.exchange_tensors(dist, send, counts, lambda n: torch.empty(max(n, 1), dtype=torch.int64))
        got = recv[:sum(rc)].numpy().view(np.uint64)
        assert np.all(multigpu.owner_of(got, world) == rank)          # only keys I own
        ks, cs = np.unique(got, return_counts=True)
        # gather the per-rank counted shares on rank 0 and compare with counting everything at once
        shares = [None] * world
        dist.all_gather_object(shares, (ks, cs))
        alls = [None] * world
        dist.all_gather_object(alls, mine)
        if rank == 0:
            ek, ec = np.unique(np.concatenate(alls), return_counts=True)
            gk = np.concatenate([s[0] for s in shares])
            gc = np.concatenate([s[1] for s in shares])
            order = np.argsort(gk)
            assert len(np.unique(gk)) == len(gk)                       # shares are disjoint
            assert np.array_equal(gk[order], ek) and np.array_equal(gc[order], ec)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_p2p_offsets_tile_every_receive_buffer():
    """slots of the fused exchange: per owner, the runs of all sources are disjoint, in source order and gap-free"""
    rng = np.random.default_rng(2)
    for world in (1, 2, 3, 8):
        M = rng.integers(0, 1000, (world, world))
        offs = [multigpu.p2p_offsets(M, r) for r in range(world)]
        for dst in range(world):
            pos = 0
            for src in range(world):
                assert offs[src][0][dst] == pos
                pos += int(M[src][dst])
            assert offs[dst][1] == pos


# ------------------------------------------------------------------------------------------------
# all-pairs distance matrix: tiling and sharding
@pytest.mark.parametrize("nsets", [1, 2, 7, 31, 32, 33, 64, 65, 130])
def test_tiles_cover_every_pair_once(nsets):
    """every pair belongs to exactly one tile, and to each of that tile's AP_KS key-range shards (work units)"""
    seen = {}
    assert multigpu.n_tiles(nsets) % multigpu.AP_KS == 0
    for u in range(multigpu.n_tiles(nsets)):
        bi, bj = multigpu.tile_blocks(nsets, u)
        assert bi <= bj
        for (i, j) in multigpu.tile_pairs(nsets, u):
            assert i < j < nsets
            seen.setdefault((i, j), []).append(u)
    assert len(seen) == nsets * (nsets - 1) // 2
    for us in seen.values():
        assert len(us) == multigpu.AP_KS and us == list(range(us[0], us[0] + multigpu.AP_KS)) and us[0] % multigpu.AP_KS == 0
    idx = sorted(multigpu.pair_index(nsets, i, j) for (i, j) in seen)
    assert idx == list(range(len(seen)))
    for world in (1, 2, 3, 8):
        rs = multigpu.tile_ranges(nsets, world)
        assert rs[0][0] == 0 and rs[-1][1] == multigpu.n_tiles(nsets)
        assert all(rs[r][1] == rs[r + 1][0] for r in range(world - 1))
        if nsets >= 2 and world == 8:   # even one tile spreads over 8 ranks (one key-range shard each)
            assert all(e > b for b, e in rs)
        # strided shares: a partition of the units, level to within one unit
        shares = [multigpu.share_units(nsets, r, world) for r in range(world)]
        assert sorted(u for sh in shares for u in sh) == list(range(multigpu.n_tiles(nsets)))
        assert max(len(sh) for sh in shares) - min(len(sh) for sh in shares) <= 1


def _host_tiles(arrs, b, e, stride=1):
    """numpy stand-in for zotmer_b200._native.allpairs_abc(sets, b, e): the cardinalities of every pair of the
    units' tiles, restricted to the k-mers of the units' key-range shards"""
    n = len(arrs)
    key_bits = max([int(a.max()).bit_length() for a in arrs if len(a)] + [1])
    out = np.zeros((n * (n - 1) // 2, 3), np.uint64)
    for u in range(b, e, stride):
        sh = u % multigpu.AP_KS
        part = [a[np.array([multigpu.key_shard(x, key_bits) == sh for x in a], bool)] if len(a) else a for a in arrs]
        for (i, j) in multigpu.tile_pairs(n, u):
            a = len(np.intersect1d(part[i], part[j], assume_unique=True))
            out[multigpu.pair_index(n, i, j)] += np.array((a, len(part[i]) - a, len(part[j]) - a), np.uint64)
    return out


def _pairs_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        pool = rng.integers(0, 2 ** 40, 4000, dtype=np.uint64)
        arrs = [np.unique(pool[rng.integers(0, len(pool), int(rng.integers(0, 300)))]) for _ in range(37)]
        full = multigpu.allpairs_sharded(lambda b, e, st: _host_tiles(arrs, b, e, st), len(arrs), dist, rank, world)
        assert np.array_equal(full, _host_tiles(arrs, 0, multigpu.n_tiles(len(arrs))))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_allpairs_sharded_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_pairs_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {r: "ok" for r in range(world)}
