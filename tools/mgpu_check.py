"""Multi-GPU parity check (run under torchrun, one rank per GPU): reads sharded over ranks, canonical k-mers routed to
their owners (fused P2P exchange, or NCCL all-to-all with ZB_EXCHANGE=nccl), counted per owner; the union of the
per-rank sets must equal the oracle's kmerize of all reads.  Also shards an all-pairs distance matrix."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from zotmer_b200 import _native as nat, multigpu
from tools import synth
from oracle import c_oracle as co

rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % dev))
K = int(os.environ.get("ZB_CHECK_K", 25))   # 31 = config[4]'s k (62-bit keys)
g = synth.genome(300000, seed=5)
shards = [synth.fastq_array(g, 20000, seed=50 + r).reshape(-1).tobytes() for r in range(world)]
mode = os.environ.get("ZB_EXCHANGE", "p2p")
# p2p: slots agreed on beforehand through a count matrix (the default); p2p_reserve: runs reserved in the owner's buffer by
# remote atomics
p2p = multigpu.P2PExchange(nat, dist, rank, world, dev, 126 * 20000 * 2, reserve=(mode == "p2p_reserve")) if mode.startswith("p2p") else None
ctx = {"world": world, "rank": rank, "dev": dev, "a2a_ms": [], "a2a_bytes": [],
       "send": torch.empty(16, dtype=torch.int64, device="cuda:%d" % dev),
       "recv": torch.empty(16, dtype=torch.int64, device="cuda:%d" % dev)}
for it in range(3):   # several steps: the double-buffered receive side is re-used
    km = nat.Kmerizer(K, dev)
    if p2p is not None and it != 1:
        p2p.prepare(km)          # owner tallies during extraction (step 1 keeps the counting pass)
    km.feed(shards[rank], False)
    if p2p is not None:
        p2p.exchange(km)
    else:
        multigpu.exchange_pending(nat, km, ctx)
    s, nr = km.finish()
    km.close()
    ks, cs = s.fetch()
    if it == 0:
        # one sorted set on rank 0 -> the file streams must be the single-process oracle's, byte for byte
        whole = multigpu.gather_counted_set(nat, s, dist, rank, world, dev)
        if rank == 0:
            ek, ec, _, _ = co.kmerize(K, [(sh, False) for sh in shards])
            wk, wc = whole.fetch()
            assert np.array_equal(wk, ek) and np.array_equal(wc, ec), "gathered set differs from the oracle"
            kw, cw = whole.encode()
            assert np.array_equal(kw, co.encode(ek, True)) and np.array_equal(cw, co.encode(ec.astype(np.uint64), False))
            if whole is not s:
                whole.free()
    s.free()
    parts = [None] * world
    dist.all_gather_object(parts, (ks, cs))
    if rank == 0:
        gk = np.concatenate([p[0] for p in parts]); gc = np.concatenate([p[1] for p in parts])
        assert len(np.unique(gk)) == len(gk), "rank shares overlap"
        order = np.argsort(gk)
        ek, ec, _, _ = co.kmerize(K, [(sh, False) for sh in shards])
        assert np.array_equal(gk[order], ek) and np.array_equal(gc[order], ec), "multi-GPU kmerize differs from the oracle"
if p2p is not None:
    # ONE kmerizer streamed through several exchanges (bench.py's human leg, tools/human_scale.py), with a rank that is
    # late on some steps: a fast peer must not route step s + 2 into a receive buffer whose step-s keys are still being
    # sorted (the reserve mode had no collective that held peers back)
    import time
    km = p2p.prepare(nat.Kmerizer(K, dev))
    for it in range(5):
        km.feed(shards[rank], False)
        if rank == (it % world) and it % 2 == 1:
            time.sleep(0.3)
        p2p.exchange(km)
    s, nr = km.finish()
    km.close()
    ks, cs = s.fetch()
    s.free()
    parts = [None] * world
    dist.all_gather_object(parts, (ks, cs))
    if rank == 0:
        gk = np.concatenate([p[0] for p in parts]); gc = np.concatenate([p[1] for p in parts])
        order = np.argsort(gk)
        ek, ec, _, _ = co.kmerize(K, [(sh, False) for sh in shards] * 5)
        assert np.array_equal(gk[order], ek) and np.array_equal(gc[order], ec), "streamed multi-step exchange differs from the oracle"
    p2p.close()
# all-pairs shards
rng = np.random.default_rng(1)
pool = rng.integers(0, 2 ** 50, 60000, dtype=np.uint64)
arrs = [np.unique(pool[rng.integers(0, len(pool), 20000)]) for _ in range(21)]
sets = [nat.KmerSet.from_arrays(a, device=dev) for a in arrs]
full = multigpu.allpairs_sharded(lambda b, e, st: nat.allpairs_abc(sets, b, e, st), len(sets), dist, rank, world, "cuda:%d" % dev)
# zot merge sharded by key range: identical to the single-GPU merge (and the oracle)
cnts = [rng.integers(1, 1000, len(a), dtype=np.uint32) for a in arrs[:9]]
msets = [nat.KmerSet.from_arrays(a, c, device=dev) for a, c in zip(arrs[:9], cnts)]
merged = multigpu.merge_sharded(nat, msets, dist, rank, world, dev)
if rank == 0:
    mk, mc = merged.fetch()
    ek, ec = co.merge([(a, c.astype(np.uint64)) for a, c in zip(arrs[:9], cnts)])
    assert np.array_equal(mk, ek) and np.array_equal(mc.astype(np.uint64), ec), "sharded merge differs from the oracle"
if rank == 0:
    I, J = np.triu_indices(len(sets), 1)
    for p in range(0, len(I), 7):
        assert tuple(int(v) for v in full[p]) == co.split(arrs[I[p]], arrs[J[p]])
    print("mgpu_check ok: world=%d exchange=%s" % (world, mode))
dist.barrier()
dist.destroy_process_group()
