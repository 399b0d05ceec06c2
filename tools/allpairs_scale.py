"""BASELINE.json config[3] at full size on one B200 (or, under torchrun, N of them): all-pairs (|X n Y|, |X \\ Y|, |Y \\ X|) over NSETS = 1,000 synthetic
bacterial k-mer sets of ~10 M k-mers (k = 25: 50-bit keys), 499,500 pairs, through zb_allpairs_abc -- the call behind
`zot dist` / `zot jaccard -a`.

The sets are made on the device (1,000 x 80 MB cannot come through PCIe in a bounded run): CLADES clade bases of
distinct random 50-bit keys; a member keeps every base key with probability 1 - q (q = share of k-mers hit by a
substitution, 2.5 % .. 22 % for 0.1 % .. 1 % per-base divergence at k = 25) and draws fresh random keys for the rest, so
pairs inside a clade share (1 - qi)(1 - qj) of a base and pairs across clades share next to nothing -- the shape
SURVEY.md 8d gives for config 4 of its table.

Checks (size-independent properties + samples, the oracle cannot run 80 GB):
  * a + b = |X_i| and a + c = |X_j| for every pair;
  * SAMPLE random pairs against the pair-at-a-time merge-path kernel (zb_pairs_abc) AND against torch
    (sort of the concatenation, count of adjacent equal keys);
  * inside a clade a / |base| is within 1 % of (1 - qi)(1 - qj); across clades a < 100.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from zotmer_b200 import _native as nat

NSETS = int(os.environ.get("NSETS", 1000))
CLADES = int(os.environ.get("CLADES", 10))
NKEYS = int(os.environ.get("NKEYS", 9950000))
SAMPLE = int(os.environ.get("SAMPLE", 48))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
di = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda:%d" % di)
torch.cuda.set_device(dev)
dist = None
if world > 1:      # torchrun: every rank builds all sets (same seed) and takes its share of the work units
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from zotmer_b200 import multigpu
g = torch.Generator(device=dev)
g.manual_seed(1000)

t0 = time.time()
bases = []
for c in range(CLADES):
    b = torch.unique(torch.randint(0, 1 << 50, (NKEYS,), generator=g, device=dev, dtype=torch.int64))   # sorted
    bases.append(b)
sets, q_of, clade_of = [], [], []
for i in range(NSETS):
    c = i % CLADES
    q = 0.025 + 0.195 * ((i // CLADES) / max(1, NSETS // CLADES - 1)) if NSETS > CLADES else 0.05
    b = bases[c]
    drop = torch.rand(b.numel(), generator=g, device=dev) < q
    fresh = torch.randint(0, 1 << 50, (int(drop.sum().item()),), generator=g, device=dev, dtype=torch.int64)
    keys = torch.unique(torch.cat([b[~drop], fresh]))
    sets.append(nat.KmerSet.from_device(keys.data_ptr(), None, keys.numel(), device=di))
    q_of.append(q)
    clade_of.append(c)
    del keys, fresh, drop
torch.cuda.synchronize()
sizes = np.array([len(s) for s in sets], np.uint64)
if rank == 0:
    print("%d sets (%d clades) of %.2f M k-mers each built on %s in %.1f s; %.1f GB of keys" % (
        NSETS, CLADES, sizes.mean() / 1e6, "the device" if world == 1 else "each of %d devices" % world, time.time() - t0,
        sizes.sum() * 8 / 1e9), flush=True)
torch.cuda.empty_cache()

npairs = NSETS * (NSETS - 1) // 2
ub, ue, ust = multigpu.unit_share(NSETS, rank, world)
if os.environ.get("ZB_CONTIGUOUS"):      # the earlier sharding: one contiguous range of units per rank
    (ub, ue), ust = multigpu.tile_ranges(NSETS, world)[rank], 1
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
nat.dbg_profile(True, di)
t0 = time.time()
abc = nat.allpairs_abc(sets, ub, ue, ust)
t_own = time.time() - t0
if world > 1:      # the "final gather": (a, b, c) restricted to a key-range shard add up over the shards
    t = torch.from_numpy(abc.view(np.int64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    abc = t.cpu().numpy().view(np.uint64)
wall = time.time() - t0
prof = nat.dbg_profile(False, di)
stages = {k: round(v[0], 2) for k, v in prof.items()}
kern = prof["allpairs"][0] if "allpairs" in prof else float("nan")
if world > 1:
    tt = torch.tensor([kern, wall, t_own], dtype=torch.float64, device=dev)
    lo = tt.clone()
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    kern, wall = float(tt[0]), float(tt[1])
    if rank == 0:
        print("%d ranks, work units range(%d, %d, %d) on rank 0; own share: slowest rank %.2f s, fastest %.2f s (kernel %.2f / %.2f s)" % (
            world, ub, ue, ust, float(tt[2]), float(lo[2]), float(tt[0]) / 1e3, float(lo[0]) / 1e3), flush=True)
if rank != 0:
    for x in sets:
        x.free()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
print("stages (ms):", stages)
bytes_model = float(8 * (abc[:, 0].astype(np.float64) * 2 + abc[:, 1] + abc[:, 2]).sum())
print("allpairs: %d pairs in %.2f s per call (kernel %.2f s) -> %.0f set-pairs/s per call, %.0f kernel-timed; "
      "8(|X|+|Y|) B/pair model: %.1f TB -> %.1f TB/s equivalent" % (
          npairs, wall, kern / 1e3, npairs / wall, npairs / kern * 1e3, bytes_model / 1e12, bytes_model / wall / 1e12), flush=True)

# ---- every pair: a + b = |X_i|, a + c = |X_j|
I, J = np.triu_indices(NSETS, 1)
assert np.array_equal(abc[:, 0] + abc[:, 1], sizes[I]) and np.array_equal(abc[:, 0] + abc[:, 2], sizes[J]), "sizes do not add up"
# ---- clade structure
cl = np.array(clade_of)
qq = np.array(q_of)
same = cl[I] == cl[J]
expect = (1 - qq[I]) * (1 - qq[J]) * np.array([b.numel() for b in bases], np.float64)[cl[I]]
rel = np.abs(abc[same, 0].astype(np.float64) - expect[same]) / expect[same]
assert rel.max() < 0.01, rel.max()
assert int(abc[~same, 0].max()) < 100 if (~same).any() else True
print("every pair adds up; %d pairs inside clades within %.3f %% of the expected intersection; across clades max |X n Y| = %d" % (
    int(same.sum()), 100 * rel.max(), int(abc[~same, 0].max()) if (~same).any() else 0), flush=True)
# ---- samples against the pair-at-a-time kernel and torch
rng = np.random.default_rng(5)
pick = rng.choice(npairs, size=min(SAMPLE, npairs), replace=False)
pick[: min(8, len(pick))] = np.flatnonzero(same)[rng.choice(int(same.sum()), size=min(8, len(pick)), replace=False)]
t0 = time.time()
ref = nat.pairs_abc(sets, I[pick], J[pick])
dt = time.time() - t0
assert np.array_equal(ref, abc[pick]), "all-pairs differs from the pair-at-a-time kernel"
for p in pick[:16]:
    xs = [s.fetch(counts=False) for s in (sets[I[p]], sets[J[p]])]
    both = torch.sort(torch.cat([torch.from_numpy(x.view(np.int64)).to(dev) for x in xs])).values
    a = int((both[1:] == both[:-1]).sum().item())
    assert (a, len(xs[0]) - a, len(xs[1]) - a) == tuple(int(v) for v in abc[p]), (p, a, abc[p])
print("%d sampled pairs equal the pair-at-a-time kernel (%.0f set-pairs/s there), 16 of them also torch sort + adjacent-equal count" % (
    len(pick), len(pick) / dt), flush=True)
print("jaccard distance of pair (0, %d): %.6f" % (CLADES, float(abc[CLADES - 1, 1] + abc[CLADES - 1, 2]) / float(abc[CLADES - 1].sum())))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
