"""all-pairs on NSETS synthetic sets of CLADES clades (bench.py's generator): sets in interleaved order (set i in clade
i % CLADES, as bench.py builds them) against the same sets ordered clade by clade -- how much does it pay to put related
sets into the same block of 32?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from zotmer_b200 import _native as nat

NSETS = int(os.environ.get("NSETS", 320)); CLADES = int(os.environ.get("CLADES", 10)); NKEYS = int(os.environ.get("NKEYS", 9950000))
dv = torch.device("cuda:0")
g = torch.Generator(device=dv); g.manual_seed(1000)
bases = [torch.unique(torch.randint(0, 1 << 50, (NKEYS,), generator=g, device=dv, dtype=torch.int64)) for _ in range(CLADES)]
sets = []
for i in range(NSETS):
    c = i % CLADES
    q = 0.025 + 0.195 * ((i // CLADES) / max(1, NSETS // CLADES - 1))
    b = bases[c]
    drop = torch.rand(b.numel(), generator=g, device=dv) < q
    fresh = torch.randint(0, 1 << 50, (int(drop.sum().item()),), generator=g, device=dv, dtype=torch.int64)
    keys = torch.unique(torch.cat([b[~drop], fresh]))
    sets.append(nat.KmerSet.from_device(keys.data_ptr(), None, keys.numel(), device=0))
del bases
torch.cuda.synchronize(); torch.cuda.empty_cache()
npairs = NSETS * (NSETS - 1) // 2
for name, order in (("interleaved", list(range(NSETS))), ("by clade", sorted(range(NSETS), key=lambda i: (i % CLADES, i)))):
    ss = [sets[i] for i in order]
    for it in range(2):
        nat.dbg_profile(True)
        abc = nat.allpairs_abc(ss)
        prof = nat.dbg_profile(False)
    print("%-12s %d sets: kernel %.1f ms -> %.0f set-pairs/s" % (name, NSETS, prof["allpairs"][0], npairs / prof["allpairs"][0] * 1e3), flush=True)
