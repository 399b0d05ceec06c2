"""sort+count timing on config-2-like keys (canonical k=25 k-mers of 30x reads): segmented vs classic"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
from tools import synth
nreads = int(os.environ.get("NREADS", 1000000))
g = synth.genome(5000000)
fq = synth.fastq_array(g, nreads).reshape(-1)
codes, nr = nat.dbg_parse(fq.tobytes(), False)
keys = nat.dbg_extract(25, codes)
print("keys", len(keys), flush=True)
for mode in (0, 3, 1):
    k, c, ms = nat.dbg_sort_count(keys, None, 50, mode, iters=4)
    print("mode %d (0 buckets, 3 segments, 1 classic): %d distinct, %.3f ms per sort+count" % (mode, len(k), ms), flush=True)
    if mode == 0:
        k0, c0 = k, c
    assert np.array_equal(k0, k) and np.array_equal(c0, c)
for mode in (0, 3):
    nat.dbg_profile(True)
    nat.dbg_sort_count(keys, None, 50, mode, iters=1)
    print(mode, {k: round(v[0], 3) for k, v in nat.dbg_profile(False).items()})
w = c0
rk = (k0 * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(14)   # a bijection of distinct keys, like rc()
for mode, nm in ((0, "weighted"), (2, "distinct+payload, buckets"), (4, "distinct+payload, segments"), (1, "classic")):
    k, c, ms = nat.dbg_sort_count(rk, w, 50, mode, iters=4)
    print("pairs %s: %d -> %.3f ms" % (nm, len(k), ms), flush=True)
