"""sort+count when many keys are highly repeated (human-like repeats: 'big' segments > 1024 keys): config[1]'s canonical
keys plus FRAC of extra keys drawn from NHOT hot k-mers (copy number in the thousands)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
from tools import synth
g = synth.genome(5000000)
fq = synth.fastq_array(g, 1000000).reshape(-1)
codes, nr = nat.dbg_parse(fq.tobytes(), False)
keys = nat.dbg_extract(25, codes)
rng = np.random.default_rng(3)
for frac, nhot in ((0.0, 0), (0.02, 2000), (0.10, 2000), (0.10, 50), (0.30, 20000)):
    if frac:
        hot = keys[rng.integers(0, len(keys), nhot)]
        extra = hot[rng.integers(0, nhot, int(len(keys) * frac))]
        ks = np.concatenate([keys, extra]); rng.shuffle(ks)
    else:
        ks = keys
    res = {}
    for mode in (0, 1):
        k, c, ms = nat.dbg_sort_count(ks, None, 50, mode, iters=3)
        res[mode] = (k, c, ms)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    nat.dbg_profile(True); nat.dbg_sort_count(ks, None, 50, 0, iters=1); prof = nat.dbg_profile(False)
    print("extra %.0f %% in %d hot keys (n = %d, max count %d): segmented %.2f ms, classic %.2f ms; stages %s" % (
        100 * frac, nhot, len(ks), int(res[0][1].max()), res[0][2], res[1][2], {a: round(b[0], 2) for a, b in prof.items()}), flush=True)
