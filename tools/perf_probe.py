import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
from tools import synth

nreads = int(os.environ.get("NREADS", 1000000))
t0 = time.time()
g = synth.genome(5000000)
fq = synth.fastq_array(g, nreads).reshape(-1)
print("synth %.1fs  fastq bytes %d" % (time.time() - t0, fq.nbytes), flush=True)

# --- sort alone, 126M canonical-like 50-bit keys
rng = np.random.default_rng(1)
n = 126 * nreads
keys = rng.integers(0, 2 ** 50, n, dtype=np.uint64)
for mb in (8, 9, 10, 11):
    _, _, ms = nat.dbg_sort(keys, None, 50, mb, iters=4)
    passes = -(-50 // mb)
    print("sort n=%d bits=50 maxbits=%d passes=%d: %.3f ms  -> %.1f GB/s algorithmic (16B/key/pass + 8B hist)" % (
        n, mb, passes, ms, n * (16 * passes + 8) / ms / 1e6), flush=True)
del keys

import torch
d = torch.from_numpy(fq).cuda()
torch.cuda.synchronize()
for it in range(3):
    nat.dbg_profile(True)
    t0 = time.time()
    km = nat.Kmerizer(25)
    km.feed_dev(d.data_ptr(), d.numel(), False)
    s, nr = km.finish()
    t1 = time.time()
    tr = s.trim(2)
    t2 = time.time()
    prof = nat.dbg_profile(False)
    print("iter %d: kmerize(dev) %.1f ms, trim %.1f ms, distinct=%d trimmed=%d reads=%d" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(s), len(tr), nr))
    print("   stages:", {k: round(v[0], 3) for k, v in prof.items()}, "sum=%.3f" % sum(v[0] for v in prof.values()), flush=True)
    km.close(); s.free(); tr.free()
for it in range(2):
    nat.dbg_profile(True)
    t0 = time.time()
    km = nat.Kmerizer(25)
    km.feed(fq, False)
    s, nr = km.finish()
    t1 = time.time()
    k_, c_ = s.fetch()
    t2 = time.time()
    prof = nat.dbg_profile(False)
    print("host iter %d: feed+finish %.1f ms, fetch %.1f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    print("   stages:", {k: round(v[0], 3) for k, v in prof.items()}, flush=True)
    km.close(); s.free()
