"""codec64 encode of a config[1]-sized set on the device, timed per call (for ncu: the enc_* kernels in isolation).
    gpurun -- 'python tools/encode_probe.py; ncu --set full -k regex:enc_ --launch-skip 6 --launch-count 3 -o gpurun_out/enc python tools/encode_probe.py'"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat

n = int(os.environ.get("N", 38800000))
rng = np.random.default_rng(1)
ks = np.unique(rng.integers(0, 2 ** 50, n, dtype=np.uint64))
cs = np.where(rng.random(len(ks)) < 0.73, 1, rng.poisson(24, len(ks))).astype(np.uint32)
cs[cs == 0] = 1
s = nat.KmerSet.from_arrays(ks, cs)
for it in range(int(os.environ.get("ITERS", 4))):
    nat.dbg_profile(True)
    t0 = time.perf_counter()
    w = s.encode_dev()
    dt = (time.perf_counter() - t0) * 1e3
    prof = nat.dbg_profile(False)
    print("encode_dev of %d entries: %.3f ms wall, stage %s, words %s" % (len(ks), dt, {k: round(v[0], 3) for k, v in prof.items()}, w.sizes()), flush=True)
    w.free()
