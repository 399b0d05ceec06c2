import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
n = int(os.environ.get("N", 126000000))
mb = int(os.environ.get("MB", 8))
rng = np.random.default_rng(1)
keys = rng.integers(0, 2 ** 50, n, dtype=np.uint64)
_, _, ms = nat.dbg_sort(keys, None, 50, mb, iters=int(os.environ.get("ITERS", 3)))
print("sort n=%d maxbits=%d: %.3f ms" % (n, mb, ms))
