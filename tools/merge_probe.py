"""zot merge kernel path: N-way union with counts summed (zb_merge) on NSETS device-resident counted sets"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
from tools import synth
nsets = int(os.environ.get("NSETS", 64))
g = synth.genome(5000000)
sets = []
t0 = time.time()
for i in range(nsets):
    h = synth.mutate(g, 0.0005 + 0.0195 * (i % 16) / 16 + 0.02 * (i // 16), 100 + i)   # 4 clades, 0.05 % .. 2 % within
    km = nat.Kmerizer(25); km.feed(synth.fasta_bytes(h), True); s, _ = km.finish(); km.close()
    sets.append(s)
tot = sum(len(s) for s in sets)
print("%d sets, %d (k-mer, count) entries in total, built in %.1f s" % (nsets, tot, time.time() - t0), flush=True)
MODES = os.environ.get("MODES", "tree,sort,auto").split(",")
for mode in MODES:
  os.environ["ZB_MERGE"] = mode
  print("ZB_MERGE=%s" % mode)
  for it in range(3):
    nat.device_sync()
    if it == 2:
        nat.dbg_profile(True)
    t0 = time.perf_counter()
    m = nat.merge(sets)
    nat.device_sync()
    dt = time.perf_counter() - t0
    if it == 2:
        print("   stages (ms):", {k: round(v[0], 3) for k, v in nat.dbg_profile(False).items()})
    print("zb_merge: %.1f ms -> %d distinct; model 12 (sum |Xi| + |U|) B = %.2f GB -> %.0f GB/s" % (
        dt * 1e3, len(m), 12 * (tot + len(m)) / 1e9, 12 * (tot + len(m)) / dt / 1e9), flush=True)
    if it < 2:
        m.free()
  if mode == "tree":
    ref_k, ref_c = m.fetch()
    m.free()
    nat.release_cache()
mk, mc = m.fetch()
if "tree" not in MODES:
    sys.exit(0)
assert np.array_equal(mk, ref_k) and np.array_equal(mc, ref_c), "n-way merge differs from the pairwise tree"
assert np.all(mk[1:] > mk[:-1])
assert int(mc.astype(np.uint64).sum()) == sum(int(s.fetch()[1].astype(np.uint64).sum()) for s in sets[:4]) + sum(
    int(s.stats()["acgt_weighted"][0] + s.stats()["acgt_weighted"][1] + s.stats()["acgt_weighted"][2] + s.stats()["acgt_weighted"][3]) for s in sets[4:])
print("sum of counts preserved")
