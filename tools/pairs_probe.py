"""all-pairs |X n Y| timing: NSETS synthetic bacterial k-mer sets (k=25, ~10 M k-mers each) -> set-pairs/s"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zotmer_b200 import _native as nat
from tools import synth
nsets = int(os.environ.get("NSETS", 32))
glen = int(os.environ.get("GLEN", 5000000))
t0 = time.time()
base = [synth.genome(glen, seed=1000 + c) for c in range(max(1, nsets // 8))]
sets = []
for i in range(nsets):
    g = synth.mutate(base[i % len(base)], 0.001 + 0.009 * (i // len(base)) / max(1, nsets // len(base)), 2000 + i)
    km = nat.Kmerizer(25); km.feed(synth.fasta_bytes(g), True); s, _ = km.finish(); km.close()
    sets.append(s.project(0))
    s.free()
print("%d sets of ~%d k-mers built in %.1f s" % (nsets, len(sets[0]), time.time() - t0), flush=True)
npairs = nsets * (nsets - 1) // 2
for it in range(3):
    nat.dbg_profile(True)
    t0 = time.time(); abc = nat.allpairs_abc(sets); dt = time.time() - t0
    prof = nat.dbg_profile(False)
    print("   stages (ms):", {k: round(v[0], 3) for k, v in prof.items()})
    print("allpairs: %d pairs in %.1f ms wall, kernel %.2f ms -> %.0f set-pairs/s (kernel), eq. %.1f TB/s of 8(|X|+|Y|) B/pair" % (
        npairs, dt * 1e3, prof["allpairs"][0], npairs / prof["allpairs"][0] * 1e3,
        sum(8 * (abc[:, 0] * 2 + abc[:, 1] + abc[:, 2])) / prof["allpairs"][0] / 1e9), flush=True)
I, J = np.triu_indices(nsets, 1)
sub = slice(0, min(npairs, 64))
t0 = time.time(); ref = nat.pairs_abc(sets, I[sub], J[sub]); dt = time.time() - t0
print("pair-at-a-time: %d pairs in %.1f ms -> %.0f set-pairs/s" % (len(ref), dt * 1e3, len(ref) / dt))
assert np.array_equal(ref, abc[sub])
print("jaccard of pair 0: %.6f" % (float(abc[0, 0]) / float(abc[0].sum())))
