import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zotmer_b200 import _native as nat
from tools import synth
g = synth.genome(5000000)
fq = synth.fastq_array(g, int(os.environ.get("NREADS", 1000000))).reshape(-1)
d = torch.from_numpy(fq).cuda()
for it in range(int(os.environ.get("ITERS", 2))):
    km = nat.Kmerizer(25); km.feed_dev(d.data_ptr(), d.numel(), False); s, nr = km.finish(); t = s.trim(2)
    print(len(s), len(t)); km.close(); s.free(); t.free()
