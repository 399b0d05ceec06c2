"""The drop-in commands on BASELINE.json's configs, timed as a user runs them (files in, files out):
config[0]: zot kmerize 25 on a 5 Mbp FASTA, then zot hist;  config[1]: zot kmerize 25 on 1,000,000 x 150 bp FASTQ reads,
then zot trim -c 2;  config[2] (bounded): zot merge of NMERGE k-mer sets;  zot dist / jaccard on a few sets.
Wall-clock per command (process already warm: library loaded, CUDA context up), page-cached files in /tmp."""
import sys, os, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from zotmer_b200 import cli, _native as nat

tmp = os.environ.get("ZB_TMP", "/tmp/zb_cli")
os.makedirs(tmp, exist_ok=True)
nmerge = int(os.environ.get("NMERGE", 16))


def run(argv, quiet=True):
    t0 = time.perf_counter()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf if quiet else sys.stdout):
        cli.main(argv)
    nat.device_sync(0)
    return (time.perf_counter() - t0) * 1e3, buf.getvalue()


g = synth.genome(5000000)
fa = os.path.join(tmp, "genome_5M.fa")
open(fa, "wb").write(synth.fasta_bytes(g))
fq = os.path.join(tmp, "reads_5M_30x.fq")
synth.fastq_array(g, 1000000).tofile(fq)
print("inputs: %s %.1f MB, %s %.1f MB" % (fa, os.path.getsize(fa) / 1e6, fq, os.path.getsize(fq) / 1e6), flush=True)
run(["kmerize", "25", os.path.join(tmp, "warm.k25"), fa])           # warm-up: library, CUDA context, allocator
for rep in range(2):
    ms, _ = run(["kmerize", "25", os.path.join(tmp, "g.k25"), fa])
    print("config[0] zot kmerize 25 g.k25 genome_5M.fa: %.1f ms (%.2f Gbases/s), output %.1f MB" % (
        ms, 5.0e6 / ms / 1e6, os.path.getsize(os.path.join(tmp, "g.k25")) / 1e6), flush=True)
ms, out = run(["hist", os.path.join(tmp, "g.k25")])
print("config[0] zot hist g.k25: %.1f ms, %d lines" % (ms, len(out.splitlines())), flush=True)
for rep in range(2):
    ms, _ = run(["kmerize", "25", os.path.join(tmp, "r.k25"), fq])
    print("config[1] zot kmerize 25 r.k25 reads_5M_30x.fq: %.1f ms (%.2f Gbases/s), output %.1f MB" % (
        ms, 150.0e6 / ms / 1e6, os.path.getsize(os.path.join(tmp, "r.k25")) / 1e6), flush=True)
    ms, _ = run(["trim", "-c", "2", os.path.join(tmp, "r2.k25"), os.path.join(tmp, "r.k25")])
    print("config[1] zot trim -c 2 r2.k25 r.k25: %.1f ms, output %.1f MB" % (ms, os.path.getsize(os.path.join(tmp, "r2.k25")) / 1e6), flush=True)
names = []
for i in range(nmerge):
    f = os.path.join(tmp, "m%02d.fa" % i)
    open(f, "wb").write(synth.fasta_bytes(synth.mutate(g, 0.0005 + 0.02 * i / nmerge, 100 + i)))
    o = os.path.join(tmp, "m%02d.k25" % i)
    run(["kmerize", "25", o, f])
    names.append(o)
for it in range(2):     # the first call also pays the first launch of the merge kernels and the growth of the allocator
    nat.dbg_profile(True)
    ms, _ = run(["merge", os.path.join(tmp, "merged.k25")] + names)
    prof = nat.dbg_profile(False)
    print("config[2] bounded: zot merge of %d sets (~9.9 M k-mers each): %.1f ms, output %.1f MB; device stages (ms): %s" % (
        nmerge, ms, os.path.getsize(os.path.join(tmp, "merged.k25")) / 1e6, {k: round(v[0], 1) for k, v in prof.items()}), flush=True)
ms, out = run(["dist", "-M", "jaccard.qual", "-M", "kulczynski.qual", "25"] + names)
print("config[3] bounded: zot dist -M jaccard.qual -M kulczynski.qual 25 on %d sets (%d pairs): %.1f ms" % (
    nmerge, nmerge * (nmerge - 1) // 2, ms))
print("\n".join(out.splitlines()[:3]))
ms, out = run(["jaccard", "-a"] + names)
print("zot jaccard -a on %d sets: %.1f ms" % (nmerge, ms))
prof_on = nat.dbg_profile(True)
run(["kmerize", "25", os.path.join(tmp, "r.k25"), fq])
print("stages of one config[1] kmerize:", {k: round(v[0], 2) for k, v in nat.dbg_profile(False).items()})
