"""BGZF input: the device inflater against the host paths, on config[1]'s FASTQ (315 MB of text).
    gpurun -- 'python tools/inflate_probe.py > gpurun_out/inflate_probe.log 2>&1'
Prints: compressed size, host inflate (zlib on one core; BGZF members on the thread pool), stage_bgzf wall (compressed H2D +
inflate kernel + sync) and the kernel alone (CUDA events), then `zot kmerize` from the .gz against the plain file."""
import os, sys, time, tempfile, zlib
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from zotmer_b200 import cli, _native as nat
from zotmer_b200.library.file import gunzipBytes

reads = int(os.environ.get("ZB_PROBE_READS", "1000000"))
level = int(os.environ.get("ZB_PROBE_LEVEL", "6"))
tmp = tempfile.mkdtemp(prefix="zb_inf_", dir=os.environ.get("ZB_TMP"))
g = synth.genome(5000000)
text = synth.fastq_array(g, reads).reshape(-1).tobytes()
t0 = time.perf_counter()
step = 65280 * 64
with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
    parts = list(ex.map(lambda o: synth.bgzf_bytes(text[o:o + step], level=level, eof=False), range(0, len(text), step)))
z = b"".join(parts) + synth.bgzf_bytes(b"")
print("text %d bytes -> BGZF %d bytes (level %d, %.1f s to compress on %d threads)" % (len(text), len(z), level, time.perf_counter() - t0, os.cpu_count()))
print("probe:", nat.bgzf_probe(z))
bases = reads * 150

def rate(ms):
    return "%.1f ms = %.2f GB/s of text = %.2f Gbases/s" % (ms, len(text) / ms / 1e6, bases / ms / 1e6)

t0 = time.perf_counter(); d = zlib.decompressobj(31); one = d.decompress(z[:len(z) // 8]); t1 = time.perf_counter()
print("host zlib, one core (first eighth, first member only):", len(one), "bytes")
t0 = time.perf_counter(); out = gunzipBytes(z); t1 = time.perf_counter()
assert out == text
print("host: BGZF members on the Python thread pool (%d cores): %s" % (os.cpu_count(), rate((t1 - t0) * 1e3)))
del out
for it in range(5):
    nat.dbg_profile(True)
    t0 = time.perf_counter()
    st, used = nat.stage_bgzf(z, 0)
    t1 = time.perf_counter()
    prof = nat.dbg_profile(False)
    print("device: stage_bgzf wall %s; kernel alone %s" % (rate((t1 - t0) * 1e3), rate(prof["inflate"][0])), flush=True)
    if it == 0:
        assert st.fetch() == text
        print("device text == original")
    st.free()
fq, gz = os.path.join(tmp, "reads.fq"), os.path.join(tmp, "reads.fq.gz")
open(fq, "wb").write(text)
open(gz, "wb").write(z)
for (name, src) in (("plain", fq), ("bgzf", gz), ("plain", fq), ("bgzf", gz)):
    ts = []
    for it in range(4):
        o = os.path.join(tmp, "o.k25")
        t0 = time.perf_counter()
        cli.main(["kmerize", "25", o, src])
        ts.append((time.perf_counter() - t0) * 1e3)
    print("zot kmerize 25 from %s: %s ms -> best %s" % (name, [round(t, 1) for t in ts], rate(min(ts))))
a = open(os.path.join(tmp, "o.k25"), "rb").read()
cli.main(["kmerize", "25", os.path.join(tmp, "p.k25"), fq])
assert a == open(os.path.join(tmp, "p.k25"), "rb").read()
print("files identical")
