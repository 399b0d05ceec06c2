"""Where the time of one `zot kmerize` call goes (ZB_CLI_TRACE phases), config[1]'s FASTQ, a warm process.
    gpurun -- 'python tools/cli_phases.py > gpurun_out/cli_phases.log 2>&1'"""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZB_CLI_TRACE"] = "1"
import numpy as np
from tools import synth
from zotmer_b200 import cli, _native as nat

tmp = tempfile.mkdtemp(prefix="zb_cli_", dir=os.environ.get("ZB_TMP"))
g = synth.genome(5000000)
fq = os.path.join(tmp, "reads.fq")
synth.fastq_array(g, 1000000).tofile(fq)
for it in range(6):
    out = os.path.join(tmp, "o%d.k25" % it)
    t0 = time.perf_counter()
    cli.main(["kmerize", "25", out, fq])
    print("run %d: %.1f ms" % (it, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
    if it == 3:
        nat.dbg_profile(True)
    if it == 4:
        print("device stages:", {k: round(v[0], 2) for k, v in nat.dbg_profile(False).items()}, file=sys.stderr)
# raw rates of the pieces: pwrite of 203 MB from the pinned ring, pread of 315 MB
import ctypes
buf = np.zeros(203 << 20, np.uint8)
for nm in ("w1", "w2"):
    t0 = time.perf_counter()
    with open(os.path.join(tmp, nm), "wb") as f:
        f.write(buf)
    print("python f.write of 203 MiB: %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
# the same bytes through a shared mapping of the output file, filled by several threads (page faults scale across
# threads; buffered write() holds the inode lock, so parallel pwrite() calls on one file run one after the other)
import mmap
from concurrent.futures import ThreadPoolExecutor
for nthreads in (1, 4, 8):
    fn = os.path.join(tmp, "m%d" % nthreads)
    t0 = time.perf_counter()
    with open(fn, "w+b") as f:
        os.ftruncate(f.fileno(), buf.nbytes)
        mm = mmap.mmap(f.fileno(), buf.nbytes, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
        dst = np.frombuffer(mm, dtype=np.uint8)
        step = 4 << 20
        def cp(o):
            dst[o:o + step] = buf[o:o + step]
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(cp, range(0, buf.nbytes, step)))
        del dst
        mm.close()
    print("mmap + %d threads, 203 MiB: %.1f ms" % (nthreads, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
for nthreads in (4, 8):
    fn = os.path.join(tmp, "p%d" % nthreads)
    t0 = time.perf_counter()
    with open(fn, "wb") as f:
        step = 4 << 20
        def pw(o):
            os.pwrite(f.fileno(), memoryview(buf)[o:o + step], o)
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(pw, range(0, buf.nbytes, step)))
    print("pwrite from %d threads, 203 MiB: %.1f ms" % (nthreads, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
t0 = time.perf_counter()
with open(fq, "rb") as f:
    d = f.read()
print("python f.read of 315 MB: %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
print("tmp is", tmp, os.popen("df -h %s | tail -1" % tmp).read(), file=sys.stderr)
shutil.rmtree(tmp, ignore_errors=True)
