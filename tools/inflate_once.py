"""A few stage_bgzf calls on a FASTQ of ZB_PROBE_READS reads (ncu target):
    ncu --set full --import-source on -k regex:bgzf_inflate --launch-skip 1 -c 1 -o gpurun_out/inflate python tools/inflate_once.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth
from zotmer_b200 import _native as nat
reads = int(os.environ.get("ZB_PROBE_READS", "300000"))
level = int(os.environ.get("ZB_PROBE_LEVEL", "6"))
g = synth.genome(5000000)
text = synth.fastq_array(g, reads).reshape(-1).tobytes()
z = synth.bgzf_bytes(text, level=level)
for it in range(3):
    nat.dbg_profile(True)
    st, used = nat.stage_bgzf(z, 0)
    ms = nat.dbg_profile(False)["inflate"][0]
    print("level %d: %d -> %d bytes, inflate kernel %.2f ms = %.1f GB/s of text" % (level, len(z), len(text), ms, len(text) / ms / 1e6))
    st.free()
