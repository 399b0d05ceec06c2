"""
Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tools/mgpu_check.py under torchrun -- reads sharded
over ranks, canonical k-mers exchanged (fused NVLink peer-memory routing and the NCCL all-to-all fallback), the union
of the per-rank counted sets compared bit for bit with the oracle; all-pairs distance tiles sharded and added up.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,k", [("p2p", 25), ("p2p_reserve", 25), ("nccl", 25), ("p2p", 31)])
def test_two_rank_kmerize_and_allpairs(mode, k):
    """k = 31 is BASELINE.json config[4]'s k (a 1/1000-scale instance of it: same code path, 62-bit keys)"""
    from zotmer_b200 import _native
    if _native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ZB_EXCHANGE=mode, ZB_CHECK_K=str(k))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mgpu_check ok" in r.stdout


# ---- one `zot` process, several GPUs (ZB_GPUS, library/devices.py): the file must be the single-GPU file, byte for byte
CLI_KMERIZE = [(5, "kat6.k5", ["kat6.fa"]), (25, "g1.k25", ["g1.fa"]), (32, "g1.k32", ["g1.fa"]), (8, "r1.k8", ["r1.fq"]),
               (25, "r1.k25", ["r1.fq"]), (31, "r1.k31", ["r1.fq"]), (25, "mix.k25", ["s0.fa", "r1.fq", "s1.fa"])]


@pytest.mark.parametrize("ngpu", [2, 3, 4, 8])
def test_cli_kmerize_several_gpus_golden_bytes(ngpu, tmp_path, monkeypatch):
    from zotmer_b200 import _native, cli
    if _native.device_count() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    golden = os.path.join(ROOT, "tests", "golden", "data")
    monkeypatch.chdir(golden)
    monkeypatch.setenv("ZB_GPUS", str(ngpu))
    for k, out, ins in CLI_KMERIZE:
        o = tmp_path / out
        cli.main(["kmerize", str(k), str(o)] + ins)
        with open(os.path.join(golden, out), "rb") as f:
            assert o.read_bytes() == f.read(), (ngpu, out)
    # capture (-C) and sub-sampling (-D) on top of the multi-device ranges
    for args, out in ((["-C", "baits.fa", "25"], "r1_C.k25"), (["-D", "0.3", "-S", "5", "25"], "r1_D03_S5.k25")):
        o = tmp_path / out
        cli.main(["kmerize"] + args + [str(o), "r1.fq"])
        with open(os.path.join(golden, out), "rb") as f:
            assert o.read_bytes() == f.read(), (ngpu, out)


def test_cli_kmerize_two_gpus_equals_one_at_size(tmp_path, monkeypatch):
    """200,000 reads (30 Mbases): several rounds' worth of keys per device, ranges of ~10 M entries, words that span
    the range boundary -- the two files must be identical"""
    import numpy as np
    from zotmer_b200 import _native, cli
    from tools import synth
    if _native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = synth.genome(1500000, seed=9)
    fq = tmp_path / "reads.fq"
    synth.fastq_array(g, 200000, seed=10).tofile(str(fq))
    fa = tmp_path / "genome.fa"
    fa.write_bytes(synth.fasta_bytes(g))
    outs = {}
    for ngpu in (1, 2):
        monkeypatch.setenv("ZB_GPUS", str(ngpu))
        o = tmp_path / ("o%d.k25" % ngpu)
        cli.main(["kmerize", "25", str(o), str(fq), str(fa)])
        outs[ngpu] = o.read_bytes()
    assert outs[1] == outs[2]
    assert len(outs[1]) > 10000000


@pytest.mark.parametrize("ngpu", [2, 4])
def test_cli_kmerize_bgzf_on_several_gpus(ngpu, tmp_path, monkeypatch):
    """block-compressed inputs on several GPUs: every device inflates a run of members, the incomplete record behind a
    run's last record boundary is handed to the device that holds the next run (library/devices.py:_stageBgzfGroup);
    the file is the golden one / the one a single GPU writes from the plain text"""
    from zotmer_b200 import _native, cli
    from zotmer_b200.library import reads
    from tools import synth
    if _native.device_count() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    golden = os.path.join(ROOT, "tests", "golden", "data")
    monkeypatch.setenv("ZB_GPUS", str(ngpu))
    for k, out, src in ((25, "r1.k25", "r1.fq"), (31, "r1.k31", "r1.fq"), (25, "g1.k25", "g1.fa")):
        data = open(os.path.join(golden, src), "rb").read()
        for block in (700, 65280):
            gz = tmp_path / (src + ".gz")
            gz.write_bytes(synth.bgzf_bytes(data, block=block))
            o = tmp_path / out
            cli.main(["kmerize", str(k), str(o), str(gz)])
            with open(os.path.join(golden, out), "rb") as f:
                assert o.read_bytes() == f.read(), (ngpu, out, block)
    # at size: several rounds of groups per device (1 MiB of text per group), FASTQ + FASTA + a plain file in one call
    g = synth.genome(1200000, seed=21)
    fq = synth.fastq_array(g, 60000, seed=22).reshape(-1).tobytes()
    fa = synth.fasta_bytes(g) + b">p2 x\n" + synth.fasta_bytes(synth.genome(300000, seed=23))[6:]
    (tmp_path / "a.fq").write_bytes(fq)
    (tmp_path / "a.fq.gz").write_bytes(synth.bgzf_bytes(fq, block=40000))
    (tmp_path / "b.fa").write_bytes(fa)
    (tmp_path / "b.fa.gz").write_bytes(synth.bgzf_bytes(fa))
    monkeypatch.setattr(reads, "BGZF_GROUP", 1 << 20)
    cli.main(["kmerize", "25", str(tmp_path / "z.k25"), str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fa.gz"), str(tmp_path / "a.fq")])
    monkeypatch.setenv("ZB_GPUS", "1")
    cli.main(["kmerize", "25", str(tmp_path / "p.k25"), str(tmp_path / "a.fq"), str(tmp_path / "b.fa"), str(tmp_path / "a.fq")])
    assert (tmp_path / "z.k25").read_bytes() == (tmp_path / "p.k25").read_bytes()
