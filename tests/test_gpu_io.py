"""
GPU tests of the round-2 entry points (pytest -m gpu): device-resident word streams (zb_set_encode_dev, zb_words_*),
the range-partitioned encoder (zb_set_encode_plan / _emit: several ranges of one sorted set -> one stream, word for word),
the host I/O runtime (zb_stage_input / zb_stage_fd / zb_kmerize_feed_staged, zb_host_count_byte, zb_words_write_fd) and
the guard-band allocator that stands in for compute-sanitizer (closed on the GPU pool, profiles/r02_sanitizer.md).
Everything is compared bit for bit with the oracle (oracle/) or with the single-call path.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nat():
    from zotmer_b200 import _native
    assert _native.device_count() >= 1, "no CUDA device"
    return _native


def random_set(rng, n, bits=50, maxcount=70000):
    ks = np.unique(rng.integers(0, 2 ** bits, n, dtype=np.uint64))
    # counts with a long tail: most fit 10 bits (six per word), a few need a word of their own
    cs = rng.integers(1, 60, len(ks), dtype=np.uint32)
    big = rng.random(len(ks)) < 0.01
    cs[big] = rng.integers(1, maxcount, int(big.sum()), dtype=np.uint32)
    return ks, cs


@pytest.mark.parametrize("n", [1, 6, 2047, 2048, 2049, 100000, 1300001])
def test_words_dev_equal_oracle_and_file(nat, n, tmp_path):
    rng = np.random.default_rng(n)
    ks, cs = random_set(rng, n)
    s = nat.KmerSet.from_arrays(ks, cs)
    w = s.encode_dev()
    ek, ec = co.encode(ks, True), co.encode(cs.astype(np.uint64), False)
    assert w.sizes() == (len(ek), len(ec))
    kw, cw = w.fetch()
    assert np.array_equal(kw, ek) and np.array_equal(cw, ec)
    # pinned destinations
    pk, pc = nat.PinnedArray(len(ek), np.uint64), nat.PinnedArray(len(ec), np.uint64)
    kw2, cw2 = w.fetch(pk.a, pc.a)
    assert np.array_equal(kw2, ek) and np.array_equal(cw2, ec)
    # the I/O threads write both streams into a file behind a header
    fn = str(tmp_path / "w.bin")
    with open(fn, "wb") as f:
        f.write(b"HEAD" * 4)
        f.flush()
        w.write_fd(f.fileno(), 16, 16 + 8 * len(ek))
    blob = open(fn, "rb").read()
    assert blob[:16] == b"HEAD" * 4
    assert blob[16:] == ek.astype("<u8").tobytes() + ec.astype("<u8").tobytes()
    pk.free(); pc.free(); w.free(); s.free()


@pytest.mark.parametrize("seed", range(6))
def test_range_partitioned_encode(nat, seed):
    """ranges of one sorted set, encoded separately with halos and chained entry states == the whole-set streams"""
    rng = np.random.default_rng(100 + seed)
    n = [40, 5000, 70000, 300000, 13, 2048 * 3][seed]
    ks, cs = random_set(rng, n, bits=[50, 50, 62, 50, 20, 50][seed])
    n = len(ks)
    nr = [4, 3, 8, 2, 7, 5][seed]
    cuts = np.sort(rng.integers(0, n + 1, nr - 1))
    if seed == 4:
        cuts = np.array([0, 1, 1, 3, 4, 4])[:nr - 1]      # empty and one-entry ranges: a word spans several of them
    bounds = [0] + [int(c) for c in cuts] + [n]
    ek, ec = co.encode(ks, True), co.encode(cs.astype(np.uint64), False)
    sets, plans, kmaps, cmaps = [], [], [], []
    for r in range(nr):
        a, b = bounds[r], bounds[r + 1]
        s = nat.KmerSet.from_arrays(ks[a:b], cs[a:b])
        prev = int(ks[a - 1]) if a > 0 else 0
        p, km, cm = s.encode_plan(prev, ks[b:b + 5], cs[b:b + 5])
        sets.append(s); plans.append(p); kmaps.append(km); cmaps.append(cm)
    kentry, koff, ktot = nat.chain_ranges(kmaps)
    centry, coff, ctot = nat.chain_ranges(cmaps)
    assert ktot == len(ek) and ctot == len(ec)
    kws, cws = [], []
    for r in range(nr):
        w = plans[r].emit(kentry[r], centry[r])
        kw, cw = w.fetch()
        assert len(kw) == kmaps[r][1][kentry[r]] and len(cw) == cmaps[r][1][centry[r]]
        kws.append(kw); cws.append(cw)
        w.free()
    assert np.array_equal(np.concatenate(kws), ek)
    assert np.array_equal(np.concatenate(cws), ec)
    for s in sets:
        s.free()


def test_range_encode_overflow_is_index_error(nat):
    # a gap of more than 60 bits at a range boundary is still the reference's IndexError (codec64.py:93-99)
    ks = np.array([5, 7, (1 << 62) + 9], np.uint64)
    s1 = nat.KmerSet.from_arrays(ks[:2], np.ones(2, np.uint32))
    s2 = nat.KmerSet.from_arrays(ks[2:], np.ones(1, np.uint32))
    p1, _, _ = s1.encode_plan(0, ks[2:], np.ones(1, np.uint32))
    p2, _, _ = s2.encode_plan(int(ks[1]), [], [])
    p1.emit(0, 0).free()
    with pytest.raises(IndexError):
        p2.emit(0, 0)


def _fastq(rng, nreads, L=100):
    g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 200000)]
    out = []
    for i in range(nreads):
        p = int(rng.integers(0, len(g) - L))
        out.append(b"@r%d\n%s\n+\n%s\n" % (i, g[p:p + L].tobytes(), b"I" * L))
    return b"".join(out)


@pytest.mark.parametrize("nreads", [1, 300, 60000])
def test_staged_feed_equals_feed(nat, nreads, tmp_path):
    rng = np.random.default_rng(nreads)
    fq = _fastq(rng, nreads)
    ek, ec, _, enr = co.kmerize(25, [(fq, False)])
    # from memory (two pieces in flight at once), and from a file descriptor
    half = fq.rfind(b"\n@", 0, len(fq) // 2) + 1 if nreads > 1 else len(fq)
    pieces = [fq[:half], fq[half:]] if half and half < len(fq) else [fq]
    km = nat.Kmerizer(25, 0)
    staged = [nat.stage_input(p) for p in pieces]
    for st in staged:
        km.feed_staged(st, False)
    s, nr = km.finish()
    km.close()
    ks, cs = s.fetch()
    assert nr == enr and np.array_equal(ks, ek) and np.array_equal(cs, ec)
    s.free()
    fn = str(tmp_path / "r.fq")
    open(fn, "wb").write(b"#" * 37 + fq)
    with open(fn, "rb") as f:
        km = nat.Kmerizer(25, 0)
        km.feed_staged(nat.stage_fd(f.fileno(), 37, len(fq)), False)
        s, nr = km.finish()
        km.close()
    ks, cs = s.fetch()
    assert nr == enr and np.array_equal(ks, ek) and np.array_equal(cs, ec)
    s.free()
    # a staged piece that is never fed can be dropped
    nat.stage_input(fq).free()


def test_host_count_byte(nat):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, 9000001, dtype=np.uint8)
    assert nat.host_count_byte(a, 10) == int(np.count_nonzero(a == 10))
    assert nat.host_count_byte(a[:5], 10) == int(np.count_nonzero(a[:5] == 10))
    assert nat.host_count_byte(b"", 10) == 0


GUARD_SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np
import __graft_entry__ as g
from zotmer_b200 import _native as nat
g.smoke()
rng = np.random.default_rng(9)
# adversarial sort + count shapes: one value repeated, skewed prefixes, tiny and odd sizes
for n in (1, 33, 4097, 200001, 1500003):
    keys = rng.integers(0, 2 ** 50, n, dtype=np.uint64)
    keys[: n // 3] = keys[0]
    keys[n // 3: n // 2] &= np.uint64(0xffff)
    ok, oc, _ = nat.dbg_sort_count(keys, None, 50)
    u, c = np.unique(keys, return_counts=True)
    assert np.array_equal(ok, u) and np.array_equal(oc, c.astype(np.uint32))
sets = [nat.KmerSet.from_arrays(np.unique(rng.integers(0, 2 ** 50, 50000, dtype=np.uint64))) for _ in range(5)]
nat.merge(sets).free()
nat.allpairs_abc(sets)
w = sets[0].encode_dev(); w.fetch(); w.free()
nb, bad = nat.guard_check(0)
print("guard: %%d blocks scanned, %%d damaged" %% (nb, bad))
assert nb > 10 and bad == 0
"""


def test_guard_bands_clean():
    """the whole smoke pipeline + adversarial shapes with 256-byte pattern bands around every device block: no kernel
    stores outside its buffers (the stand-in for compute-sanitizer memcheck on this pool)"""
    env = dict(os.environ, ZB_GUARD="1")
    r = subprocess.run([sys.executable, "-c", GUARD_SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "0 damaged" in r.stdout


def test_guard_bands_detect_damage(nat):
    """the check itself works: without ZB_GUARD the entry point refuses, so a clean report is never vacuous"""
    if os.environ.get("ZB_GUARD"):
        pytest.skip("guards are on in this process")
    with pytest.raises(nat.NativeError):
        nat.guard_check(0)


def test_repeat_determinism(nat):
    """a shared-memory race in the chained-scan sort, the bucket hash tables or the mirror merge would show up as a run
    that differs: 20 runs of kmerize+count on the same reads, each compared with the oracle"""
    rng = np.random.default_rng(77)
    fq = _fastq(rng, 40000)
    ek, ec, _, _ = co.kmerize(25, [(fq, False)])
    for _ in range(20):
        km = nat.Kmerizer(25, 0)
        km.feed(fq, False)
        s, _ = km.finish()
        km.close()
        ks, cs = s.fetch()
        assert np.array_equal(ks, ek) and np.array_equal(cs, ec)
        s.free()


# ------------------------------------------------------------------------------------------------------------------
# block-compressed input: BGZF members inflated on the device (zb_stage_bgzf, csrc/inflate.cu), record-aligned cuts on
# the device (zb_staged_cut), `zot kmerize reads.fq.gz`
# ------------------------------------------------------------------------------------------------------------------
def _texts():
    from tools import synth
    rng = np.random.default_rng(77)
    g = synth.genome(60000, seed=3)
    fq = synth.fastq_array(g, 3000, seed=4).reshape(-1).tobytes()
    fa = synth.fasta_bytes(g)
    rnd = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()            # incompressible: stored / near-stored blocks
    two = rng.choice(np.frombuffer(b"ab", dtype=np.uint8), 200001).tobytes()  # short codes, long matches
    runs = b"".join(bytes([65 + i % 7]) * int(n) for i, n in enumerate(rng.integers(1, 900, 700)))   # distance < length
    return {"fastq": fq, "fasta": fa, "random": rnd, "two": two, "runs": runs, "tiny": b"ACGT\n", "empty": b""}


@pytest.mark.parametrize("name", ["fastq", "fasta", "random", "two", "runs", "tiny", "empty"])
def test_bgzf_inflate_on_device_equals_zlib(nat, name):
    import zlib
    from tools import synth
    data = _texts()[name]
    for (level, block, strategy) in [(6, 65280, 0), (1, 65280, 0), (9, 30000, 0), (0, 65000, 0), (6, 4097, zlib.Z_FIXED),
                                     (6, 65280, zlib.Z_HUFFMAN_ONLY), (4, 777, 0)]:
        if len(data) > 300000 and block < 4000:
            continue
        z = synth.bgzf_bytes(data, level=level, block=block, strategy=strategy)
        probe = nat.bgzf_probe(z)
        assert probe is not None and probe[1] == len(data)
        st, used = nat.stage_bgzf(z, 0)
        assert used == len(z) and len(st) == len(data)
        assert st.fetch() == data, (name, level, block, strategy)
        st.free()


def test_bgzf_groups_and_carry(nat):
    """a file inflated group by group: every group's text follows the carried tail of the previous piece"""
    from tools import synth
    data = _texts()["fastq"]
    z = synth.bgzf_bytes(data, block=20000)
    za = np.frombuffer(z, dtype=np.uint8)
    off, got, prev, prev_cut = 0, b"", None, 0
    while off < len(z):
        carried = (len(prev) - prev_cut) if prev is not None else 0
        st, used = nat.stage_bgzf(za[off:], 0, carried + 70000, prev, prev_cut)
        assert 0 < used
        off += used
        text = st.fetch()
        if prev is not None:
            assert text[:carried] == got[len(got) - carried:]
            got += text[carried:]
            prev.free()
        else:
            got = text
        prev, prev_cut = st, (len(st) * 2) // 3
    prev.free()
    assert got == data


def test_bgzf_corrupt_member_is_an_error(nat):
    from tools import synth
    data = _texts()["fastq"]
    z = bytearray(synth.bgzf_bytes(data))
    z[5000] ^= 0x55
    z[5001] ^= 0xAA
    z[5002] ^= 0x0F
    st = None
    try:
        st, _ = nat.stage_bgzf(bytes(z), 0)
        text = st.fetch()
    except Exception as e:
        assert "inflate" in str(e) or "BGZF" in str(e)
    else:
        # (a flipped literal inflates to the right size: only the CRC, which `gunzip -c | reader` never waits for, differs)
        assert len(text) == len(data)
    if st is not None:
        st.free()
    # the library still works afterwards
    st, _ = nat.stage_bgzf(synth.bgzf_bytes(b"ACGT\n" * 10), 0)
    assert st.fetch() == b"ACGT\n" * 10
    st.free()


@pytest.mark.parametrize("kind", ["fastq", "fastq_no_final_newline", "fastq_partial", "fasta", "fasta_one_record", "short"])
def test_staged_cut_equals_host_pieces(nat, kind):
    from zotmer_b200.library.reads import pieces
    t = _texts()
    data = {"fastq": t["fastq"], "fastq_no_final_newline": t["fastq"][:-1], "fastq_partial": t["fastq"][:len(t["fastq"]) - 100],
            "fasta": t["fasta"] + b">second\nACGT\nAC\n>third x\nGGGG", "fasta_one_record": t["fasta"], "short": b"@r\nAC\n"}[kind]
    fa = kind.startswith("fasta")
    st = nat.stage_input(data, 0)
    cut = st.cut(fa)
    st.free()
    if fa:
        want = data.rfind(b"\n>") + 1
    else:
        nl = data.count(b"\n")
        want = 0
        if nl >= 4:
            at = len(data)
            for _ in range(nl % 4 + 1):
                at = data.rfind(b"\n", 0, at)
            want = at + 1
    assert cut == want
    if want:
        # the same boundary the host splitter picks when it has to cut inside this text
        ps = list(pieces(data, fa, max_piece=len(data) - 1))
        assert len(ps) >= 2 and len(bytes(ps[0])) <= want


@pytest.mark.parametrize("k,group", [(25, None), (25, 200000), (31, 70000)])
def test_kmerize_bgzf_file_equals_plain(nat, tmp_path, k, group, monkeypatch):
    """`zot kmerize` of a bgzip'd FASTQ / FASTA (inflated on the device, in groups with carried tails) gives the set of
    the plain file, which the other tests pin to the oracle"""
    from tools import synth
    from zotmer_b200.commands.kmerize import kmerizeFiles
    from zotmer_b200.library import reads
    if group:
        monkeypatch.setattr(reads, "BGZF_GROUP", group)
    t = _texts()
    for (name, ext) in (("fastq", ".fq"), ("fasta", ".fa")):
        data = t[name] if name == "fastq" else t["fasta"] + b">p2 plasmid\n" + t["fasta"][7:5000] + b"\n>p3\nACGTTGCA\n"
        plain, gz = str(tmp_path / ("x" + ext)), str(tmp_path / ("x" + ext + ".gz"))
        open(plain, "wb").write(data)
        open(gz, "wb").write(synth.bgzf_bytes(data, block=30000))
        before = nat.launch_count(0)
        (a, na) = kmerizeFiles(k, [gz], 0)
        assert nat.launch_count(0) > before
        (b, nb) = kmerizeFiles(k, [plain], 0)
        ak, ac = a.fetch()
        bk, bc = b.fetch()
        assert na == nb and np.array_equal(ak, bk) and np.array_equal(ac, bc)
        ek, ec, _, enr = co.kmerize(k, [(data, name == "fasta")])
        assert enr == na and np.array_equal(ak, ek) and np.array_equal(ac, ec)
        a.free(); b.free()


def test_cli_kmerize_bgzf_golden(tmp_path):
    """the golden k-mer set of r1.fq, from the same reads bgzip'd"""
    from tools import synth
    gold = os.path.join(ROOT, "tests", "golden", "data")
    src = os.path.join(gold, "r1.fq")
    want = os.path.join(gold, "r1.k25")
    if not (os.path.exists(src) and os.path.exists(want)):
        pytest.skip("golden r1.fq / r1.k25 not present")
    gz = str(tmp_path / "r1.fq.gz")
    open(gz, "wb").write(synth.bgzf_bytes(open(src, "rb").read(), block=5000))
    out = str(tmp_path / "o.k25")
    subprocess.check_call([sys.executable, "-m", "zotmer_b200.cli", "kmerize", "25", out, gz], cwd=ROOT)
    assert open(out, "rb").read() == open(want, "rb").read()


def test_stage_input_from_pinned_memory(nat):
    """a source that is pinned already is copied from where it lies (no trip through the pinned ring): same bytes, same set"""
    data = _texts()["fastq"]
    pin = nat.PinnedArray(len(data), np.uint8)
    pin.a[:] = np.frombuffer(data, dtype=np.uint8)
    st = nat.stage_input(pin.a, 0)
    assert st.fetch() == data
    st.free()
    km = nat.Kmerizer(25, 0)
    km.feed_staged(nat.stage_input(pin.a, 0), False)
    a, na = km.finish()
    km.close()
    km = nat.Kmerizer(25, 0)
    km.feed_staged(nat.stage_input(data, 0), False)
    b, nb = km.finish()
    km.close()
    ak, ac = a.fetch()
    bk, bc = b.fetch()
    assert na == nb and np.array_equal(ak, bk) and np.array_equal(ac, bc)
    z = __import__("tools.synth", fromlist=["x"]).bgzf_bytes(data)
    pz = nat.PinnedArray(len(z), np.uint8)
    pz.a[:] = np.frombuffer(z, dtype=np.uint8)
    st, used = nat.stage_bgzf(pz.a, 0)
    assert used == len(z) and st.fetch() == data
    st.free(); pin.free(); pz.free()


@pytest.mark.parametrize("k", [25, 31])
def test_kmerize_fasta_record_longer_than_a_piece(nat, k):
    """a chromosome that does not fit a feed piece is cut inside (library/reads.py:splitPieces): same set, same record count"""
    from tools import synth
    from zotmer_b200.library.reads import stagedPieces
    g = synth.genome(400000, seed=41)
    fa = synth.fasta_bytes(g) + b">p2\n" + synth.fasta_bytes(synth.genome(90000, seed=42), width=100000)[6:] + b">p3\nACGTTGCA\n"
    km = nat.Kmerizer(k, 0)
    fake = pieces_fed = 0
    for (st, is_fa) in stagedPieces([("x.fa", fa)], 0, max_piece=70000, k=k):
        fake += getattr(st, "fake_records", 0)
        pieces_fed += 1
        km.feed_staged(st, is_fa)
    s, nr = km.finish()
    km.close()
    assert pieces_fed >= 6 and fake >= 4
    ks, cs = s.fetch()
    ek, ec, _, enr = co.kmerize(k, [(fa, True)])
    assert nr - fake == enr == 3 and np.array_equal(ks, ek) and np.array_equal(cs, ec)


def test_kmerize_bgzf_with_a_damaged_member_stops_where_gunzip_stops(nat, tmp_path, monkeypatch):
    """file.py:93-97 pipes `gunzip -c` and never looks at its exit status: what was written before the damaged member
    is processed, the rest is not.  Same here: the group that fails on the device is inflated on the host up to the damage."""
    from tools import synth
    from zotmer_b200.commands.kmerize import kmerizeFiles
    from zotmer_b200.library import reads
    from zotmer_b200.library.file import gunzipBytes
    data = _texts()["fastq"]
    z = bytearray(synth.bgzf_bytes(data, block=30000))
    members = []
    p = 0
    while p < len(z):
        bsize = (z[p + 16] | (z[p + 17] << 8)) + 1
        members.append((p, bsize))
        p += bsize
    assert len(members) > 12
    for victim, group in ((1, None), (9, 70000), (len(members) - 2, 200000)):
        zz = bytearray(z)
        (mp, bs) = members[victim]
        zz[mp + 40:mp + 90] = b"\xff" * 50          # inside the member's deflate stream
        want = gunzipBytes(bytes(zz))
        assert 0 < len(want) < len(data) and data.startswith(want)
        gz = tmp_path / ("bad%d.fq.gz" % victim)
        gz.write_bytes(bytes(zz))
        monkeypatch.setattr(reads, "BGZF_GROUP", group or reads.BGZF_GROUP)
        (a, na) = kmerizeFiles(25, [str(gz)], 0)
        ak, ac = a.fetch()
        ek, ec, _, enr = co.kmerize(25, [(want, False)])
        assert na == enr and np.array_equal(ak, ek) and np.array_equal(ac, ec), victim
        a.free()
