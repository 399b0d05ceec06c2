"""
CPU-only tests of the host layer: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), the docopt-compatible parser, the casket container against
the reference-written golden files, the metadata-only commands, and the no-CPU-fallback rule.
"""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def g(name):
    return os.path.join(GOLDEN, name)


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zotmer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from zotmer_b200 import _native
    L = _native.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), s
        assert s in _native.SIGNATURES, "binding missing for " + s
    assert set(_native.SIGNATURES) == set(syms)
    assert L.zb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    from zotmer_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError) as ei:
        _native.Kmerizer(25)
    assert ei.value.code == _native.ZB_E_NOGPU
    with pytest.raises(_native.NativeError):
        _native.KmerSet.from_arrays(np.array([1, 2, 3], np.uint64))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "zotmer_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libzot_oracle|zot_oracle", txt, re.M):
                    bad.append(fn)
    assert bad == []


def test_docopt_mini_grammars():
    from zotmer_b200 import docopt_mini as d
    from zotmer_b200 import cli
    from zotmer_b200.commands import kmerize, dist, jaccard, trim, merge, hist
    o = d.docopt(kmerize.__doc__, ["kmerize", "-m", "5", "-v", "25", "out.k25", "a.fa", "b.fq"])
    assert o["-m"] == "5" and o["-v"] is True and o["-D"] is None and o["<k>"] == "25"
    assert o["<output>"] == "out.k25" and o["<input>"] == ["a.fa", "b.fq"]
    o = d.docopt(dist.__doc__, ["dist", "-M", "jaccard.qual", "-M", "*.qual", "25", "a", "b"])
    assert o["-M"] == ["jaccard.qual", "*.qual"] and o["<k>"] == "25" and o["<input>"] == ["a", "b"]
    assert d.docopt(dist.__doc__, ["dist", "25", "a"])["-M"] == []
    o = d.docopt(jaccard.__doc__, ["jaccard", "-ap", "0.9", "a", "b"])
    assert o["-a"] is True and o["-b"] is False and o["-p"] == "0.9"
    o = d.docopt(trim.__doc__, ["trim", "o", "i"])
    assert o["-c"] == "0" and o["-C"] == "0" and o["<input>"] == "i"
    assert d.docopt(trim.__doc__, ["trim", "-c", "2", "o", "i"])["-c"] == "2"
    assert d.docopt(merge.__doc__, ["merge", "o", "i"])["<input>"] == ["i"]
    assert d.docopt(hist.__doc__, ["hist", "x", "y"])["<input>"] == ["x", "y"]
    o = d.docopt(cli.__doc__, ["kmerize", "-m", "5", "25", "o", "i"], options_first=True, version="v")
    assert o["<command>"] == "kmerize" and o["<args>"] == ["-m", "5", "25", "o", "i"]
    with pytest.raises(SystemExit):
        d.docopt(trim.__doc__, ["trim", "o"])
    with pytest.raises(SystemExit):
        d.docopt(kmerize.__doc__, ["kmerize", "--bogus", "25", "o", "i"])


def test_casket_reads_reference_files_and_rewrites_identically(tmp_path):
    from zotmer_b200.library.kmers import kmers
    for name in ("kat6.k5", "r1_c2.k25", "m2.k25", "m5.k25"):
        with kmers(g(name), "r") as z:
            meta = z.meta
            blobs = {nm: z.open(nm).read() for nm in ("kmers", "counts")}
            toc_order = list(z.toc.keys())
        assert toc_order == ["kmers", "counts", "__meta__"]
        o = tmp_path / name
        with kmers(str(o), "w") as w:
            with w.add_stream("kmers") as f:
                f.write(blobs["kmers"])
            w.add_content("counts", blobs["counts"])
            w.meta = meta
        assert o.read_bytes() == open(g(name), "rb").read()


def test_casket_missing_entry_keyerror():
    from zotmer_b200.library.casket import casket
    with casket(g("kat6.k5"), "r") as z:
        with pytest.raises(KeyError):
            z.open("nope")
        assert z.list() == [("__meta__", 176), ("counts", 8), ("kmers", 8)]


def test_hist_and_info_commands(in_golden_dir, capsys):
    from zotmer_b200 import cli
    cli.main(["hist", "g1.k25", "r1.k25", "r1_c2.k25", "m5.k25", "m2.k25"])
    assert capsys.readouterr().out == open(g("hist.txt")).read()
    cli.main(["info", "kat6.k5", "m3.k25"])
    assert capsys.readouterr().out == open(g("info.txt")).read()


def test_cli_help_and_unknown_command(capsys):
    from zotmer_b200 import cli
    cli.main(["help"])
    out = capsys.readouterr().out
    for c in ("kmerize", "merge", "dist", "jaccard", "trim", "hist", "info", "dump"):
        assert "\t" + c in out
    cli.main(["help", "trim"])
    assert "zot trim [-c CUTOFF] <output> <input>" in capsys.readouterr().out
    cli.main(["frobnicate"])
    assert "unable to load command `frobnicate'" in capsys.readouterr().err


def test_reads_rules_and_pieces():
    from zotmer_b200.library.reads import isFasta, pieces
    assert isFasta("a.fa") and isFasta("a.fasta.gz") and isFasta("x.fna.bz2") and isFasta("y.fas")
    assert not isFasta("a.fq") and not isFasta("a.fastq.gz") and not isFasta("-") and not isFasta("a.fa.txt")
    fq = b"".join(b"@r%d\nACGT\n+\nIIII\n" % i for i in range(1000))
    ps = [bytes(p) for p in pieces(fq, False, max_piece=1000)]
    assert b"".join(ps) == fq and len(ps) > 10
    assert all(p.count(b"\n") % 4 == 0 and p[:1] == b"@" for p in ps)
    fa = b"".join(b">s%d\nACGTACGTAC\nGGG\n" % i for i in range(500))
    ps = [bytes(p) for p in pieces(fa, True, max_piece=777)]
    assert b"".join(ps) == fa and all(p[:1] == b">" for p in ps)


def test_host_formulas_match_reference_vectors():
    from zotmer_b200.library import dist as D
    from zotmer_b200.library import basics as B
    from zotmer_b200.commands import jaccard as J
    kat = json.load(open(g("kat.json")))
    for e in kat["split"]:
        if "jaccard" not in e:
            continue
        a, b, c = e["split"]
        for nm in ("brayCurtis", "chord", "hellinger", "jaccard", "kulczynski", "ochiai", "sorensen", "whittaker"):
            assert float(getattr(D, nm)(a, b, c)).hex() == e["m_" + nm]
    for e in kat["beta"]:
        assert float(J.logIx(e["p"], e["m"], e["n"])).hex() == e["logIx"]
        assert float(J.quantBeta(0.05, e["m"], e["n"])).hex() == e["q05"]
        assert float(J.quantBeta(0.95, e["m"], e["n"])).hex() == e["q95"]
    for e in kat["render"]:
        assert B.render(e["k"], int(e["x"])) == e["out"]
        assert B.renderMany(e["k"], np.array([int(e["x"])], np.uint64)) == [e["out"]]
    for e in kat["rc"]:
        assert str(B.rc(e["k"], int(e["x"]))) == e["out"]
    for e in kat["murmer"]:
        assert str(B.murmer(int(e["x"]), e["s"])) == e["out"]


def test_fasta_line_rules_host_build(tmp_path):
    """the byte-classification rules of the FASTA parse kernel, compiled for the host and fuzzed
    against the oracle's read_fasta (zotmer_b200/csrc/fasta_rules.cuh)"""
    import random
    from oracle import zot_oracle as zo
    exe = str(tmp_path / "parse_rules_host")
    subprocess.check_call(["g++", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host", "parse_rules_host.cpp")])

    def expected(data):
        out = bytearray()
        recs = zo.read_fasta(data)
        for _, seq in recs:
            out.append(4)
            out.extend(4 if zo.NUC[b] is None else zo.NUC[b] for b in seq)
        return bytes(out), len(recs)

    rng = random.Random(5)
    cases = [open(g("g1.fa"), "rb").read(), open(g("kat6.fa"), "rb").read(), b"", b"ACGT", b">", b">a",
             b"\n\n  \n>x\nAC GT\n \n\nAC\n"]
    alphabets = [b"ACGT\n", b"ACGTN \n\r\t>", b"AC >\n", b" \n>", b"ACGTacgtnN>\n\r \t\x0b\x0c-", b"A\n"]
    for it in range(400):
        al = rng.choice(alphabets)
        n = rng.choice([0, 1, 2, 15, 16, 17, 31, 32, 33, 50, 100, 200, 500])
        data = bytes(rng.choices(al, weights=[rng.choice([1, 1, 1, 5, 20]) for _ in al], k=n))
        cases.append((b">h\n" + data) if rng.random() < 0.5 else data)
    fa, co = str(tmp_path / "t.fa"), str(tmp_path / "t.codes")
    for data in cases:
        open(fa, "wb").write(data)
        r = subprocess.run([exe, fa, co], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exp, nrec = expected(data)
        assert open(co, "rb").read() == exp and int(r.stdout) == nrec


def test_compressed_inputs_inflate_like_gunzip(tmp_path):
    """.gz (one member, several members, BGZF in parallel) and .bz2 (one and several streams) give the bytes
    `gunzip -c` / `bunzip2 -c` give (file.py:93-102)"""
    import bz2, gzip, struct, subprocess, zlib
    import numpy as np
    from zotmer_b200.library import file as zfile
    rng = np.random.default_rng(3)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), 100).astype(np.uint8)), b"I" * 100) for i in range(40000))
    # plain gzip, two concatenated members
    p1 = tmp_path / "a.fq.gz"
    p1.write_bytes(gzip.compress(text[:1000000], 1) + gzip.compress(text[1000000:], 1))
    assert zfile.readBytes(str(p1)) == text
    # BGZF: 64 KB members with the BC extra field, as bgzip writes them, plus its empty EOF member
    def member(chunk):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8
        return (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body +
                struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    bg = b"".join(member(text[i:i + 65280]) for i in range(0, len(text), 65280)) + member(b"")
    p2 = tmp_path / "b.fq.gz"
    p2.write_bytes(bg)
    assert len(bg) >= (1 << 20) and zfile._bgzf_blocks(bg) is not None
    assert zfile.readBytes(str(p2)) == text
    assert gzip.decompress(bg) == text                      # it really is a valid multi-member gzip file
    try:
        ref = subprocess.run(["gunzip", "-c", str(p2)], stdout=subprocess.PIPE, check=False).stdout
        assert ref == text
    except FileNotFoundError:
        pass
    # truncated file: the complete members and whatever the broken one still yields, as `gunzip -c` prints before it
    # complains (the reference never looks at its exit status)
    p3 = tmp_path / "c.fq.gz"
    p3.write_bytes(gzip.compress(text[:5000]) + gzip.compress(text[5000:9000])[:-20])
    got = zfile.readBytes(str(p3))
    assert got.startswith(text[:5000]) and text[:9000].startswith(got)
    # bz2, two streams
    p4 = tmp_path / "d.fa.bz2"
    p4.write_bytes(bz2.compress(text[:300000]) + bz2.compress(text[300000:700000]))
    assert zfile.readBytes(str(p4)) == text[:700000]


def test_map_bytes_and_pieces(tmp_path):
    """mapBytes: plain files come back as a read-only mapping that `pieces` can cut (len / rfind / buffer protocol) and
    numpy can view without a copy; empty, compressed and stdin-like inputs as bytes"""
    import gzip
    import numpy as np
    from zotmer_b200.library import file as zfile
    from zotmer_b200.library.reads import pieces
    fq = b"".join(b"@r%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(3000))
    fa = b"".join(b">s%d\nACGTTGCA\nAACC\n" % i for i in range(3000))
    for name, text, is_fa in (("a.fq", fq, False), ("b.fa", fa, True)):
        p = tmp_path / name
        p.write_bytes(text)
        m = zfile.mapBytes(str(p))
        assert not isinstance(m, bytes) and len(m) == len(text)
        assert np.frombuffer(m, dtype=np.uint8).tobytes() == text
        parts = [bytes(x) for x in pieces(m, is_fa, max_piece=4096)]
        assert len(parts) > 5 and b"".join(parts) == text
        assert parts == [bytes(x) for x in pieces(text, is_fa, max_piece=4096)]
        assert all(x.startswith(b">" if is_fa else b"@r") for x in parts)
    e = tmp_path / "e.fq"
    e.write_bytes(b"")
    assert zfile.mapBytes(str(e)) == b""
    z = tmp_path / "z.fq.gz"
    z.write_bytes(gzip.compress(fq))
    assert zfile.mapBytes(str(z)) == fq


def test_casket_entry_views_equal_reads(tmp_path):
    """the word streams of a k-mer set file come back as views of ONE mapping per container (casket.Entry.view), with
    the bytes read() returns; a partly read entry yields its rest; a ragged blob still trips readWords' assertion"""
    import numpy as np
    from zotmer_b200.library.kmers import kmers
    from zotmer_b200.library.casket import casket
    from zotmer_b200.library.files import readWords
    for name in ("kat6.k5", "r1_c2.k25", "m5.k25"):
        with kmers(g(name), "r") as z:
            for nm in ("kmers", "counts"):
                want = z.open(nm).read()
                got = readWords(z.open(nm))
                assert got.dtype == np.dtype("<u8") and got.tobytes() == want
                e = z.open(nm)
                head = e.read(8)
                assert head + bytes(e.view()) == want and e.read() == b""
            assert z._map is not None
            m = z._map
            readWords(z.open("kmers"))
            assert z._map is m                      # mapped once
    p = tmp_path / "ragged.k"
    with casket(str(p), "w") as w:
        w.add_content("kmers", b"\x01" * 13)
        w.add_content("empty", b"")
    with casket(str(p), "r") as z:
        assert len(readWords(z.open("empty"))) == 0
        with pytest.raises(AssertionError):
            readWords(z.open("kmers"))


def test_tools_and_bench_compile():
    """the probes under tools/ and bench.py are run on GPU boxes only: at least they must parse here"""
    import glob
    import py_compile
    for fn in sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        py_compile.compile(fn, doraise=True)


def test_chain_ranges_and_merge_hists():
    """host halves of the range-partitioned encode and of the multi-device stats (no GPU needed)"""
    from zotmer_b200 import _native
    from zotmer_b200.library import devices
    # three ranges; maps: entry state -> (exit state, words).  state 0 in front of the first range
    maps = [([2, 0, 1, 2, 3, 4], [10, 9, 9, 9, 9, 9]), ([0, 1, 2, 3, 4, 5], [0, 0, 0, 0, 0, 0]), ([0, 0, 5, 0, 0, 0], [7, 7, 6, 6, 6, 6])]
    entry, offs, total = _native.chain_ranges(maps)
    assert entry == [0, 2, 2] and offs == [0, 10, 10] and total == 16
    # first-occurrence order along consecutive ranges, frequencies added up
    h = devices.mergeHists([[(3, 5), (1, 2)], [(1, 1), (7, 4)], [], [(3, 1)]])
    assert h == [(3, 6), (1, 3), (7, 4)]


def test_piece_rounds_are_record_aligned(tmp_path):
    """ZB_GPUS=N: every input is cut into ~N record-aligned pieces, dealt out N per round; nothing is lost or doubled"""
    from zotmer_b200.library import devices
    rng = np.random.default_rng(5)
    recs = []
    for i in range(9000):
        L = int(rng.integers(30, 400))
        s = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * L))
    fq = tmp_path / "a.fq"
    fq.write_bytes(b"".join(recs))
    fa = tmp_path / "one_record.fa"     # a single record larger than a share: goes to one device whole
    fa.write_bytes(b">chr\n" + b"\n".join(bytes(rng.choice(list(b"ACGT"), 80).tolist()) for _ in range(30000)) + b"\n")
    fb = tmp_path / "many.fa"
    fb.write_bytes(b"".join(b">s%d\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), 3000).tolist())) for i in range(900)))
    rounds = devices._pieceRounds([str(fq), str(fa), str(fb)], 3)
    assert all(1 <= len(r) <= 3 for r in rounds)
    flat = [(bytes(p), is_fa) for r in rounds for (p, is_fa) in r]
    assert b"".join(p for p, f in flat if not f) == fq.read_bytes()
    assert b"".join(p for p, f in flat if f) == fa.read_bytes() + fb.read_bytes()
    for p, is_fa in flat:
        if is_fa:
            assert p[:1] == b">"
        else:
            assert p.count(b"\n") % 4 == 0 and p[:1] == b"@"
    assert sum(1 for p, f in flat if not f) >= 3 and sum(1 for p, f in flat if f and p.startswith(b">chr")) == 1
    # with the k-mer length known the long record is spread over the devices: cut inside, prefix + piece per cut
    from oracle import c_oracle as co
    rounds = devices._pieceRounds([str(fa)], 3, k=25)
    jobs = [j for r in rounds for j in r]
    assert len(jobs) >= 3
    texts, fake = [], 0
    for (src, is_fa) in jobs:
        if isinstance(src, tuple):
            texts.append(src[0] + bytes(src[1]))
            fake += src[2]
        else:
            texts.append(bytes(src))
    ek, ec, _, enr = co.kmerize(25, [(fa.read_bytes(), True)])
    gk, gc, _, gnr = co.kmerize(25, [(t, True) for t in texts])
    assert np.array_equal(ek, gk) and np.array_equal(ec, gc) and gnr - fake == enr == 1


def test_device_inflater_core_on_the_host_equals_zlib(tmp_path):
    """zotmer_b200/csrc/inflate_core.cuh (what one warp runs per BGZF member) compiled for the host with one lane:
    stored / fixed / dynamic blocks, every zlib strategy, multi-block streams with flush points, wrong sizes and
    corrupt streams (must return an error, never hang or read out of bounds)"""
    import ctypes
    import random
    import zlib
    so = str(tmp_path / "libzi_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(ROOT, "tests", "host", "inflate_host.cpp")])
    L = ctypes.CDLL(so)
    L.zi_inflate_host.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]
    L.zi_inflate_host_lanes.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]
    rng = random.Random(1)
    lanes = [0]

    def inflate(comp, n):
        """one lane, and -- every few calls -- 2, 4 or 8 lanes on as many threads with a real barrier (the lanes' protocol:
        parked literals, two-literal entries, shared copies, synchronisation on overlap); both must agree"""
        out = ctypes.create_string_buffer(n + 72)
        rc = L.zi_inflate_host(comp, len(comp), out, n)
        lanes[0] += 1
        if lanes[0] % 3 == 0:
            w = (2, 4, 8)[(lanes[0] // 3) % 3]
            out2 = ctypes.create_string_buffer(n + 72)
            rc2 = L.zi_inflate_host_lanes(w, comp, len(comp), out2, n)
            assert (rc2 == 0) == (rc == 0), (rc, rc2, w)
            if rc == 0:
                assert out2.raw[:n] == out.raw[:n], w
        return rc, out.raw[:n]

    def check(data, level, strategy, wbits=-15, mem=9):
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategy)
        comp = c.compress(data) + c.flush()
        rc, out = inflate(comp, len(data))
        assert rc == 0 and out == data, (rc, len(data), level, strategy)
        assert inflate(comp, len(data) + 1)[0] != 0
        if data:
            assert inflate(comp, len(data) - 1)[0] != 0

    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGT") for _ in range(150)), b"I" * 150) for i in range(200))
    cases = [b"", b"a", b"ab" * 10, b"\x00" * 70000, bytes(rng.getrandbits(8) for _ in range(65536)), fq[:65536], b"ACGT" * 16384]
    for n in [1, 2, 3, 17, 257, 258, 259, 4095, 32768, 65535]:
        cases.append(bytes(rng.choice(b"ACGTN\n") for _ in range(n)))
        cases.append(bytes(rng.getrandbits(8) for _ in range(n)))
        cases.append(bytes(rng.choice(b"ab") for _ in range(n)))
    for d in cases:
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                check(d, level, strat)
    for trial in range(100):
        c = zlib.compressobj(rng.choice([1, 6, 9]), zlib.DEFLATED, -15, rng.choice([1, 5, 9]),
                             rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED]))
        parts, data = [], b""
        for j in range(rng.randint(1, 6)):
            al = rng.choice([b"ACGT\n", b"ab", bytes(range(256)), b"ACGTNIIIII@+\n0123456789"])
            d = bytes(rng.choice(al) for _ in range(rng.choice([0, 1, 10, 300, 5000])))
            data += d
            parts.append(c.compress(d))
            parts.append(c.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH, zlib.Z_BLOCK])))
        parts.append(c.flush())
        comp = b"".join(parts)
        rc, out = inflate(comp, len(data))
        assert rc == 0 and out == data
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    base = c.compress(fq[:30000]) + c.flush()
    rejected = 0
    for trial in range(1500):
        b = bytearray(base)
        for _ in range(rng.randint(1, 4)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        if rng.random() < 0.3:
            b = b[:rng.randrange(1, len(b))]
        rejected += inflate(bytes(b), 30000)[0] != 0
    assert rejected > 500


def test_bgzf_probe_without_a_device():
    """BGZF members are recognised from their headers on the host (zb_bgzf_probe); an ordinary .gz is not BGZF"""
    import gzip
    from tools import synth
    from zotmer_b200 import _native as nat
    from zotmer_b200.library.file import gunzipBytes
    data = b"@r\nACGT\n+\nIIII\n" * 9000
    z = synth.bgzf_bytes(data, block=10000)
    assert gzip.decompress(z) == data and gunzipBytes(z) == data
    assert nat.bgzf_probe(z) == (len(data) // 10000 + 1 + 1, len(data))
    assert nat.bgzf_probe(synth.bgzf_bytes(data, eof=False)) == ((len(data) + 65279) // 65280, len(data))
    assert nat.bgzf_probe(gzip.compress(data)) is None
    assert nat.bgzf_probe(z[:-1]) is None and nat.bgzf_probe(z + b"\x00" * 40) is None and nat.bgzf_probe(b"") is None
    # runs of whole members for several devices (zb_bgzf_groups): they tile the file, each holds <= max_text bytes of text
    for max_text in (10000, 35000, 70000, 10 ** 9):
        starts = nat.bgzf_groups(z, max_text)
        assert starts[0] == 0 and starts == sorted(set(starts))
        total = 0
        for a, b in zip(starts, starts[1:] + [len(z)]):
            m, t = nat.bgzf_probe(z[a:b])
            assert m >= 1 and (t <= max(max_text, 65536) or m == 1)
            assert gzip.decompress(z[a:b]) == data[total:total + t]
            total += t
        assert total == len(data)
    with pytest.raises(Exception):
        nat.bgzf_groups(gzip.compress(data), 70000)


def test_bgzf_groups_handed_from_device_to_device(tmp_path, monkeypatch):
    """library/devices.py: the runs of members of a block-compressed file are dealt out to the devices; every device
    puts the leftover of the previous run in front of its text, cuts at its last record boundary and hands the rest on.
    The device calls (inflate, concat, cut) are stood in for by host code here; the threads, the order of the hand-over
    and the cuts are the real ones.  (The multi-GPU tests that run this on devices need more than one GPU.)"""
    import gzip
    import threading
    from tools import synth
    from zotmer_b200.library import devices, reads

    class FakeStaged(object):
        def __init__(self, data):
            self.data = bytes(data)

        def __len__(self):
            return len(self.data)

        def cut(self, fa):
            d = self.data
            if fa:
                return d.rfind(b"\n>") + 1
            nl = d.count(b"\n")
            if nl < 4:
                return 0
            at = len(d)
            for _ in range(nl % 4 + 1):
                at = d.rfind(b"\n", 0, at)
            return at + 1

        def fetch_range(self, off, n):
            return self.data[off:off + n]

        def set_len(self, n):
            self.data = self.data[:n]

        def free(self):
            self.data = b""

    class FakeNative(object):
        bgzf_probe = staticmethod(devices._native.bgzf_probe)
        bgzf_groups = staticmethod(devices._native.bgzf_groups)

        @staticmethod
        def stage_bgzf(comp, dev):
            return FakeStaged(gzip.decompress(bytes(comp))), len(comp)

        @staticmethod
        def stage_concat(prefix, body, dev):
            return FakeStaged(bytes(prefix) + body.data)

    monkeypatch.setattr(devices, "_native", FakeNative)
    monkeypatch.setattr(reads, "BGZF_GROUP", 1 << 20)
    g = synth.genome(400000, seed=31)
    fq = synth.fastq_array(g, 12000, seed=32).reshape(-1).tobytes()[:-1]            # no final newline
    big = synth.genome(2600000, seed=34)        # one record longer than two runs: they hold no record boundary
    fa = synth.fasta_bytes(g) + b">p2\n" + synth.fasta_bytes(big)[6:] + b">p3\n" + synth.fasta_bytes(synth.genome(90000, seed=33))[6:] + b">p4\nACGT"
    for name, text, is_fa in (("a.fq.gz", fq, False), ("b.fa.gz", fa, True)):
        fn = tmp_path / name
        fn.write_bytes(synth.bgzf_bytes(text, block=50000))
        for n in (2, 3, 8):
            jobs = []
            assert devices._bgzfJobs(str(fn), is_fa, n, jobs) and len(jobs) >= 2
            out = [None] * len(jobs)
            errors = []

            def work(j):
                grp = jobs[j][0]
                try:
                    st = devices._stageBgzfGroup(grp, jobs[j - 1][0] if not grp.first else None, is_fa, j % n, errors)
                    out[j] = st.data if st is not None else b""
                except BaseException as e:      # noqa: B902
                    errors.append(e)

            # rounds of n jobs, the devices of a round at once, later rounds' threads started before earlier ones end
            ths = [threading.Thread(target=work, args=(j,)) for j in reversed(range(len(jobs)))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            assert not errors, errors
            assert b"".join(out) == text
            for piece in out[:-1]:
                if is_fa:
                    assert piece == b"" or (piece[:1] == b">" and piece[-1:] == b"\n")
                else:
                    assert piece.count(b"\n") % 4 == 0 and (piece == b"" or piece[-1:] == b"\n")


def test_long_fasta_records_are_cut_inside():
    """library/reads.py:splitPieces -- a FASTA record longer than a piece is cut after a line (or inside a line longer than
    the piece); the next piece starts with an empty header + the k - 1 characters before the cut.  Feeding the pieces
    must give the k-mers, counts and (after subtracting the extra headers) records of the whole file: oracle on both
    sides, with the reference's line rules (indented headers, '>' and blanks inside lines, CRLF, blank lines, text
    before the first header)."""
    import random
    from oracle import c_oracle as co
    from zotmer_b200.library.reads import splitPieces

    def rnd_fasta(rng, nrec, maxlen, linelens, weird):
        out = []
        if weird and rng.random() < 0.3:
            out.append(b"ACGTACGTAAAA junk before the first header\n")
        for r in range(nrec):
            out.append(b">rec%d some text\n" % r if not (weird and rng.random() < 0.2) else b"  >indented header\n")
            L = rng.randrange(0, maxlen)
            seq = bytes(rng.choice(b"ACGT" if rng.random() < 0.97 else b"NnRY") for _ in range(L))
            i = 0
            while i < L:
                w = rng.choice(linelens)
                line = seq[i:i + w]
                i += w
                if weird and rng.random() < 0.1:
                    line = b"  " + line + b" \t"
                if weird and rng.random() < 0.05 and len(line) > 4:
                    line = line[:2] + b" " + line[2:]
                if weird and rng.random() < 0.02 and len(line) > 4:
                    line = line[:3] + b">" + line[3:]
                out.append(line + (b"\r\n" if weird and rng.random() < 0.1 else b"\n"))
                if weird and rng.random() < 0.03:
                    out.append(b"\n")
        data = b"".join(out)
        return data[:-1] if rng.random() < 0.5 and data.endswith(b"\n") else data

    rng = random.Random(11)
    done = 0
    for trial in range(260):
        data = rnd_fasta(rng, rng.randint(1, 4), rng.choice([50, 400, 3000]), rng.choice([[60], [7, 13], [100000], [1, 2, 3]]), trial % 2 == 1)
        k = rng.choice([1, 2, 5, 16, 25, 31, 32])
        mp = rng.choice([40, 97, 300, 1000])
        if len(data) <= mp:
            continue
        ps = list(splitPieces(data, True, k, mp))
        assert b"".join(bytes(v) for _, v, _ in ps) == data and all(len(v) <= mp for _, v, _ in ps)
        fake = sum(f for _, _, f in ps)
        ek, ec, _, enr = co.kmerize(k, [(data, True)])
        gk, gc, _, gnr = co.kmerize(k, [(pre + bytes(v), True) for pre, v, _ in ps])
        assert np.array_equal(ek, gk) and np.array_equal(ec, gc) and gnr - fake == enr, (trial, k, mp)
        done += 1
    assert done > 150
    # FASTQ and short inputs go through `pieces` unchanged
    fq = b"@r\nACGT\n+\nIIII\n" * 50
    assert [(p, bytes(v), f) for p, v, f in splitPieces(fq, False, 25, 100)] == [(b"", bytes(v), 0) for v in __import__("zotmer_b200.library.reads", fromlist=["x"]).pieces(fq, False, 100)]


def test_bgzf_pieces_groups_carry_and_damage(monkeypatch):
    """library/reads.py:bgzfPieces with the device calls stood in for by host code: groups of members, the incomplete
    record carried to the next piece, a record that fills several groups, and a damaged member (the input ends where
    `gunzip -c` would have stopped writing, file.py:93-97)"""
    import gzip
    from tools import synth
    from zotmer_b200.library import reads
    from zotmer_b200.library.file import gunzipBytes
    from zotmer_b200 import _native as real_nat

    class FakeStaged(object):
        def __init__(self, data):
            self.data = bytes(data)

        def __len__(self):
            return len(self.data)

        def cut(self, fa):
            d = self.data
            if fa:
                return d.rfind(b"\n>") + 1
            nl = d.count(b"\n")
            if nl < 4:
                return 0
            at = len(d)
            for _ in range(nl % 4 + 1):
                at = d.rfind(b"\n", 0, at)
            return at + 1

        def fetch_range(self, off, n):
            return self.data[off:off + n]

        def set_len(self, n):
            self.data = self.data[:n]

        def free(self):
            self.data = b""

    class FakeNative(object):
        @staticmethod
        def stage_bgzf(data, device=0, max_out=0, carry=None, carry_off=0):
            nat = real_nat
            a = bytes(data)
            head = carry.data[carry_off:] if carry is not None else b""
            # whole members while the text fits max_out (at least one), as zb_stage_bgzf takes them
            starts = nat.bgzf_groups(a, 65536)
            text, used = b"", 0
            for s0, s1 in zip(starts, starts[1:] + [len(a)]):
                try:
                    t = gzip.decompress(a[s0:s1])
                except Exception:
                    raise AssertionError("BGZF member does not inflate")
                if used and max_out and len(head) + len(text) + len(t) > max_out:
                    break
                text += t
                used = s1
            return FakeStaged(head + text), used

        @staticmethod
        def stage_input(data, device=0):
            return FakeStaged(bytes(data))

    import zotmer_b200
    monkeypatch.setattr(zotmer_b200, "_native", FakeNative, raising=False)
    g = synth.genome(300000, seed=51)
    fq = synth.fastq_array(g, 4000, seed=52).reshape(-1).tobytes()
    fa = synth.fasta_bytes(g) + b">p2\nACGT\nAC\n>p3 x\nGGGG"
    for text, is_fa in ((fq, False), (fa, True)):
        z = synth.bgzf_bytes(text, block=20000)
        for group in (70000, 150000, 10 ** 9):
            out = [st.data for st in reads.bgzfPieces(z, is_fa, 0, group=group)]
            assert b"".join(out) == text
            for piece in out[:-1]:
                assert piece[-1:] == b"\n" and (piece[:1] == b">" if is_fa else piece.count(b"\n") % 4 == 0)
    # a damaged member in the third group
    z = bytearray(synth.bgzf_bytes(fq, block=20000))
    p, k = 0, 0
    while k < 9:
        p += (z[p + 16] | (z[p + 17] << 8)) + 1
        k += 1
    z[p + 40:p + 90] = b"\xff" * 50
    want = gunzipBytes(bytes(z))
    assert 0 < len(want) < len(fq)
    out = b"".join(st.data for st in reads.bgzfPieces(bytes(z), False, 0, group=70000))
    assert out == want
