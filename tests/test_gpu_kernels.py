"""
GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through the C ABI
(zotmer_b200/_native.py -> libzot_b200.so) and compares with the oracle (oracle/) on the same
seeded inputs, or with the committed golden fixtures produced by the reference itself.
Bit-exact: everything on this path is integer work.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import c_oracle as co
from oracle import zot_oracle as zo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from zotmer_b200 import _native
    assert _native.device_count() >= 1, "no CUDA device"
    return _native


def g(name):
    return os.path.join(GOLDEN, name)


def rd(name):
    with open(g(name), "rb") as f:
        return f.read()


def rnd_dna(rng, n):
    return np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()


def make_fastq(rng, genome, nreads, L, err=0.01, pn=0.002):
    out = []
    G = np.frombuffer(genome, np.uint8)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    for i in range(nreads):
        p = int(rng.integers(0, len(G) - L))
        s = G[p:p + L].copy()
        if rng.random() < 0.5:
            s = comp[s[::-1]]
        e = rng.random(L) < err
        s[e] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(e.sum()))]
        nn = rng.random(L) < pn
        s[nn] = ord("N")
        out.append(b"@r%d\n%s\n+\n%s\n" % (i, s.tobytes(), b"I" * L))
    return b"".join(out)


def make_fasta(rng, n, width=80, nrec=3):
    out = []
    for r in range(nrec):
        s = rnd_dna(rng, n // nrec)
        out.append(b">rec%d\n" % r + b"\n".join(s[i:i + width] for i in range(0, len(s), width)) + b"\n")
    return b"".join(out)


# ----------------------------------------------------------------------------- radix sort
@pytest.mark.parametrize("n", [1, 5, 4095, 4096, 4097, 100000, 1 << 20, 3000001])
@pytest.mark.parametrize("bits,maxbits", [(50, 8), (64, 8), (62, 10), (50, 10), (10, 8), (50, 11), (33, 9)])
def test_sort_keys(nat, n, bits, maxbits):
    rng = np.random.default_rng(n * 131 + bits)
    keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    out, _, _ = nat.dbg_sort(keys, None, bits, maxbits)
    assert np.array_equal(out, np.sort(keys))


@pytest.mark.parametrize("n", [7, 4097, 250000])
def test_sort_pairs_stable(nat, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 50, n, dtype=np.uint64)  # many duplicates: stability is visible through the payload
    vals = np.arange(n, dtype=np.uint32)
    out, v, _ = nat.dbg_sort(keys, vals, 50, 8)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(out, keys[order])
    assert np.array_equal(v, vals[order])


def test_sort_skewed(nat):
    n = 500000
    rng = np.random.default_rng(3)
    keys = np.zeros(n, np.uint64)
    keys[rng.integers(0, n, 1000)] = rng.integers(0, 2 ** 50, 1000, dtype=np.uint64)
    out, _, _ = nat.dbg_sort(keys, None, 50, 8)
    assert np.array_equal(out, np.sort(keys))
    keys[:] = np.uint64(2 ** 64 - 1)
    out, _, _ = nat.dbg_sort(keys, None, 64, 8)
    assert np.array_equal(out, keys)


# ----------------------------------------------------------------------------- sort + count (segsort.cu)
def np_count(keys, weights=None):
    if len(keys) == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    u, inv = np.unique(keys, return_inverse=True)
    w = np.ones(len(keys), np.uint64) if weights is None else weights.astype(np.uint64)
    c = np.zeros(len(u), np.uint64)
    np.add.at(c, inv, w)
    return u, c


def check_sort_count(nat, keys, bits, weights=None):
    ek, ec = np_count(keys, weights)
    for mode in (0, 3, 1):   # bucket route, segment route and the classic full sort + RLE must agree with numpy and each other
        k, c, _ = nat.dbg_sort_count(keys, weights, bits, mode)
        assert np.array_equal(k, ek), "mode %d: keys differ" % mode
        assert np.array_equal(c.astype(np.uint64), ec), "mode %d: counts differ" % mode


@pytest.mark.parametrize("n", [1, 2, 17, 3071, 3072, 3073, 4096, 4097, 6144, 100000, 1 << 20, 2500001])
@pytest.mark.parametrize("bits,dup", [(50, 1), (50, 6), (62, 3), (64, 1), (32, 2), (24, 1)])
def test_sort_count_random(nat, n, bits, dup):
    rng = np.random.default_rng(n * 7 + bits + dup)
    m = max(1, n // dup)
    base = rng.integers(0, 2 ** 63, m, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, m, dtype=np.uint64)
    if bits < 64:
        base &= np.uint64((1 << bits) - 1)
    keys = base[rng.integers(0, m, n)]
    check_sort_count(nat, keys, bits)


@pytest.mark.parametrize("seglen", [1, 2, 15, 16, 17, 100, 511, 513, 1023, 1024, 1025, 3072, 5000])
def test_sort_count_segment_lengths(nat, seglen):
    """keys sharing their top bits in groups of exactly `seglen`: around the 1024-key limit of the in-shared-memory
    path and across CTA boundaries (3072 owned positions), distinct and duplicated low parts"""
    bits = 50
    nseg = max(3, 60000 // seglen)
    rng = np.random.default_rng(seglen)
    n = nseg * seglen
    lg = int(np.ceil(np.log2(n)))
    T = 8 * max(1, (lg - 4 + 7) // 8)
    low = bits - T
    assert low >= 8
    tops = rng.choice(1 << min(T, 20), nseg, replace=False).astype(np.uint64)
    for distinct_low in (True, False):
        if distinct_low:
            lows = rng.integers(0, 1 << low, n, dtype=np.uint64)
        else:
            lows = rng.integers(0, 3, n, dtype=np.uint64) * np.uint64(12345)
        keys = (np.repeat(tops, seglen) << np.uint64(low)) | lows
        rng.shuffle(keys)
        check_sort_count(nat, keys, bits)


def test_sort_count_heavy_repeats(nat):
    """a few keys repeated very often (big segments -> side path) mixed with unique keys"""
    rng = np.random.default_rng(11)
    n = 700000
    keys = rng.integers(0, 2 ** 50, n, dtype=np.uint64)
    keys[rng.integers(0, n, 200000)] = np.uint64(0)                       # poly-A like
    keys[rng.integers(0, n, 50000)] = np.uint64((1 << 50) - 1)
    hot = rng.integers(0, 2 ** 50, 40, dtype=np.uint64)
    keys[rng.integers(0, n, 100000)] = hot[rng.integers(0, 40, 100000)]   # ~2500 copies each
    check_sort_count(nat, keys, 50)
    check_sort_count(nat, np.full(100000, 7, np.uint64), 50)              # one key only
    check_sort_count(nat, np.full(1025, (1 << 64) - 1, np.uint64), 64)


@pytest.mark.parametrize("dups", [1, 5])
def test_sort_count_skewed_buckets(nat, dups):
    """bucket route: key-range buckets of 2x, 5x, 25x and 75x the average size (done in 2..16 rounds inside the kernel,
    the last one handed to the segment route), keys only and distinct keys with payload"""
    rng = np.random.default_rng(90 + dups)
    n = 300000
    bits = 50
    parts = [rng.integers(0, 2 ** bits, n * 50 // 100, dtype=np.uint64)]
    for top, share in ((0x03, 0.30), (0x57, 0.10), (0xa1, 0.02), (0xfe, 0.008)):
        m = int(n * share)
        parts.append((np.uint64(top) << np.uint64(bits - 8)) | rng.integers(0, 2 ** (bits - 8), m, dtype=np.uint64))
    base = np.unique(np.concatenate(parts))
    keys = base[rng.integers(0, len(base), len(base) * dups)] if dups > 1 else base.copy()
    rng.shuffle(keys)
    check_sort_count(nat, keys, bits)
    # distinct + payload
    w = rng.integers(0, 2 ** 32 - 1, len(base), dtype=np.uint32)
    perm = rng.permutation(len(base))
    for mode in (2, 4):
        k, c, _ = nat.dbg_sort_count(base[perm], w[perm], bits, mode)
        assert np.array_equal(k, base) and np.array_equal(c, w)
    # canonical-like skew at a size where the fullest buckets pass the capacity: 7/16 of the keys below 2^48
    n2 = 2200000
    top2 = rng.choice(4, n2, p=[7 / 16, 5 / 16, 3 / 16, 1 / 16]).astype(np.uint64)
    keys2 = (top2 << np.uint64(48)) | rng.integers(0, 2 ** 48, n2, dtype=np.uint64)
    check_sort_count(nat, keys2, bits)


@pytest.mark.parametrize("n", [5, 4097, 300000])
def test_sort_count_weighted(nat, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 50, max(1, n // 2), dtype=np.uint64)[rng.integers(0, max(1, n // 2), n)]
    w = rng.integers(1, 1000, n, dtype=np.uint32)
    check_sort_count(nat, keys, 50, w)
    # distinct keys with payload = the mirror sort of kmerize
    keys = rng.permutation(n).astype(np.uint64) * np.uint64(1099511627791) & np.uint64((1 << 50) - 1)
    keys = np.unique(keys)
    rng.shuffle(keys)
    w = rng.integers(1, 2 ** 32 - 1, len(keys), dtype=np.uint32)
    check_sort_count(nat, keys, 50, w)


@pytest.mark.parametrize("n", [3, 5000, 400000])
def test_sort_count_distinct_payload(nat, n):
    """mode 2: keys promised distinct, weights are a payload (the mirror sort of kmerize)"""
    rng = np.random.default_rng(n + 1)
    keys = np.unique(rng.integers(0, 2 ** 50, n, dtype=np.uint64))
    rng.shuffle(keys)
    w = rng.integers(0, 2 ** 32 - 1, len(keys), dtype=np.uint32)
    order = np.argsort(keys)
    for mode in (2, 4):   # bucket route, segment route
        k, c, _ = nat.dbg_sort_count(keys, w, 50, mode)
        assert np.array_equal(k, keys[order]) and np.array_equal(c, w[order])
        if n >= 5000:   # a broken promise is reported, not silently mis-sorted
            bad = keys.copy()
            bad[len(bad) // 2] = bad[len(bad) // 2 + 1]
            with pytest.raises(Exception):
                nat.dbg_sort_count(bad, w, 50, mode)


def test_sort_count_weight_overflow(nat):
    keys = np.full(3, 12345678901, np.uint64)
    keys = np.concatenate([keys, np.arange(5000, dtype=np.uint64) << np.uint64(20)])
    w = np.full(len(keys), 2 ** 31, np.uint32)
    for mode in (0, 3, 1):
        with pytest.raises(IndexError):
            nat.dbg_sort_count(keys, w, 50, mode)


# ----------------------------------------------------------------------------- parse
def expected_fasta_codes(data):
    out = bytearray()
    recs = zo.read_fasta(data)
    for _, seq in recs:
        out.append(4)
        out.extend(4 if zo.NUC[b] is None else zo.NUC[b] for b in seq)
    return bytes(out), len(recs)


def kmers_from_codes(codes, k):
    """canonical k-mers of a code stream (python, small inputs)"""
    out = []
    x = 0
    run = 0
    msk = (1 << (2 * k)) - 1
    for c in codes:
        if c > 3:
            run = 0
            x = 0
            continue
        x = ((x << 2) | int(c)) & msk
        run += 1
        if run >= k:
            out.append(min(x, zo.rc(k, x)))
    return sorted(out)


@pytest.mark.parametrize("name", ["g1.fa", "kat6.fa", "s0.fa", "j.fa", "empty.fa"])
def test_parse_fasta_golden_inputs(nat, name):
    data = rd(name)
    codes, nrec = nat.dbg_parse(data, True)
    exp, enrec = expected_fasta_codes(data)
    assert nrec == enrec
    assert codes.tobytes() == exp


def test_parse_fasta_fuzz(nat):
    import random
    rng = random.Random(11)
    alphabets = [b"ACGT\n", b"ACGTN \n\r\t>", b"AC >\n", b" \n>", b"ACGTacgtnN>\n\r \t\x0b\x0c-", b"A\n"]
    for it in range(150):
        al = rng.choice(alphabets)
        n = rng.choice([1, 15, 16, 17, 100, 4095, 4096, 4097, 16383, 16384, 16385, 40000, 70000])
        w = [rng.choice([1, 1, 1, 5, 20]) for _ in al]
        data = bytes(rng.choices(al, weights=w, k=n))
        if rng.random() < 0.5:
            data = b">h\n" + data
        codes, nrec = nat.dbg_parse(data, True)
        exp, enrec = expected_fasta_codes(data)
        assert nrec == enrec, it
        assert codes.tobytes() == exp, it


def test_parse_fasta_long_lines_and_blank_runs(nat):
    rng = np.random.default_rng(4)
    one_line = b">single line chromosome\n" + rnd_dna(rng, 300000) + b"\n"
    blanks = b">a\nACGTACGTAC" + b" \n" * 30000 + b"GTACGTACGT\n>b  " + b"x" * 50000 + b"\nAC" + b" " * 40000 + b"GT\n"
    for data in (one_line, blanks, one_line + blanks):
        codes, nrec = nat.dbg_parse(data, True)
        exp, enrec = expected_fasta_codes(data)
        assert nrec == enrec
        assert codes.tobytes() == exp


@pytest.mark.parametrize("name", ["r1.fq", "r2.fq"])
def test_parse_fastq_golden_inputs(nat, name):
    data = rd(name)
    codes, nrec = nat.dbg_parse(data, False)
    recs = zo.read_fastq(data)
    assert nrec == len(recs)
    for k in (8, 25):
        exp = sorted(min(x, zo.rc(k, x)) for r in recs for x in zo.kmers_list(k, r[1], False))
        assert kmers_from_codes(codes, k) == exp


def test_parse_fastq_edge_cases(nat):
    cases = [b"", b"\n", b"@a\nACGT\n+\nIIII", b"@a\nACGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\n",
             b"@a\nACGT\n+\nIIII\n@b\nACGT", b"\n\n\n\n\n\n\n\n", b"@a\r\nAC GT\r\n+\r\nIIII\r\n" * 3,
             b"@a\nACGTACGTAC\n+\nIIIIIIIIII\n" * 5000 + b"@tail\nACGTACGT\n"]
    for data in cases:
        codes, nrec = nat.dbg_parse(data, False)
        recs = zo.read_fastq(data)
        assert nrec == len(recs), data[:40]
        exp = sorted(min(x, zo.rc(4, x)) for r in recs for x in zo.kmers_list(4, r[1], False))
        assert kmers_from_codes(codes, 4) == exp, data[:40]


# ----------------------------------------------------------------------------- extract
@pytest.mark.parametrize("k", [1, 2, 5, 16, 25, 31, 32])
def test_extract(nat, k):
    rng = np.random.default_rng(k)
    for n in (0, 1, 31, 32, 33, 4095, 4096, 4097, 20000):
        codes = rng.integers(0, 4, n, dtype=np.uint8)
        codes[rng.random(n) < 0.01] = 4
        got = np.sort(nat.dbg_extract(k, codes))
        assert [int(x) for x in got] == kmers_from_codes(codes, k), (k, n)


# ----------------------------------------------------------------------------- kmerize end to end
KMERIZE = [(5, "kat6.k5", ["kat6.fa"]), (5, "g1.k5", ["g1.fa"]), (16, "g1.k16", ["g1.fa"]), (25, "g1.k25", ["g1.fa"]),
           (30, "g1.k30", ["g1.fa"]), (31, "g1.k31", ["g1.fa"]), (32, "g1.k32", ["g1.fa"]),
           (8, "r1.k8", ["r1.fq"]), (25, "r1.k25", ["r1.fq"]), (31, "r1.k31", ["r1.fq"]),
           (21, "r2.k21", ["r2.fq"]), (25, "mix.k25", ["s0.fa", "r1.fq", "s1.fa"])]


def run_kmerize(nat, k, inputs):
    km = nat.Kmerizer(k)
    for data, is_fa in inputs:
        km.feed(data, is_fa)
    s, nr = km.finish()
    km.close()
    return s, nr


@pytest.mark.parametrize("k,out,ins", KMERIZE)
def test_kmerize_golden(nat, k, out, ins):
    z = zo.CasketReader(g(out))
    xs, cs = zo.read_kmers_and_counts(z)
    s, nr = run_kmerize(nat, k, [(rd(i), zo.is_fasta(i)) for i in ins])
    ks, cc = s.fetch()
    assert nr == z.meta["reads"]
    assert [int(x) for x in ks] == xs
    assert [int(c) for c in cc] == cs
    st = s.stats()
    assert [(str(v), f) for v, f in st["hist"]] == list(z.meta["hist"].items())
    n = float(sum(st["acgt_weighted"]))
    assert [a / n for a in st["acgt_weighted"]] == z.meta["acgt"]
    kw, cw = s.encode()
    assert kw.tobytes() == z.blob("kmers") and cw.tobytes() == z.blob("counts")


@pytest.mark.parametrize("k", [25, 31, 12])
def test_kmerize_fastq_vs_c_oracle(nat, k):
    rng = np.random.default_rng(100 + k)
    genome = rnd_dna(rng, 200000)
    fq = make_fastq(rng, genome, 20000, 150)
    s, nr = run_kmerize(nat, k, [(fq, False)])
    ks, cc = s.fetch()
    ek, ec, eacgt, enr = co.kmerize(k, [(fq, False)])
    assert nr == enr
    assert np.array_equal(ks, ek) and np.array_equal(cc, ec)
    st = s.stats()
    assert st["acgt_weighted"] == eacgt
    assert st["hist"] == co.hist(ec.astype(np.uint64))


@pytest.mark.parametrize("k", [25, 32, 6])
def test_kmerize_fasta_vs_c_oracle(nat, k):
    rng = np.random.default_rng(200 + k)
    fa = make_fasta(rng, 1500000)
    s, nr = run_kmerize(nat, k, [(fa, True)])
    ks, cc = s.fetch()
    ek, ec, eacgt, enr = co.kmerize(k, [(fa, True)])
    assert nr == enr
    assert np.array_equal(ks, ek) and np.array_equal(cc, ec)
    assert s.stats()["acgt_weighted"] == eacgt


@pytest.mark.parametrize("max_runs", [None, 3, 8])
def test_kmerize_batched_flush(nat, monkeypatch, max_runs):
    """small ZB_MAX_PENDING forces several sort+count rounds; the runs are united at the end (default: they fit the
    memory budget) or every max_runs batches; the result must not change"""
    rng = np.random.default_rng(7)
    genome = rnd_dna(rng, 50000)
    fq = make_fastq(rng, genome, 8000, 100)
    monkeypatch.setenv("ZB_MAX_PENDING", "65536")
    if max_runs:
        monkeypatch.setenv("ZB_MAX_RUNS", str(max_runs))
    s, nr = run_kmerize(nat, 21, [(fq, False), (fq[:len(fq) // 2 - (len(fq) // 2) % 1], False)])
    monkeypatch.delenv("ZB_MAX_PENDING")
    ks, cc = s.fetch()
    ek, ec, _, enr = co.kmerize(21, [(fq, False), (fq[:len(fq) // 2], False)])
    assert nr == enr
    assert np.array_equal(ks, ek) and np.array_equal(cc, ec)


@pytest.mark.parametrize("k", [25, 24, 31])
def test_kmerize_low_complexity(nat, k):
    """homopolymers, dinucleotide repeats and a tiny alphabet: a handful of k-mers with counts in the hundreds of thousands
    (segments far beyond the shared-memory limit of the sort+count kernel), mixed with ordinary reads"""
    rng = np.random.default_rng(k)
    L = 100
    reads = []
    reads += [b"A" * L] * 3000 + [b"T" * L] * 2500 + [b"AT" * (L // 2)] * 2000 + [b"ACG" * (L // 3) + b"A"] * 1500
    genome = rnd_dna(rng, 20000)
    two = np.frombuffer(b"AC", np.uint8)[rng.integers(0, 2, 30000)].tobytes()
    for i in range(6000):
        p = int(rng.integers(0, len(genome) - L))
        reads.append(genome[p:p + L])
        p = int(rng.integers(0, len(two) - L))
        reads.append(two[p:p + L])
    order = rng.permutation(len(reads))
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, reads[i], b"I" * L) for i in order)
    s, nr = run_kmerize(nat, k, [(fq, False)])
    ks, cc = s.fetch()
    ek, ec, eacgt, enr = co.kmerize(k, [(fq, False)])
    assert nr == enr
    assert np.array_equal(ks, ek) and np.array_equal(cc, ec)
    assert int(cc.max()) > 100000
    st = s.stats()
    assert st["acgt_weighted"] == eacgt and st["hist"] == co.hist(ec.astype(np.uint64))


@pytest.mark.parametrize("skew", [False, True])
def test_kmerize_mirror_routes(nat, monkeypatch, skew):
    """the both-strand union at the end of kmerize: the fused route (mirrored half ordered by its top bits only, united with
    the canonical half inside key-range buckets in shared memory) against the classic route (ZB_SORT_COUNT=1: full sort of
    the mirrored half + merge path) and the oracle.  skew: a two-letter genome puts most canonical k-mers into 16 of 256
    buckets, which overflow, so the fused route must hand over to sort + merge after its top-bit passes."""
    rng = np.random.default_rng(77 + skew)
    genome = rnd_dna(rng, 150000)
    reads = make_fastq(rng, genome, 12000, 150)
    if skew:
        two = np.frombuffer(b"AC", np.uint8)[rng.integers(0, 2, 300000)].tobytes()
        reads += make_fastq(rng, two, 24000, 150)
    ek, ec, _, enr = co.kmerize(25, [(reads, False)])
    for mode in (None, "1"):
        if mode is None:
            monkeypatch.delenv("ZB_SORT_COUNT", raising=False)
        else:
            monkeypatch.setenv("ZB_SORT_COUNT", mode)
        s, nr = run_kmerize(nat, 25, [(reads, False)])
        ks, cc = s.fetch()
        assert nr == enr
        assert np.array_equal(ks, ek) and np.array_equal(cc, ec), mode
    monkeypatch.delenv("ZB_SORT_COUNT", raising=False)
    nat.Kmerizer(25, 0).close()     # re-read the switches: back to the default route


def test_kmerize_palindromes_even_k(nat):
    pal = b"ACGTTGCAAGCTTGCAACGT"
    fa = b">p\n" + pal * 50 + b"\n>q\n" + b"AT" * 100 + b"\n"
    for k in (2, 4, 8, 20):
        s, _ = run_kmerize(nat, k, [(fa, True)])
        ks, cc = s.fetch()
        ek, ec, _, _ = co.kmerize(k, [(fa, True)])
        assert np.array_equal(ks, ek) and np.array_equal(cc, ec), k


@pytest.mark.parametrize("k", [9, 25, 32])
def test_kmerize_capture_vs_oracle(nat, k):
    """capture mode on a few thousand reads / a multi-record FASTA spanning many tiles: every record is judged on its
    own (kmerize.py:507-517), records are told apart by the parser's record marks"""
    rng = np.random.default_rng(k)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, 60000))
    baits = ">b\n" + genome[1000:1200] + "\n>c\n" + genome[30000:30100] + "\n"
    recs = []
    for i in range(4000):
        L = int(rng.choice([150, 150, 80, 40, k, k - 1, 0]))
        p = int(rng.integers(0, len(genome) - 150))
        r = genome[p:p + L]
        if L > 20 and rng.random() < 0.2:
            q = int(rng.integers(0, L))
            r = r[:q] + "N" + r[q + 1:]
        recs.append(r)
    fq = "".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(recs)).encode()
    fa = "".join(">s%d\n%s\n" % (i, "\n".join(r[j:j + 60] for j in range(0, len(r), 60))) for i, r in enumerate(recs[:1500]))
    fa = (fa + ">long\n" + genome[900:21000] + "\n>other\n" + genome[40000:52000] + "\n").encode()
    B = set()
    for seq in zo.sequences("b.fa", baits.encode()):
        B |= set(zo.kmers_list(k, seq, True))
    bs, _ = run_kmerize(nat, k, [(baits.encode(), True)])
    assert sorted(B) == [int(x) for x in bs.fetch()[0]]
    for name, data, is_fa in (("x.fq", fq, False), ("x.fa", fa, True)):
        buf = []
        nrec = 0
        for seq in zo.sequences(name, data):
            xs = zo.kmers_list(k, seq, True)
            if any(x in B for x in xs):
                buf.extend(xs)
            nrec += 1
        buf.sort()
        ek, ec = zo.count_sorted(buf)
        km = nat.Kmerizer(k)
        km.set_baits(bs)
        km.feed(data, is_fa)
        s, nr = km.finish()
        km.close()
        ks, cc = s.fetch()
        assert nr == nrec
        assert len(ek) > 0 and [int(x) for x in ks] == ek and [int(c) for c in cc] == ec, name


def test_concurrent_host_threads(nat):
    """four host threads drive the same GPU at once (every thread has its own library context): kmerize + trim + stats
    + merge + fetch, ten rounds each, every result identical to the single-threaded one"""
    import threading
    from tools import synth
    g = synth.genome(300000, seed=21)
    inputs = [synth.fastq_array(g, 20000, seed=30 + i).reshape(-1).tobytes() for i in range(4)]

    def work(data):
        km = nat.Kmerizer(25)
        km.feed(data, False)
        s, nr = km.finish()
        km.close()
        t = s.trim(2)
        m = nat.merge([s, t, s])
        out = (s.fetch(), t.fetch(), m.fetch(), s.stats()["hist"], nr)
        s.free(); t.free(); m.free()
        return out

    def same(a, b):
        return all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a[:3], b[:3])) and a[3:] == b[3:]

    expect = [work(d) for d in inputs]
    errors = []

    def runner(i):
        try:
            for _ in range(10):
                if not same(work(inputs[i]), expect[i]):
                    errors.append("thread %d: result differs" % i)
                    return
        except Exception as e:   # pragma: no cover
            errors.append("thread %d: %r" % (i, e))

    ths = [threading.Thread(target=runner, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors


def test_kmerize_empty(nat):
    s, nr = run_kmerize(nat, 25, [(b">nothing\nACGT\n", True)])
    assert len(s) == 0 and nr == 1
    assert s.stats()["total"] == 0


# ----------------------------------------------------------------------------- set algebra
def random_set(rng, n, bits=50, maxc=5):
    k = np.unique(rng.integers(0, 2 ** bits, n, dtype=np.uint64))
    c = rng.integers(1, maxc + 1, len(k), dtype=np.uint32)
    return k, c


@pytest.mark.parametrize("sizes", [[0, 0], [1, 0], [10, 10, 10], [5000, 3, 70000], [100000] * 5, [4096, 4096], [30000] * 9])
def test_merge(nat, sizes):
    rng = np.random.default_rng(sum(sizes) + len(sizes))
    sets = [random_set(rng, n, bits=18) for n in sizes]  # 18 bits: plenty of shared k-mers
    hs = [nat.KmerSet.from_arrays(k, c) for k, c in sets]
    m = nat.merge(hs)
    mk, mc = m.fetch()
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    assert np.array_equal(mk, ek) and np.array_equal(mc.astype(np.uint64), ec)
    st = m.stats()
    assert st["hist"] == co.hist(ec)
    assert st["acgt_weighted"] == [int(ec[(ek & np.uint64(3)) == np.uint64(b)].sum()) for b in range(4)]
    assert st["acgt_plain"] == [int(((ek & np.uint64(3)) == np.uint64(b)).sum()) for b in range(4)]


@pytest.mark.parametrize("route,nsets", [("", 5), ("", 2), ("tree", 3), ("sort", 6), ("", 40)])
def test_merge_counts_beyond_u32(nat, monkeypatch, route, nsets):
    """merge.py:56 / :145-146 add Python ints and codec64 carries 60 bits: sums past 2^32-1 come out exact (a WIDE set:
    u64 counts on fetch, true values in hist / acgt and in the encoded stream), whichever merge route meets them"""
    if route:
        monkeypatch.setenv("ZB_MERGE", route)
    rng = np.random.default_rng(900 + nsets)
    pool = np.unique(rng.integers(0, 2 ** 50, 60000, dtype=np.uint64))
    hot = rng.choice(pool, 7, replace=False)          # k-mers every input holds with a huge count
    sets = []
    for i in range(nsets):
        k = np.union1d(np.sort(rng.choice(pool, 20000 + 1000 * (i % 3), replace=False)), hot)
        c = rng.integers(1, 3000, len(k), dtype=np.uint32)
        c[np.isin(k, hot)] = rng.integers(2 ** 31, 2 ** 32 - 1, 7, dtype=np.uint64).astype(np.uint32)
        c[np.isin(k, hot[:1])] = np.uint32(2 ** 32 - 1)
        if i == 0:
            c[np.isin(k, hot[1:2])] = np.uint32(2 ** 32 - 2)      # with the 1s below: a sum of exactly 2^32-1 + ...
        sets.append((k, c))
    hs = [nat.KmerSet.from_arrays(k, c) for k, c in sets]
    m = nat.merge(hs)
    assert m.is_wide()
    mk, mc = m.fetch()
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    assert int(ec.max()) > 2 ** 32 and mc.dtype == np.uint64
    assert np.array_equal(mk, ek) and np.array_equal(mc, ec)
    st = m.stats()
    assert st["hist"] == co.hist(ec)
    assert st["acgt_weighted"] == [int(ec[(ek & np.uint64(3)) == np.uint64(b)].sum()) for b in range(4)]
    assert st["total"] == int(ec.sum())
    kw, cw = m.encode()
    assert np.array_equal(kw, co.encode(ek, True)) and np.array_equal(cw, co.encode(ec, False))
    # subsets of a wide set keep the true counts of what they keep (trim.py:54-62 compares Python ints)
    t = m.trim(3)
    tk, tc = t.fetch()
    xk, xc = co.trim(ek, ec, 3, 0)
    assert t.is_wide() and np.array_equal(tk, xk) and np.array_equal(tc, xc) and t.stats()["hist"] == co.hist(xc)
    t2 = m.trim(1, 1000)                       # the upper cutoff drops every huge count: an ordinary set again
    t2k, t2c = t2.fetch()
    x2k, x2c = co.trim(ek, ec, 1, 1000)
    assert not t2.is_wide() and np.array_equal(t2k, x2k) and np.array_equal(t2c.astype(np.uint64), x2c)
    with pytest.raises(Exception):
        m.trim(1, 2 ** 32 + 5)                 # a cutoff beyond 2^32-1 that would have to drop a saturated entry: not supported
    sl = m.slice(5, len(ek) - 5)
    sk, sc = sl.fetch()
    assert np.array_equal(sk, ek[5:-5]) and np.array_equal(sc.astype(np.uint64), ec[5:-5])
    # the streams read back are the same wide set (files.py:219-227 decodes Python ints)
    back = nat.KmerSet.from_streams(kw, cw)
    bk, bc = back.fetch()
    assert back.is_wide() and np.array_equal(bk, ek) and np.array_equal(bc, ec) and back.stats()["hist"] == st["hist"]
    # and a wide set is an input like any other: four 16-bit planes instead of two
    m2 = nat.merge([back, hs[0], hs[1]])
    k2, c2 = m2.fetch()
    e2k, e2c = co.merge([(ek, ec), (sets[0][0], sets[0][1].astype(np.uint64)), (sets[1][0], sets[1][1].astype(np.uint64))])
    assert m2.is_wide() and np.array_equal(k2, e2k) and np.array_equal(c2, e2c)
    assert m2.stats()["hist"] == co.hist(e2c)
    w2k, w2c = m2.encode()
    assert np.array_equal(w2k, co.encode(e2k, True)) and np.array_equal(w2c, co.encode(e2c, False))
    # sums that stay below 2^32-1 are an ordinary set
    small = nat.merge(hs[:1] + [nat.KmerSet.from_arrays(sets[0][0][:10], np.ones(10, np.uint32))])
    assert not small.is_wide()


def test_merge_nway_equals_pairwise_tree(nat, monkeypatch):
    """>= 3 inputs are merged bucket by bucket of the key space in shared memory (nwaymerge.cu); it must equal the
    reference-shaped pairwise tree (ZB_MERGE=tree) and the oracle, including counts that add up past 2^16 and empty inputs"""
    rng = np.random.default_rng(44)
    pool = np.unique(rng.integers(0, 2 ** 50, 150000, dtype=np.uint64))
    sets = []
    for i in range(9):
        n = [40000, 0, 70000, 1, 65000, 30000, 150000, 5, 90000][i]
        k = np.sort(rng.choice(pool, min(n, len(pool)), replace=False))
        c = rng.integers(1, 200000, len(k), dtype=np.uint32)
        sets.append((k, c))
    hs = [nat.KmerSet.from_arrays(k, c) for k, c in sets]
    a = nat.merge(hs)
    monkeypatch.setenv("ZB_MERGE", "tree")
    b = nat.merge(hs)
    monkeypatch.delenv("ZB_MERGE")
    ak, ac = a.fetch()
    bk, bc = b.fetch()
    assert np.array_equal(ak, bk) and np.array_equal(ac, bc)
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    assert np.array_equal(ak, ek) and np.array_equal(ac.astype(np.uint64), ec)
    # a sum beyond 2^32-1 comes out exact on every path (merge.py:145-146 adds Python ints; test_merge_counts_beyond_u32)
    big = [nat.KmerSet.from_arrays(np.array([7, 9], np.uint64), np.array([2 ** 31, 1], np.uint32)) for _ in range(4)]
    for mode in (None, "sort", "tree"):
        if mode:
            monkeypatch.setenv("ZB_MERGE", mode)
        w = nat.merge(big)
        wk, wc = w.fetch()
        assert w.is_wide() and wk.tolist() == [7, 9] and wc.tolist() == [2 ** 33, 4]


def _merge_all_modes(nat, monkeypatch, sets):
    """zb_merge through its three routes (key-range buckets in shared memory = default, weighted sort of the
    concatenation, pairwise tree) against the oracle"""
    hs = [nat.KmerSet.from_arrays(k, c) for k, c in sets]
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    for mode in (None, "sort", "tree"):
        if mode:
            monkeypatch.setenv("ZB_MERGE", mode)
        else:
            monkeypatch.delenv("ZB_MERGE", raising=False)
        m = nat.merge(hs)
        mk, mc = m.fetch()
        assert np.array_equal(mk, ek) and np.array_equal(mc.astype(np.uint64), ec), mode
        m.free()
    monkeypatch.delenv("ZB_MERGE", raising=False)
    for h in hs:
        h.free()


@pytest.mark.parametrize("nsets,n,bits", [(3, 50000, 50), (64, 20000, 50), (200, 3000, 50), (7, 100000, 64), (5, 2000, 10),
                                          (16, 300, 3), (33, 1, 50), (1024, 500, 40)])
def test_merge_buckets_shapes(nat, monkeypatch, nsets, n, bits):
    rng = np.random.default_rng(nsets * 1000 + bits)
    pool = np.unique(rng.integers(0, 2 ** bits if bits < 64 else 2 ** 64 - 1, 3 * n, dtype=np.uint64, endpoint=False))
    sets = []
    for i in range(nsets):
        k = np.sort(rng.choice(pool, min(n, len(pool)) if i % 7 != 6 else 0, replace=False))
        sets.append((k, rng.integers(1, 1000, len(k), dtype=np.uint32)))
    _merge_all_modes(nat, monkeypatch, sets)


def test_merge_buckets_skewed_keys(nat, monkeypatch):
    """key spaces that the interpolated bucket search and the fixed-prefix buckets do not like: a dense cluster
    (more bucket bits needed), everything in one tiny range (falls back to the sort), identical inputs"""
    rng = np.random.default_rng(77)
    # 90 % of the keys inside 2^-20 of the key space
    dense = np.unique(np.concatenate([rng.integers(2 ** 45, 2 ** 45 + 2 ** 30, 180000, dtype=np.uint64),
                                      rng.integers(0, 2 ** 50, 20000, dtype=np.uint64)]))
    sets = []
    for i in range(12):
        k = np.sort(rng.choice(dense, 60000, replace=False))
        sets.append((k, rng.integers(1, 9, len(k), dtype=np.uint32)))
    _merge_all_modes(nat, monkeypatch, sets)
    # consecutive integers plus one far outlier: almost every key shares all of its top bits
    base = np.arange(1, 200001, dtype=np.uint64)
    sets = [(np.concatenate([base[i::3], np.array([2 ** 63 + 5], np.uint64)]), np.full(len(base[i::3]) + 1, i + 1, np.uint32))
            for i in range(6)]
    _merge_all_modes(nat, monkeypatch, sets)
    # the same set 40 times: every bucket holds each key 40 times
    k, c = random_set(rng, 30000)
    _merge_all_modes(nat, monkeypatch, [(k, c)] * 40)


def test_merge_more_inputs_than_buckets_take(nat, monkeypatch):
    """> 1024 inputs do not fit the bucket kernel's slice table: the sort route takes over, same result"""
    rng = np.random.default_rng(78)
    pool = np.unique(rng.integers(0, 2 ** 50, 5000, dtype=np.uint64))
    sets = []
    for i in range(1100):
        k = np.sort(rng.choice(pool, 40, replace=False))
        sets.append((k, rng.integers(1, 5, len(k), dtype=np.uint32)))
    hs = [nat.KmerSet.from_arrays(k, c) for k, c in sets]
    m = nat.merge(hs)
    mk, mc = m.fetch()
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    assert np.array_equal(mk, ek) and np.array_equal(mc.astype(np.uint64), ec)


def test_stats_large_counts(nat):
    rng = np.random.default_rng(5)
    k, _ = random_set(rng, 50000)
    c = rng.integers(1, 100000, len(k), dtype=np.uint32)
    c[100] = 4000000000
    s = nat.KmerSet.from_arrays(k, c)
    assert s.stats()["hist"] == co.hist(c.astype(np.uint64))


@pytest.mark.parametrize("cmin,cmax", [(1, 0), (2, 0), (2, 3), (6, 0), (3, 3)])
def test_trim(nat, cmin, cmax):
    rng = np.random.default_rng(cmin * 10 + cmax)
    for n in (0, 1, 4095, 4097, 200000):
        k, c = random_set(rng, n)
        s = nat.KmerSet.from_arrays(k, c)
        t = s.trim(cmin, cmax)
        tk, tc = t.fetch()
        ek, ec = co.trim(k, c.astype(np.uint64), cmin, cmax)
        assert np.array_equal(tk, ek) and np.array_equal(tc.astype(np.uint64), ec)


@pytest.mark.parametrize("shift", [0, 2, 26, 40])
def test_project(nat, shift):
    rng = np.random.default_rng(shift)
    for n in (0, 1, 5000, 300000):
        k, c = random_set(rng, n)
        p = nat.KmerSet.from_arrays(k, c).project(shift)
        assert np.array_equal(p.fetch(counts=False), co.project(k, shift))


def test_pairs_abc(nat):
    rng = np.random.default_rng(9)
    base, _ = random_set(rng, 60000, bits=20)
    sets = []
    for i, n in enumerate([0, 1, 50, 4096, 8191, 30000, 60000]):
        pick = np.sort(rng.choice(len(base), min(n, len(base)), replace=False))
        sets.append(base[pick])
    sets.append(np.array([2 ** 64 - 1], np.uint64))
    sets.append(np.array([5, 2 ** 64 - 1], np.uint64))
    hs = [nat.KmerSet.from_arrays(k) for k in sets]
    I, J = np.triu_indices(len(sets), 1)
    abc = nat.pairs_abc(hs, I, J)
    for p in range(len(I)):
        assert tuple(int(v) for v in abc[p]) == co.split(sets[I[p]], sets[J[p]]), (I[p], J[p])


def test_sample_and_restrict_vs_oracle(nat):
    """zb_sample (both float expressions of the reference, evaluated in IEEE double on the device) and zb_restrict"""
    rng = np.random.default_rng(31)
    k = np.unique(rng.integers(0, 2 ** 62, 120000, dtype=np.uint64))
    c = rng.integers(1, 100, len(k), dtype=np.uint32)
    s = nat.KmerSet.from_arrays(k, c)
    M = 0xFFFFFFFFFF
    for (p, seed) in ((0.3, 7), (0.01, 0), (1.0, 2 ** 63 + 5), (0.999999, 1)):
        keep0 = np.array([float(zo.murmer(int(x), seed) & M) / float(M) < p for x in k])
        keep1 = np.array([zo.sub(seed, p, int(x)) for x in k])
        for mode, keep in ((0, keep0), (1, keep1)):
            t = s.sample(p, seed, mode)
            tk, tc = t.fetch()
            assert np.array_equal(tk, k[keep]) and np.array_equal(tc, c[keep]), (p, seed, mode)
            t.free()
    ref = np.unique(np.concatenate([k[::3], rng.integers(0, 2 ** 62, 50000, dtype=np.uint64)]))
    r = nat.KmerSet.from_arrays(ref)
    t = s.restrict(r)
    tk, tc = t.fetch()
    keep = np.isin(k, ref)
    assert np.array_equal(tk, k[keep]) and np.array_equal(tc, c[keep])
    e = nat.KmerSet.from_arrays(np.zeros(0, np.uint64))
    assert len(s.restrict(e)) == 0 and len(e.restrict(s)) == 0 and len(e.sample(0.5)) == 0


def _rand_sets(rng, sizes, bits, share=0.5):
    pool = np.unique(rng.integers(0, 2 ** bits, int(max(sizes) * 2) + 8, dtype=np.uint64))
    out = []
    for n in sizes:
        m = min(n, len(pool))
        own = np.unique(rng.integers(0, 2 ** bits, int(n * (1 - share)) + 1, dtype=np.uint64))
        a = np.unique(np.concatenate([rng.choice(pool, int(m * share), replace=False), own]))
        out.append(a if n else np.zeros(0, np.uint64))
    return out


@pytest.mark.parametrize("sizes,bits", [
    ([3000, 1, 0, 2500, 7], 50),                       # fewer sets than one block, an empty set
    ([20000] * 8 + [500], 50),                         # a block boundary (9 sets)
    ([60000, 100, 40000, 0, 0, 3, 70000, 65000, 12, 9999, 30000, 30001, 64, 63, 65, 50000, 1000], 62),
    ([5000] * 19, 20),                                 # dense small key space (20 bits: many shared keys)
    ([150000] * 5, 64),
    ([3000] * 31 + [0, 50, 4000], 40),                 # 34 sets: two blocks of 32 (a diagonal and a cross tile)
    ([800] * 70, 24),                                  # three blocks, dense key space
])
def test_allpairs_abc(nat, sizes, bits):
    """tiled all-pairs cardinalities against the two-pointer oracle (library/dist.py:241-265) and the
    pair-at-a-time kernel, whole matrix and tile shards"""
    rng = np.random.default_rng(len(sizes) * 1000 + bits)
    arrs = _rand_sets(rng, sizes, bits)
    sets = [nat.KmerSet.from_arrays(a) for a in arrs]
    n = len(sets)
    I, J = np.triu_indices(n, 1)
    abc = nat.allpairs_abc(sets)
    ref = nat.pairs_abc(sets, I, J)
    assert np.array_equal(abc, ref)
    for p in range(0, len(I), max(1, len(I) // 40)):
        assert tuple(int(v) for v in abc[p]) == co.split(arrs[I[p]], arrs[J[p]]), (I[p], J[p])
    # parts over ranges of work units (tile x key-range shard) add up to the whole matrix; a part is zero outside
    # the pairs of its units' tiles, and what it holds is the pair restricted to the units' key shards
    nt = nat.allpairs_tiles(n)
    from zotmer_b200 import multigpu
    assert nt == multigpu.n_tiles(n)
    tot = np.zeros_like(abc)
    cuts = sorted(set([0, 1, nt // 3, (2 * nt) // 3, nt - 1, nt]))
    key_bits = max([int(a.max()).bit_length() for a in arrs if len(a)] + [1])
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = nat.allpairs_abc(sets, a, b)
        tot += part
        mine = np.zeros(len(abc), bool)
        for u in range(a, b):
            for (i, j) in multigpu.tile_pairs(n, u):
                mine[multigpu.pair_index(n, i, j)] = True
        assert not part[~mine].any()
        if b - a == 1:   # one unit: check the key-shard restriction against numpy
            sh = a % multigpu.AP_KS
            assert key_bits >= 3 and multigpu.AP_KS == 8
            sub = [x[(x >> np.uint64(key_bits - 3)) == np.uint64(sh)] if len(x) else x for x in arrs]
            for (i, j) in multigpu.tile_pairs(n, a)[:50]:
                isec = len(np.intersect1d(sub[i], sub[j], assume_unique=True))
                assert tuple(int(v) for v in part[multigpu.pair_index(n, i, j)]) == (isec, len(sub[i]) - isec, len(sub[j]) - isec)
    assert np.array_equal(tot, abc)
    # strided shares (rank r of W takes units r, r + W, ...: some key-range shards of every tile) add up as well
    for world in (2, 3, 8):
        tot = np.zeros_like(abc)
        for r in range(world):
            b, e, st = multigpu.unit_share(n, r, world)
            tot += nat.allpairs_abc(sets, b, e, st)
        assert np.array_equal(tot, abc), world


def test_allpairs_strided_fall_back(nat, monkeypatch):
    """strided shares on the pair-at-a-time route (forced): every pair is computed by exactly one share"""
    rng = np.random.default_rng(15)
    arrs = _rand_sets(rng, [3000] * 70, 40)
    sets = [nat.KmerSet.from_arrays(a) for a in arrs]
    from zotmer_b200 import multigpu
    abc = nat.allpairs_abc(sets)
    monkeypatch.setenv("ZB_ALLPAIRS_PAIRWISE", "1")
    for world in (1, 3, 8):
        tot = np.zeros_like(abc)
        for r in range(world):
            b, e, st = multigpu.unit_share(len(sets), r, world)
            tot += nat.allpairs_abc(sets, b, e, st)
        assert np.array_equal(tot, abc), world


def test_allpairs_all_ones_key(nat):
    """k = 32: the k-mer TTT...T is 2^64-1, the value the kernel's hash tables use as their empty marker"""
    rng = np.random.default_rng(9)
    top = np.uint64(2 ** 64 - 1)
    arrs = []
    for i in range(11):
        a = np.unique(rng.integers(0, 2 ** 63, 3000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 3000, dtype=np.uint64))
        a = a[a != top]
        if i % 3 != 1:
            a = np.concatenate([a, np.array([top], np.uint64)])
        arrs.append(a)
    arrs[4] = np.concatenate([arrs[0][::2], arrs[4]])
    arrs[4] = np.unique(arrs[4])
    sets = [nat.KmerSet.from_arrays(a) for a in arrs]
    I, J = np.triu_indices(len(sets), 1)
    abc = nat.allpairs_abc(sets)
    for p in range(len(I)):
        assert tuple(int(v) for v in abc[p]) == co.split(arrs[I[p]], arrs[J[p]]), (I[p], J[p])


def test_allpairs_skewed_keys_fall_back(nat):
    """all keys share their top bits (one bucket holds everything): the tiled kernel cannot stage that and
    the pair-at-a-time path must take over with the same result"""
    rng = np.random.default_rng(5)
    arrs = [np.unique(rng.integers(0, 40000, 20000, dtype=np.uint64)) | np.uint64(1 << 49) for _ in range(10)]
    arrs[3] = np.concatenate([np.arange(10, dtype=np.uint64), arrs[3]])
    sets = [nat.KmerSet.from_arrays(a) for a in arrs]
    I, J = np.triu_indices(len(sets), 1)
    abc = nat.allpairs_abc(sets)
    for p in range(len(I)):
        assert tuple(int(v) for v in abc[p]) == co.split(arrs[I[p]], arrs[J[p]])


def test_codec_streams(nat):
    rng = np.random.default_rng(12)
    for n in (0, 1, 6, 7, 1000, 100000):
        v = (rng.integers(0, 2 ** 62, n, dtype=np.uint64) >> rng.integers(2, 62, n).astype(np.uint64))
        w = nat.encode_stream(v, False)
        assert np.array_equal(w, co.encode(v, False))
        assert np.array_equal(nat.decode_stream(w, False), v)
        s = np.sort(v)
        w = nat.encode_stream(s, True)
        assert np.array_equal(w, co.encode(s, True))
        assert np.array_equal(nat.decode_stream(w, True), s)
    with pytest.raises(IndexError):
        nat.encode_stream(np.array([2 ** 61], np.uint64))


@pytest.mark.parametrize("n", [2047, 2048, 2049, 4096 + 3, 6 * 2048, 700001])
@pytest.mark.parametrize("kind", ["small", "mixed", "runs", "wide"])
def test_codec_device_vs_oracle(nat, n, kind):
    """the device codec against the C restatement of codec64.py: word-for-word, around the 2048-value encode tiles
    and the 512-word decode tiles, with group patterns that straddle tile edges"""
    rng = np.random.default_rng(n + len(kind))
    if kind == "small":
        v = rng.integers(0, 1024, n, dtype=np.uint64)                      # 6 per word
    elif kind == "mixed":
        v = rng.integers(0, 2 ** 60, n, dtype=np.uint64) >> rng.integers(0, 60, n).astype(np.uint64)
    elif kind == "runs":                                                   # long runs of one width, then a switch
        width = np.repeat(rng.choice([3, 10, 12, 15, 20, 30, 60], n // 97 + 1), 97)[:n].astype(np.uint64)
        v = rng.integers(0, 2 ** 60, n, dtype=np.uint64) >> (np.uint64(60) - width)
    else:
        v = rng.integers(2 ** 59, 2 ** 60, n, dtype=np.uint64)             # 1 per word
    w = nat.encode_stream(v, False)
    assert np.array_equal(w, co.encode(v, False))
    assert np.array_equal(nat.decode_stream(w, False), v)
    s = np.cumsum(v >> np.uint64(24), dtype=np.uint64)                     # ascending "k-mers", gaps of mixed widths
    w = nat.encode_stream(s, True)
    assert np.array_equal(w, co.encode(s, True))
    assert np.array_equal(nat.decode_stream(w, True), s)


def test_codec_set_streams_and_errors(nat):
    rng = np.random.default_rng(77)
    n = 300000
    k = np.unique(rng.integers(0, 2 ** 50, n, dtype=np.uint64))
    c = rng.integers(1, 5000, len(k), dtype=np.uint32)
    c[::1000] = 2 ** 32 - 1
    s = nat.KmerSet.from_arrays(k, c)
    kw, cw = s.encode()
    assert np.array_equal(kw, co.encode(k, True)) and np.array_equal(cw, co.encode(c.astype(np.uint64), False))
    t = nat.KmerSet.from_streams(kw, cw)
    tk, tc = t.fetch()
    assert np.array_equal(tk, k) and np.array_equal(tc, c)
    u = nat.KmerSet.from_streams(kw, None)                                 # counts default to 1 (files.py:152-156)
    uk, uc = u.fetch()
    assert np.array_equal(uk, k) and (uc == 1).all()
    e = nat.KmerSet.from_streams(np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    assert len(e) == 0 and e.encode()[0].size == 0
    with pytest.raises(AssertionError):                                    # streams of different length (files.py:182)
        nat.KmerSet.from_streams(kw, cw[:-1])
    bad = kw.copy()
    bad[len(bad) // 2] = (bad[len(bad) // 2] & ~np.uint64(15)) | np.uint64(9)
    with pytest.raises(AssertionError):                                    # unknown tag (codec64.py:128)
        nat.decode_stream(bad, True)
    big = co.encode(np.array([5, 2 ** 32, 7], np.uint64), False)           # a count beyond u32: a wide set (merge.py:145-146)
    w = nat.KmerSet.from_streams(co.encode(np.array([1, 2, 3], np.uint64), True), big)
    wk, wc = w.fetch()
    assert w.is_wide() and wk.tolist() == [1, 2, 3] and wc.tolist() == [5, 2 ** 32, 7]
    assert w.stats()["hist"] == [(5, 1), (2 ** 32, 1), (7, 1)] and np.array_equal(w.encode()[1], big)
    wide = np.array([1, 2 ** 62], np.uint64)                               # gap > 60 bits
    with pytest.raises(IndexError):
        nat.encode_stream(wide, True)


def test_device_bucketing_matches_host_owner_function(nat):
    """zb_kmerize_take_bucketed_dev (csrc/extract.cu owner_of) vs zotmer_b200.multigpu.owner_of"""
    import torch
    from zotmer_b200 import multigpu
    rng = np.random.default_rng(21)
    fq = make_fastq(rng, rnd_dna(rng, 30000), 3000, 100)
    for world in (1, 2, 3, 8):
        km = nat.Kmerizer(25)
        km.feed(fq, False)
        n = km.pending()
        buf = torch.empty(n + 16, dtype=torch.int64, device="cuda:0")
        counts = km.take_bucketed_dev(world, buf.data_ptr())
        torch.cuda.synchronize()
        keys = buf[:n].cpu().numpy().view(np.uint64)
        assert sum(counts) == n
        exp_owner = np.repeat(np.arange(world), counts)
        assert np.array_equal(multigpu.owner_of(keys, world), exp_owner)
        # hand everything back (as if received from peers) and finish: result == single-GPU oracle
        km.add_canonical_dev(buf.data_ptr(), n)
        s, nr = km.finish()
        km.close()
        ks, cc = s.fetch()
        ek, ec, _, _ = co.kmerize(25, [(fq, False)])
        assert np.array_equal(ks, ek) and np.array_equal(cc, ec)
