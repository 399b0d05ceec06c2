"""
Parity at BASELINE.json's full sizes (run on the B200 box: pytest -m gpu).  config[0] (5 Mbp FASTA) and config[1]
(1,000,000 x 150 bp reads, 150 Mbases) go through the C ABI and are compared bit for bit with the C restatement of
the reference (oracle/zot_oracle.c, ~40 s on the host for config[1]) AND through size-independent properties:
strict order, sum of counts = 2 x valid windows counted independently from the text, count(x) == count(rc(x)),
sum(hist c * freq) = sum of counts, encode -> decode round trip, merge(s, s) = doubled counts, all sort+count
routes identical, all-pairs cardinalities consistent with the two-pointer oracle.  config[2] (merge of 64 sets, 637 M
entries) at full size: three independent merge routes bit-identical + count sum + a key slice against numpy;
config[3] on a bounded instance (70 sets = every tile kind): all-pairs kernel == pair-at-a-time kernel == oracle.
"""
import numpy as np
import pytest

from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
K = 25


@pytest.fixture(scope="module")
def nat():
    from zotmer_b200 import _native
    assert _native.device_count() >= 1, "no CUDA device"
    return _native


def rc_np(x, k):
    x = ~x
    m = np.uint64(0x3333333333333333)
    x = ((x >> np.uint64(2)) & m) | ((x & m) << np.uint64(2))
    m = np.uint64(0x0F0F0F0F0F0F0F0F)
    x = ((x >> np.uint64(4)) & m) | ((x & m) << np.uint64(4))
    x = x.byteswap()
    return x >> np.uint64(64 - 2 * k)


def valid_windows(reads_2d, k):
    """windows of k consecutive ACGT bytes per read, counted with numpy from the read matrix [nreads, L]"""
    ok = np.isin(reads_2d, np.frombuffer(b"ACGTacgtUu", np.uint8))
    bad = (~ok).astype(np.int32)
    cs = np.concatenate([np.zeros((len(bad), 1), np.int32), np.cumsum(bad, axis=1)], axis=1)
    return int(((cs[:, k:] - cs[:, :-k]) == 0).sum())


@pytest.fixture(scope="module")
def config1(nat):
    from tools import synth
    g = synth.genome(5000000)
    rec = synth.fastq_array(g, 1000000)            # [1,000,000, 315]
    fq = rec.reshape(-1).tobytes()
    km = nat.Kmerizer(K)
    km.feed(fq, False)
    s, nr = km.finish()
    km.close()
    assert nr == 1000000
    return {"rec": rec, "fq": fq, "set": s}


def test_config1_properties(nat, config1):
    s = config1["set"]
    ks, cs = s.fetch()
    assert len(ks) > 30000000
    assert np.all(ks[1:] > ks[:-1]), "k-mers not strictly ascending"
    windows = valid_windows(config1["rec"][:, 11:161], K)
    assert int(cs.astype(np.uint64).sum()) == 2 * windows, "sum of counts != 2 x valid windows"
    r = rc_np(ks, K)
    pos = np.searchsorted(ks, r)
    assert np.array_equal(ks[pos], r), "set not closed under reverse complement"
    assert np.array_equal(cs[pos], cs), "count(x) != count(rc(x))"
    st = s.stats()
    assert sum(int(v) * int(f) for v, f in st["hist"]) == int(cs.astype(np.uint64).sum())
    assert sum(int(f) for v, f in st["hist"]) == len(ks)
    assert sum(st["acgt_weighted"]) == int(cs.astype(np.uint64).sum())
    bc = np.bincount((ks & np.uint64(3)).astype(np.int64), weights=cs.astype(np.float64), minlength=4)
    assert [int(v) for v in bc] == [int(v) for v in st["acgt_weighted"]]


def test_config1_all_sort_count_paths_agree(nat, config1, monkeypatch):
    ks, cs = config1["set"].fetch()
    for route in ("1", "2"):   # classic: full LSD sort + reduce-by-key; segment route (default = bucket route)
        monkeypatch.setenv("ZB_SORT_COUNT", route)
        km = nat.Kmerizer(K)
        km.feed(config1["fq"], False)
        s2, _ = km.finish()
        km.close()
        k2, c2 = s2.fetch()
        s2.free()
        assert np.array_equal(ks, k2) and np.array_equal(cs, c2), route
    monkeypatch.delenv("ZB_SORT_COUNT")
    nat.Kmerizer(K).close()                        # re-reads the environment: back to the default path


def test_config1_vs_c_oracle_trim_codec_merge(nat, config1):
    s = config1["set"]
    ks, cs = s.fetch()
    ek, ec, eacgt, enr = co.kmerize(K, [(config1["fq"], False)])
    assert enr == 1000000
    assert np.array_equal(ks, ek) and np.array_equal(cs, ec), "config[1] kmerize+count differs from the oracle"
    # zot trim -c 2
    t = s.trim(2)
    tk, tc = t.fetch()
    keep = ec >= 2
    assert np.array_equal(tk, ek[keep]) and np.array_equal(tc, ec[keep])
    # file streams: word for word, and back
    kw, cw = t.encode()
    assert np.array_equal(kw, co.encode(tk, True)) and np.array_equal(cw, co.encode(tc.astype(np.uint64), False))
    back = nat.KmerSet.from_streams(kw, cw)
    bk, bc = back.fetch()
    assert np.array_equal(bk, tk) and np.array_equal(bc, tc)
    # merge(s, t): counts add where both have the k-mer
    m = nat.merge([s, t])
    mk, mc = m.fetch()
    assert np.array_equal(mk, ek)
    exp = ec.astype(np.uint64) + np.where(keep, ec, 0).astype(np.uint64)
    assert np.array_equal(mc.astype(np.uint64), exp)
    # pair cardinalities of (s, t, m): t is a subset of s = the k-mers of m
    abc = nat.allpairs_abc([s, t, m])
    assert tuple(int(v) for v in abc[0]) == (len(tk), len(ek) - len(tk), 0)
    assert tuple(int(v) for v in abc[1]) == (len(ek), 0, 0)
    assert tuple(int(v) for v in abc[2]) == (len(tk), 0, len(ek) - len(tk))
    for x in (t, back, m):
        x.free()


def test_config0_fasta_vs_c_oracle(nat):
    from tools import synth
    g = synth.genome(5000000)
    fa = synth.fasta_bytes(g)
    km = nat.Kmerizer(K)
    km.feed(fa, True)
    s, nr = km.finish()
    km.close()
    ks, cs = s.fetch()
    ek, ec, eacgt, enr = co.kmerize(K, [(fa, True)])
    assert nr == enr == 1
    assert np.array_equal(ks, ek) and np.array_equal(cs, ec)
    assert s.stats()["acgt_weighted"] == eacgt
    assert int(cs.astype(np.uint64).sum()) == 2 * (5000000 - K + 1)
    s.free()


def test_config2_merge_64_sets(nat, monkeypatch):
    """BASELINE.json config[2] at full size: zb_merge of 64 synthetic bacterial k-mer sets (k=25, ~10 M (k-mer, count)
    entries each, 637 M in all).  Three independent routes -- key-range buckets in shared memory (default), weighted
    sort + count of the concatenation, pairwise merge tree -- must agree bit for bit; the sum of counts is preserved,
    the result is strictly ascending, and a slice of the key space is checked against numpy."""
    from tools import synth
    g = synth.genome(5000000)
    sets = []
    for i in range(64):
        h = synth.mutate(g, 0.0005 + 0.0195 * (i % 16) / 16 + 0.02 * (i // 16), 100 + i)   # 4 clades, 0.05 % .. 2 % within
        km = nat.Kmerizer(K)
        km.feed(synth.fasta_bytes(h), True)
        s, _ = km.finish()
        km.close()
        sets.append(s)
    total = sum(int(s.stats()["total"]) for s in sets)
    results = {}
    for mode in (None, "sort", "tree"):
        if mode:
            monkeypatch.setenv("ZB_MERGE", mode)
        else:
            monkeypatch.delenv("ZB_MERGE", raising=False)
        m = nat.merge(sets)
        results[mode] = m.fetch()
        assert int(m.stats()["total"]) == total, mode
        m.free()
        nat.release_cache()
    monkeypatch.delenv("ZB_MERGE", raising=False)
    mk, mc = results[None]
    assert len(mk) > 300000000 and np.all(mk[1:] > mk[:-1])
    for mode in ("sort", "tree"):
        assert np.array_equal(results[mode][0], mk) and np.array_equal(results[mode][1], mc), mode
    # one 1/4096 slice of the key space against numpy
    lo, hi = np.uint64(1234) << np.uint64(38), np.uint64(1235) << np.uint64(38)
    parts = []
    for s in sets:
        k, c = s.fetch()
        a, b = np.searchsorted(k, lo), np.searchsorted(k, hi)
        parts.append((k[a:b], c[a:b].astype(np.uint64)))
    allk = np.concatenate([p[0] for p in parts])
    allc = np.concatenate([p[1] for p in parts])
    u, inv = np.unique(allk, return_inverse=True)
    sums = np.zeros(len(u), np.uint64)
    np.add.at(sums, inv, allc)
    a, b = np.searchsorted(mk, lo), np.searchsorted(mk, hi)
    assert len(u) > 10000 and np.array_equal(mk[a:b], u) and np.array_equal(mc[a:b].astype(np.uint64), sums)
    for s in sets:
        s.free()


def test_config3_allpairs_bounded(nat):
    """BASELINE.json config[3] on a bounded instance that exercises every tile kind: 70 synthetic genomes (0.5 Mbp, 5
    clades) = three blocks of 32 sets.  The all-pairs kernel (per-key set masks) against the pair-at-a-time merge-path
    kernel for ALL 2,415 pairs, against the two-pointer oracle for a sample, and as the sum of its work units."""
    from tools import synth
    base = [synth.genome(500000, seed=700 + c) for c in range(5)]
    arrs, sets = [], []
    for i in range(70):
        g = synth.mutate(base[i % 5], 0.001 + 0.002 * (i // 5), 800 + i)
        km = nat.Kmerizer(K)
        km.feed(synth.fasta_bytes(g), True)
        s, _ = km.finish()
        km.close()
        p = s.project(0)
        s.free()
        sets.append(p)
        arrs.append(p.fetch(counts=False))
    n = len(sets)
    I, J = np.triu_indices(n, 1)
    abc = nat.allpairs_abc(sets)
    ref = nat.pairs_abc(sets, I, J)
    assert np.array_equal(abc, ref)
    assert abc[:, 0].max() > 500000            # related genomes share most of their k-mers
    for p in range(0, len(I), 97):
        assert tuple(int(v) for v in abc[p]) == co.split(arrs[I[p]], arrs[J[p]])
    nt = nat.allpairs_tiles(n)
    tot = np.zeros_like(abc)
    for a, b in ((0, 5), (5, nt // 2), (nt // 2, nt)):
        tot += nat.allpairs_abc(sets, a, b)
    assert np.array_equal(tot, abc)
    # 40 of them: two blocks = the single folded tile
    abc40 = nat.allpairs_abc(sets[:40])
    I4, J4 = np.triu_indices(40, 1)
    assert np.array_equal(abc40, nat.pairs_abc(sets[:40], I4, J4))
    for s in sets:
        s.free()
