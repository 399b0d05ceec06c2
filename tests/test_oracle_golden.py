"""
Pins the oracle (oracle/zot_oracle.py and oracle/zot_oracle.c) against the fixtures under
tests/golden/data, which are outputs of the reference's own code (tests/golden/make_golden.py).
CPU only.
"""
import json
import os

import numpy as np
import pytest

from oracle import zot_oracle as zo
from oracle import c_oracle as co

from conftest import GOLDEN


def g(name):
    return os.path.join(GOLDEN, name)


def rd(name):
    with open(g(name), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def kat():
    with open(g("kat.json")) as f:
        return json.load(f)


# ------------------------------------------------------------------ function-level KATs

def test_kmers_list(kat):
    for e in kat["kmersList"]:
        seq = e["seq"].encode("latin-1")
        assert [str(x) for x in zo.kmers_list(e["k"], seq, e["both"])] == e["out"], (e["k"], e["seq"])


def test_kmers_list_c(kat):
    for e in kat["kmersList"]:
        if not e["both"]:
            continue
        fq = b"@r\n" + e["seq"].encode("latin-1") + b"\n+\n\n"
        if any(c in e["seq"] for c in "\n"):
            continue
        keys, acgt, nr = co.extract(e["k"], fq, False)
        # a FASTQ line is strip()ped: only compare when stripping is a no-op
        if e["seq"].encode("latin-1").strip(zo.PY2_SPACE) != e["seq"].encode("latin-1"):
            continue
        assert [str(int(x)) for x in keys] == e["out"], (e["k"], e["seq"])
        assert nr == 1


def test_rc_rev_murmer_render(kat):
    for e in kat["rc"]:
        assert str(zo.rc(e["k"], int(e["x"]))) == e["out"]
    for e in kat["rev"]:
        assert str(zo.rev(int(e["x"]))) == e["out"]
    for e in kat["murmer"]:
        assert str(zo.murmer(int(e["x"]), e["s"])) == e["out"]
    for e in kat["render"]:
        assert zo.render(e["k"], int(e["x"])) == e["out"]


def test_codec64(kat):
    assert [list(t) for t in zo.LOOKUP[:12]] == kat["codec64_lookup"]
    for e in kat["codec64"]:
        xs = [int(x) for x in e["xs"]]
        ws = [int(w) for w in e["ws"]]
        assert zo.encode(xs) == ws
        assert zo.decode(ws) == xs
        assert [int(w) for w in co.encode(np.array(xs, np.uint64))] == ws
        assert [int(x) for x in co.decode(np.array(ws, np.uint64))] == xs
    assert kat["codec64_overflow_first"] == "IndexError"
    with pytest.raises(IndexError):
        zo.encode([2 ** 61])
    with pytest.raises(IndexError):
        co.encode(np.array([2 ** 61], np.uint64))


def test_split_and_measures(kat):
    for e in kat["split"]:
        assert list(zo.split(e["xs"], e["ys"])) == e["split"]
        assert list(co.split(e["xs"], e["ys"])) == e["split"]
        if "jaccard" in e:
            j = zo.jaccard(e["xs"], e["ys"])
            assert [j[0], j[1], float(j[2]).hex()] == e["jaccard"]
            a, b, c = e["split"]
            for nm, key in (("brayCurtis", "bray.curtis.qual"), ("chord", "chord.qual"), ("hellinger", "hellinger.qual"),
                            ("jaccard", "jaccard.qual"), ("kulczynski", "kulczynski.qual"), ("ochiai", "ochiai.qual"),
                            ("sorensen", "sorensen.qual"), ("whittaker", "whittaker.qual")):
                assert float(zo.QUAL[key](a, b, c)).hex() == e["m_" + nm], nm


# ------------------------------------------------------------------ command-level goldens

KMERIZE = [(5, "kat6.k5", ["kat6.fa"]), (5, "g1.k5", ["g1.fa"]), (16, "g1.k16", ["g1.fa"]), (25, "g1.k25", ["g1.fa"]),
           (30, "g1.k30", ["g1.fa"]), (31, "g1.k31", ["g1.fa"]), (32, "g1.k32", ["g1.fa"]),
           (8, "r1.k8", ["r1.fq"]), (25, "r1.k25", ["r1.fq"]), (31, "r1.k31", ["r1.fq"]),
           (21, "r2.k21", ["r2.fq"]), (25, "mix.k25", ["s0.fa", "r1.fq", "s1.fa"]),
           (25, "s3.k25", ["s3.fa"]), (16, "s4.k16", ["s4.fa"])]


@pytest.mark.parametrize("k,out,ins", KMERIZE)
def test_cmd_kmerize(tmp_path, k, out, ins):
    o = str(tmp_path / out)
    zo.cmd_kmerize(k, o, [g(i) for i in ins])
    assert open(o, "rb").read() == rd(out)


@pytest.mark.parametrize("k,out,ins", KMERIZE)
def test_c_kmerize_matches(k, out, ins):
    z = zo.CasketReader(g(out))
    xs, cs = zo.read_kmers_and_counts(z)
    ks, cc, acgt, nr = co.kmerize(k, [(rd(i), zo.is_fasta(i)) for i in ins])
    assert [int(x) for x in ks] == xs
    assert [int(c) for c in cc] == cs
    assert nr == z.meta["reads"]
    n = float(sum(acgt))
    assert [a / n for a in acgt] == z.meta["acgt"]
    assert [(str(v), f) for v, f in co.hist(cc.astype(np.uint64))] == list(z.meta["hist"].items())


def test_kmerize_empty_divides_by_zero(tmp_path, kat):
    assert kat["kmerize_empty"] == "ZeroDivisionError"
    with pytest.raises(ZeroDivisionError):
        zo.cmd_kmerize(25, str(tmp_path / "e.k25"), [g("empty.fa")])


@pytest.mark.parametrize("out,ins", [("m3.k25", ["s0.k25", "s1.k25", "s2.k25"]),
                                     ("m4.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25"]),
                                     ("m5.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25", "s4.k25"]),
                                     ("m5dup.k25", ["s0.k25", "s0.k25", "g1.k25", "r1.k25", "s0.k25"]),
                                     ("m2.k25", ["s0.k25", "s1.k25"])])
def test_cmd_merge(tmp_path, out, ins):
    o = str(tmp_path / out)
    zo.cmd_merge(o, [g(i) for i in ins])
    assert open(o, "rb").read() == rd(out)
    sets = []
    for i in ins:
        xs, cs = zo.read_kmers_and_counts(zo.CasketReader(g(i)))
        sets.append((np.array(xs, np.uint64), np.array(cs, np.uint64)))
    mk, mc = co.merge(sets)
    xs, cs = zo.read_kmers_and_counts(zo.CasketReader(g(out), with_meta=True))
    assert [int(x) for x in mk] == xs and [int(c) for c in mc] == cs


def test_merge_behavioural(tmp_path, kat, capsys):
    assert kat["merge_one_input"] == "ZeroDivisionError"
    with pytest.raises(ZeroDivisionError):
        zo.cmd_merge(str(tmp_path / "m1"), [g("s0.k25")])
    with pytest.raises(SystemExit) as ei:
        zo.cmd_merge(str(tmp_path / "mb"), [g("s0.k25"), g("s1.k25"), g("s2.k16")])
    assert ei.value.code == 1
    assert capsys.readouterr().err == kat["merge_mismatched_K"]["stderr"]


@pytest.mark.parametrize("out,inp,c,C", [("r1_c2.k25", "r1.k25", 2, 0), ("r1_c2_C3.k25", "r1.k25", 2, 3),
                                         ("r2_c3.k21", "r2.k21", 3, 0), ("m5_c2.k25", "m5.k25", 2, 0),
                                         ("r1_c1000.k25", "r1.k25", 1000, 0)])
def test_cmd_trim(tmp_path, out, inp, c, C):
    o = str(tmp_path / out)
    zo.cmd_trim(o, g(inp), c, C)
    assert open(o, "rb").read() == rd(out)
    xs, cs = zo.read_kmers_and_counts(zo.CasketReader(g(inp)))
    ox, oc = co.trim(np.array(xs, np.uint64), np.array(cs, np.uint64), c, C)
    gx, gc = zo.read_kmers_and_counts(zo.CasketReader(g(out)))
    assert [int(x) for x in ox] == gx and [int(v) for v in oc] == gc


def test_trim_c0_typeerror(tmp_path, kat):
    assert kat["trim_c0"] == "TypeError"
    with pytest.raises(TypeError):
        zo.cmd_trim(str(tmp_path / "t"), g("r1.k25"), 0)


def test_cmd_hist_dump(in_golden_dir):
    assert zo.cmd_hist(["g1.k25", "r1.k25", "r1_c2.k25", "m5.k25", "m2.k25"]) == rd("hist.txt").decode()
    assert zo.cmd_dump("kat6.k5") == rd("dump_kat6.txt").decode()
    assert zo.cmd_dump("r1_c2.k25") == rd("dump_r1_c2.txt").decode()


def test_cmd_dist(in_golden_dir):
    sets = ["s%d.k25" % i for i in range(5)]
    assert zo.cmd_dist(list(zo.QUAL), 25, sets) == rd("dist_qual_25.txt").decode()
    assert zo.cmd_dist(list(zo.QUAL), 12, sets) == rd("dist_qual_12_of_25.txt").decode()
    assert zo.cmd_dist(["jaccard.qual", "kulczynski.qual"], 25, sets[:3] + ["g1.k25"]) == rd("dist_two_25.txt").decode()
    # C restatement of prep()+split()
    xs = np.array(zo.read_kmers(zo.CasketReader("s0.k25")), np.uint64)
    ys = np.array(zo.read_kmers(zo.CasketReader("s1.k25")), np.uint64)
    assert co.split(co.project(xs, 26), co.project(ys, 26)) == zo.split(zo.prep(12, "s0.k25"), zo.prep(12, "s1.k25"))


def test_cmd_jaccard(in_golden_dir):
    sets = ["s%d.k25" % i for i in range(5)]
    assert zo.cmd_jaccard(sets, False) == rd("jaccard_first.txt").decode()
    assert zo.cmd_jaccard(sets, True) == rd("jaccard_all.txt").decode()


# ----------------------------------------------------------------------------- SURVEY.md 8f row 2
def test_sub_kat(kat):
    for s_, p_, x_, want in kat["sub"]:
        assert zo.sub(s_, p_, x_) == want


@pytest.mark.parametrize("out,inp,p,S", [("r1_P03_S7.k25", "r1.k25", 0.3, 7), ("g1_Pdef.k25", "g1.k25", 0.01, 0),
                                         ("m5_P05_D.k25", "m5.k25", 0.5, 0), ("r1_P1.k25", "r1.k25", 1.0, 123456789)])
def test_cmd_sample(tmp_path, out, inp, p, S):
    o = str(tmp_path / out)
    zo.cmd_sample(o, g(inp), p, S)
    assert open(o, "rb").read() == rd(out)


@pytest.mark.parametrize("ref,out,inp", [("r1_c2.k25", "proj_r1c2_r1.k25", "r1.k25"), ("s0.k25", "proj_s0_s1.k25", "s1.k25"),
                                         ("s1.k25", "proj_s1_m5.k25", "m5.k25")])
def test_cmd_project(tmp_path, ref, out, inp):
    o = str(tmp_path / out)
    zo.cmd_project(g(ref), o, g(inp))
    assert open(o, "rb").read() == rd(out)


@pytest.mark.parametrize("k,out,ins,d,S", [(25, "r1_D03_S5.k25", ["r1.fq"], 0.3, 5), (25, "g1_D05.k25", ["g1.fa"], 0.5, 0),
                                           (16, "g1_D2.k16", ["g1.fa"], 2.0, 9)])
def test_cmd_kmerize_D(tmp_path, k, out, ins, d, S):
    o = str(tmp_path / out)
    zo.cmd_kmerize_D(k, o, [g(i) for i in ins], d, S)
    assert open(o, "rb").read() == rd(out)


@pytest.mark.parametrize("k,out,ins,baits", [(25, "r1_C.k25", ["r1.fq"], "baits.fa"), (25, "g1_C.k25", ["g1.fa"], "baits.fa"),
                                             (16, "mix_C.k16", ["r1.fq", "g1.fa"], "baits.fa"),
                                             (25, "r1_Cself.k25", ["r1.fq"], "g1.fa")])
def test_cmd_kmerize_C(tmp_path, k, out, ins, baits):
    """capture mode (kmerize.py:478-483, :507-517) against the files the reference wrote"""
    o = str(tmp_path / out)
    zo.cmd_kmerize_C(k, o, [g(i) for i in ins], g(baits))
    assert open(o, "rb").read() == rd(out)
