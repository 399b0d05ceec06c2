"""
Command-level parity (run on the B200 box: pytest -m gpu): every `zot` sub-command of the hot path is
run through zotmer_b200.cli exactly as a user would, and its output FILE BYTES / stdout are compared
with the fixtures the reference itself produced (tests/golden/make_golden.py).
"""
import os
import shutil

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def rd(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture()
def zot(in_golden_dir):
    from zotmer_b200 import cli
    from zotmer_b200 import _native
    assert _native.device_count() >= 1

    def run(*argv):
        cli.main([str(a) for a in argv])
    return run


KMERIZE = [(5, "kat6.k5", ["kat6.fa"]), (5, "g1.k5", ["g1.fa"]), (16, "g1.k16", ["g1.fa"]), (25, "g1.k25", ["g1.fa"]),
           (30, "g1.k30", ["g1.fa"]), (31, "g1.k31", ["g1.fa"]), (32, "g1.k32", ["g1.fa"]),
           (8, "r1.k8", ["r1.fq"]), (25, "r1.k25", ["r1.fq"]), (31, "r1.k31", ["r1.fq"]),
           (21, "r2.k21", ["r2.fq"]), (25, "mix.k25", ["s0.fa", "r1.fq", "s1.fa"]),
           (25, "s2.k25", ["s2.fa"]), (16, "s4.k16", ["s4.fa"])]


@pytest.mark.parametrize("k,out,ins", KMERIZE)
def test_kmerize_file_bytes(zot, tmp_path, k, out, ins):
    o = tmp_path / out
    zot("kmerize", k, o, *ins)
    assert o.read_bytes() == rd(out)


def test_kmerize_m_option_and_gz(zot, tmp_path):
    import gzip
    gz = tmp_path / "r1.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(rd("r1.fq"))
    o = tmp_path / "o.k25"
    zot("kmerize", "-m", "1", 25, o, gz)
    assert o.read_bytes() == rd("r1.k25")


def test_kmerize_empty_raises_like_reference(zot, tmp_path):
    from zotmer_b200.commands import kmerize
    with pytest.raises(ZeroDivisionError):
        kmerize.main(["kmerize", "25", str(tmp_path / "e.k25"), "empty.fa"])


@pytest.mark.parametrize("out,ins", [("m3.k25", ["s0.k25", "s1.k25", "s2.k25"]),
                                     ("m4.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25"]),
                                     ("m5.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25", "s4.k25"]),
                                     ("m5dup.k25", ["s0.k25", "s0.k25", "g1.k25", "r1.k25", "s0.k25"]),
                                     ("m2.k25", ["s0.k25", "s1.k25"])])
def test_merge_file_bytes(zot, tmp_path, out, ins):
    o = tmp_path / out
    zot("merge", o, *ins)
    assert o.read_bytes() == rd(out)


def test_merge_counts_beyond_u32_file(zot, tmp_path):
    """`zot merge` of sets whose counts add up past 2^32-1 (the reference adds Python ints, merge.py:145-146, and codec64
    writes them): the file holds the exact sums, hist / acgt are computed from them"""
    import numpy as np
    from oracle import c_oracle as co
    from oracle import zot_oracle as zo
    from zotmer_b200.library.kmers import kmers
    from zotmer_b200.library.files import writeKmersAndCounts2, readWords
    rng = np.random.default_rng(5)
    pool = np.unique(rng.integers(0, 2 ** 50, 3000, dtype=np.uint64))
    sets = []
    for i in range(3):
        k = np.sort(rng.choice(pool, 2000, replace=False))
        k = np.union1d(k, pool[:3])
        c = rng.integers(1, 50, len(k), dtype=np.uint32)
        c[np.isin(k, pool[:3])] = np.array([2 ** 32 - 1, 3000000000, 2 ** 31], dtype=np.uint32)
        sets.append((k, c))
        with kmers(str(tmp_path / ("w%d.k25" % i)), "w") as z:
            writeKmersAndCounts2(z, k, c)
            z.meta["K"] = 25
            z.meta["kmers"] = "kmers"
            z.meta["counts"] = "counts"
    out = tmp_path / "wide.k25"
    zot("merge", out, *[tmp_path / ("w%d.k25" % i) for i in range(3)])
    ek, ec = co.merge([(k, c.astype(np.uint64)) for k, c in sets])
    assert int(ec.max()) == 3 * (2 ** 32 - 1)
    with kmers(str(out), "r") as z:
        kw = np.array(readWords(z.open("kmers")), dtype=np.uint64)
        cw = np.array(readWords(z.open("counts")), dtype=np.uint64)
        meta = dict(z.meta)
    assert np.array_equal(kw, co.encode(ek, True)) and np.array_equal(cw, co.encode(ec, False))
    assert zo.decode([int(w) for w in cw]) == [int(v) for v in ec]
    want_hist = {}
    for v in ec:
        want_hist[int(v)] = want_hist.get(int(v), 0) + 1
    assert {int(a): b for a, b in meta["hist"].items()} == want_hist
    tot = float(int(ec.sum()))
    assert meta["acgt"] == [int(ec[(ek & np.uint64(3)) == np.uint64(b)].sum()) / tot for b in range(4)]
    # the file is an input like any other: merged once more with its own inputs
    out2 = tmp_path / "wide2.k25"
    zot("merge", out2, out, tmp_path / "w0.k25", tmp_path / "w1.k25")
    e2k, e2c = co.merge([(ek, ec)] + [(k, c.astype(np.uint64)) for k, c in sets[:2]])
    with kmers(str(out2), "r") as z:
        kw2 = np.array(readWords(z.open("kmers")), dtype=np.uint64)
        cw2 = np.array(readWords(z.open("counts")), dtype=np.uint64)
    assert np.array_equal(kw2, co.encode(e2k, True)) and np.array_equal(cw2, co.encode(e2c, False))


def test_merge_behavioural(zot, tmp_path, capsys):
    from zotmer_b200.commands import merge
    with pytest.raises(ZeroDivisionError):
        merge.main(["merge", str(tmp_path / "m1"), "s0.k25"])
    with pytest.raises(SystemExit) as ei:
        merge.main(["merge", str(tmp_path / "mb"), "s0.k25", "s1.k25", "s2.k16"])
    assert ei.value.code == 1
    assert capsys.readouterr().err == "mismatched K\n"


@pytest.mark.parametrize("out,inp,args", [("r1_c2.k25", "r1.k25", ["-c", "2"]), ("r2_c3.k21", "r2.k21", ["-c", "3"]),
                                          ("m5_c2.k25", "m5.k25", ["-c", "2"]), ("r1_c1000.k25", "r1.k25", ["-c", "1000"])])
def test_trim_file_bytes(zot, tmp_path, out, inp, args):
    o = tmp_path / out
    zot("trim", *args, o, inp)
    assert o.read_bytes() == rd(out)


def test_trim_upper_cutoff_and_c0(zot, tmp_path):
    from zotmer_b200.commands import trim
    from zotmer_b200 import docopt_mini
    o = tmp_path / "t.k25"
    # -C cannot be given on the command line (it is not in the usage pattern); exercise it like the golden did
    real = docopt_mini.docopt
    try:
        docopt_mini.docopt = lambda doc, argv=None, **kw: {"<output>": str(o), "<input>": "r1.k25", "-c": "2", "-C": "3"}
        trim.main(["trim"])
    finally:
        docopt_mini.docopt = real
    assert o.read_bytes() == rd("r1_c2_C3.k25")
    with pytest.raises(TypeError):
        trim.main(["trim", str(o), "r1.k25"])


def test_hist_dump_info(zot, capsys):
    zot("hist", "g1.k25", "r1.k25", "r1_c2.k25", "m5.k25", "m2.k25")
    assert capsys.readouterr().out == rd("hist.txt").decode()
    zot("dump", "kat6.k5")
    assert capsys.readouterr().out == rd("dump_kat6.txt").decode()
    zot("dump", "r1_c2.k25")
    assert capsys.readouterr().out == rd("dump_r1_c2.txt").decode()
    zot("info", "kat6.k5", "m3.k25")
    assert capsys.readouterr().out == rd("info.txt").decode()


def test_dist(zot, capsys):
    sets = ["s%d.k25" % i for i in range(5)]
    zot("dist", "-M", "*.qual", 25, *sets)
    assert capsys.readouterr().out == rd("dist_qual_25.txt").decode()
    zot("dist", "-M", "jaccard.qual", "-M", "kulczynski.qual", 25, "s0.k25", "s1.k25", "s2.k25", "g1.k25")
    assert capsys.readouterr().out == rd("dist_two_25.txt").decode()
    zot("dist", "-M", "*.qual", 12, *sets)
    assert capsys.readouterr().out == rd("dist_qual_12_of_25.txt").decode()
    zot("dist", "-M", "list", 25, *sets)
    assert capsys.readouterr().out == rd("dist_list.txt").decode()
    zot("dist", "-M", "nosuch", 25, *sets)
    cap = capsys.readouterr()
    assert cap.out == rd("dist_bad.txt").decode() and cap.err == rd("dist_bad.txt.stderr").decode()
    zot("dist", 25, *sets)
    assert capsys.readouterr().out == ""


def test_dist_behavioural(zot):
    from zotmer_b200.commands import dist
    from zotmer_b200.library.exceptions import MismatchedK
    with pytest.raises(TypeError):
        dist.main(["dist", "-M", "jaccard.ab", "5", "kat6.k5", "g1.k5"])
    with pytest.raises(MismatchedK):
        dist.main(["dist", "-M", "jaccard.qual", "25", "s0.k16", "s1.k16"])


def test_jaccard(zot, capsys):
    sets = ["s%d.k25" % i for i in range(5)]
    zot("jaccard", *sets)
    assert capsys.readouterr().out == rd("jaccard_first.txt").decode()
    zot("jaccard", "-a", *sets)
    assert capsys.readouterr().out == rd("jaccard_all.txt").decode()
    zot("jaccard", "-a", "-p", "0.9", *sets[:3])
    assert capsys.readouterr().out == rd("jaccard_p.txt").decode()
    zot("jaccard", "-ap", "0.9", *sets[:3])
    assert capsys.readouterr().out == rd("jaccard_p.txt").decode()
    zot("jaccard", "-a", "j.fa")
    assert capsys.readouterr().out == rd("jaccard_fasta.txt").decode()


def test_jaccard_mismatched_K(zot, capsys):
    from zotmer_b200.commands import jaccard
    with pytest.raises(SystemExit) as ei:
        jaccard.main(["jaccard", "s0.k25", "s1.k16"])
    assert ei.value.code == 1
    assert capsys.readouterr().err == "mismatched K: s1.k16\n"


# ----------------------------------------------------------------------------- SURVEY.md 8f row 2
@pytest.mark.parametrize("out,inp,args", [("r1_P03_S7.k25", "r1.k25", ["-P", "0.3", "-S", "7"]), ("g1_Pdef.k25", "g1.k25", []),
                                          ("m5_P05_D.k25", "m5.k25", ["-D", "-P", "0.5"]),
                                          ("r1_P1.k25", "r1.k25", ["-P", "1.0", "-S", "123456789"])])
def test_sample_file_bytes(zot, tmp_path, out, inp, args):
    o = tmp_path / out
    zot("sample", *(args + [o, inp]))
    assert o.read_bytes() == rd(out)


@pytest.mark.parametrize("ref,out,inp", [("r1_c2.k25", "proj_r1c2_r1.k25", "r1.k25"), ("s0.k25", "proj_s0_s1.k25", "s1.k25"),
                                         ("s1.k25", "proj_s1_m5.k25", "m5.k25")])
def test_project_file_bytes(zot, tmp_path, ref, out, inp):
    o = tmp_path / out
    zot("project", ref, o, inp)
    assert o.read_bytes() == rd(out)


def test_project_mismatched_K(zot, tmp_path, capsys):
    from zotmer_b200.commands import project
    with pytest.raises(SystemExit):
        project.main(["project", "s0.k25", str(tmp_path / "x.k25"), "s1.k16"])
    assert capsys.readouterr().err == "mismatched K (16)\n"


@pytest.mark.parametrize("k,out,ins,baits", [(25, "r1_C.k25", ["r1.fq"], "baits.fa"), (25, "g1_C.k25", ["g1.fa"], "baits.fa"),
                                             (16, "mix_C.k16", ["r1.fq", "g1.fa"], "baits.fa"),
                                             (25, "r1_Cself.k25", ["r1.fq"], "g1.fa")])
def test_kmerize_C_file_bytes(zot, tmp_path, k, out, ins, baits):
    """capture mode: records holding a bait k-mer are kept whole (device: capture_records), acgt / reads cover all"""
    o = tmp_path / out
    zot("kmerize", "-C", baits, k, o, *ins)
    assert o.read_bytes() == rd(out)


@pytest.mark.parametrize("k,out,ins,args", [(25, "r1_D03_S5.k25", ["r1.fq"], ["-D", "0.3", "-S", "5"]),
                                            (25, "g1_D05.k25", ["g1.fa"], ["-D", "0.5"]),
                                            (16, "g1_D2.k16", ["g1.fa"], ["-D", "2.0", "-S", "9"])])
def test_kmerize_D_file_bytes(zot, tmp_path, k, out, ins, args):
    o = tmp_path / out
    zot("kmerize", *(args + [k, o] + ins))
    assert o.read_bytes() == rd(out)
