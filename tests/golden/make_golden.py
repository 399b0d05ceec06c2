#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/data/ by RUNNING THE REFERENCE ITSELF
(/root/reference, translated py2->py3 in memory by py2shim.py; no reference source is copied).

Run inside the build container only:   python tests/golden/make_golden.py
The outputs (small FASTA/FASTQ inputs, the k-mer set files the reference wrote for them, captured
stdout of dist/jaccard/hist/dump/info, and function-level known-answer vectors in kat.json) are
committed; the GPU box and the test-suite only ever read the committed files.
"""
import contextlib
import io
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import py2shim  # noqa: E402

DATA = os.path.join(HERE, "data")


def rnd_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def mutate(rng, s, rate):
    b = list(s)
    for i in np.nonzero(rng.random(len(b)) < rate)[0]:
        b[i] = "ACGT"[("ACGT".index(b[i]) + int(rng.integers(1, 4))) % 4]
    return "".join(b)


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def wrap(s, w):
    return "\n".join(s[i:i + w] for i in range(0, len(s), w))


def write(path, text):
    with open(path, "wb") as f:
        f.write(text.encode("latin-1"))


def make_inputs():
    rng = np.random.default_rng(17)
    # ---- g1.fa : multi-record FASTA with every lexical wrinkle of readFasta (file.py:19-36)
    chrom = rnd_seq(rng, 3000)
    rep = rnd_seq(rng, 120)
    chrom = chrom[:500] + rep + chrom[500:1500] + revcomp(rep) + chrom[1500:2200] + rep + chrom[2200:]
    pal = "ACGTTGCAAGCTTGCAACGT"  # even-length reverse-palindrome (matters for even k)
    rec2 = rnd_seq(rng, 400)
    rec2 = rec2[:100] + "NNNNNNNNNN" + rec2[100:200].lower() + "R" + rec2[200:300].replace("T", "U") + pal + rec2[300:]
    rec3 = rnd_seq(rng, 30)  # shorter than k=31/32 but longer than 25
    rec4 = rnd_seq(rng, 10)  # shorter than k
    rec5 = rnd_seq(rng, 200)
    fa = []
    fa.append("this line precedes any header and is ignored ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
    fa.append(">chr1 synthetic chromosome\n" + wrap(chrom, 60) + "\n")
    fa.append(">rec2 wrinkles\n" + wrap(rec2, 70) + "\n\n")
    fa.append("  >rec3 header with leading blanks\n" + rec3 + "\n")
    fa.append(">rec4\n" + rec4 + "\n")
    fa.append(">empty\n")
    fa.append(">rec5 crlf and inner blanks\r\n" + rec5[:80] + "\r\n  " + rec5[80:120] + " " + rec5[120:160] + "\t\r\n" + rec5[160:])
    write(os.path.join(DATA, "g1.fa"), "".join(fa))  # no trailing newline on purpose

    # ---- kat6.fa : SURVEY.md 8c KAT6
    write(os.path.join(DATA, "kat6.fa"), ">s1 desc\nACGTACGTNACG\nTACgu\n>s2\nAAAAAAA\n")

    # ---- r1.fq : FASTQ, 4-line records, errors, N, short reads, CRLF, trailing partial record
    genome = rnd_seq(rng, 2000)
    fq = []
    for i in range(300):
        L = int(rng.choice([100, 100, 100, 75, 30, 24, 0]))
        p = int(rng.integers(0, len(genome) - 100))
        s = genome[p:p + L]
        if rng.random() < 0.5:
            s = revcomp(s)
        s = mutate(rng, s, 0.01)
        if L and rng.random() < 0.1:
            q = int(rng.integers(0, L))
            s = s[:q] + "N" + s[q + 1:]
        if L and rng.random() < 0.05:
            s = s.lower()
        eol = "\r\n" if i % 50 == 7 else "\n"
        fq.append("@read%d/1%s%s%s+%s%s%s" % (i, eol, s, eol, eol, "I" * L, eol))
    fq.append("@partial\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\n")  # dropped: only 3 lines (file.py:45-52)
    write(os.path.join(DATA, "r1.fq"), "".join(fq))

    # ---- baits.fa : bait sequences for `zot kmerize -C` (pieces of the genome r1.fq was read from, one of them
    # reverse-complemented, and a piece of g1.fa's chr1); no random numbers drawn here
    write(os.path.join(DATA, "baits.fa"), ">b1\n%s\n>b2 rc\n%s\n>b3 chr1\n%s\n" % (
        genome[300:380], revcomp(genome[1200:1260]), wrap(chrom[700:790], 40)))

    # ---- r2.fq : enough records (2500 x 60 bp) to force >= 2 spills with -m 1 (kmerize.py:527-539)
    fq = []
    for i in range(2500):
        p = int(rng.integers(0, len(genome) - 60))
        s = mutate(rng, genome[p:p + 60], 0.005)
        if rng.random() < 0.5:
            s = revcomp(s)
        fq.append("@r%d\n%s\n+\n%s\n" % (i, s, "I" * 60))
    write(os.path.join(DATA, "r2.fq"), "".join(fq))

    # ---- s0..s4.fa : related genomes for merge / dist / jaccard
    base = rnd_seq(rng, 1500)
    for i, rate in enumerate([0.0, 0.01, 0.03, 0.1, 0.75]):
        g = mutate(rng, base, rate)
        write(os.path.join(DATA, "s%d.fa" % i), ">s%d\n%s\n" % (i, wrap(g, 80)))
    # ---- j.fa : multi-record FASTA for `zot jaccard`'s single-FASTA mode (jaccard.py:110-142)
    write(os.path.join(DATA, "j.fa"), "".join(">rec%d some description\n%s\n" % (i, wrap(mutate(rng, base[:400], r), 50))
                                               for i, r in enumerate([0.0, 0.01, 0.05])))


def run_cmd(modname, opts, out_txt=None, expect_exc=None):
    """Call zotmer.commands.<modname>.main with docopt stubbed to return `opts`."""
    mod = sys.modules["zotmer.commands." + modname]
    py2shim.set_docopt(lambda doc, argv=None, **kw: dict(opts))
    so, se = io.StringIO(), io.StringIO()
    exc = None
    with contextlib.redirect_stdout(so), contextlib.redirect_stderr(se):
        try:
            with sys.modules["zotmer.library.file"].autoremove():
                mod.main([modname])
        except SystemExit as e:
            exc = "SystemExit(%r)" % (e.code,)
        except Exception as e:  # behavioural KATs: the reference's crash class is the golden
            exc = type(e).__name__
    if expect_exc == "*":
        pass
    elif expect_exc is not None:
        assert exc == expect_exc, (modname, opts, exc, se.getvalue())
    else:
        assert exc is None, (modname, opts, exc, se.getvalue())
    if out_txt is not None:
        with open(os.path.join(DATA, out_txt), "w") as f:
            f.write(so.getvalue())
        if se.getvalue():
            with open(os.path.join(DATA, out_txt + ".stderr"), "w") as f:
                f.write(se.getvalue())
    return so.getvalue(), se.getvalue(), exc


def kmerize(k, out, inputs, mem=None):
    return run_cmd("kmerize", {"<k>": str(k), "<output>": os.path.join(DATA, out),
                               "<input>": [os.path.join(DATA, i) for i in inputs],
                               "-m": mem, "-C": None, "-D": None, "-S": None, "-v": False})


def main():
    if os.path.isdir(DATA):
        shutil.rmtree(DATA)
    os.makedirs(DATA)
    os.chdir(DATA)  # so that file names printed by dist/jaccard/hist are relative
    py2shim.load()
    from zotmer.library import basics, codec64, bits
    import zotmer.library.dist as ldist
    import zotmer.commands.jaccard as cj

    make_inputs()
    kat = {}

    # ------------------------------------------------------------------ function-level KATs
    rng = np.random.default_rng(99)
    seqs = ["", "A", "ACGT", "ACGTACGTNACGTACgu", "A" * 31 + "CG", "N" * 40, "acgtnACGTUuRYacgtacgtacgt",
            rnd_seq(rng, 200), rnd_seq(rng, 64).lower(), "ACGTTGCAAGCTTGCAACGT" * 3,
            rnd_seq(rng, 40) + "N" + rnd_seq(rng, 31) + "-" + rnd_seq(rng, 32) + " " + rnd_seq(rng, 33) + "\xff"]
    kl = []
    for s in seqs:
        for k in (1, 4, 5, 16, 25, 30, 31, 32):
            for both in (True, False):
                kl.append({"k": k, "seq": s, "both": both,
                           "out": [str(x) for x in basics.kmersList(k, s, both)]})
                assert list(basics.kmers(k, s, both)) == basics.kmersList(k, s, both)
    kat["kmersList"] = kl
    kat["rc"] = [{"k": k, "x": str(x), "out": str(basics.rc(k, x))}
                 for k in (1, 5, 16, 25, 30, 31, 32)
                 for x in [int(v) & ((1 << (2 * k)) - 1) for v in rng.integers(0, 2 ** 63, 6)]]
    kat["rev"] = [{"x": str(int(x)), "out": str(bits.rev(int(x)))} for x in rng.integers(0, 2 ** 63, 8)]
    kat["murmer"] = [{"x": str(x), "s": s, "out": str(basics.murmer(x, s))}
                     for (x, s) in [(0, 0), (0x1234567, 17), (2 ** 50 - 1, 0), (2 ** 62 + 12345, 3), (1, 1)]]
    kat["render"] = [{"k": k, "x": str(x), "out": basics.render(k, x)} for (k, x) in [(5, 108), (25, 2 ** 50 - 7), (31, 12345678901234567)]]

    enc = []
    lists = [[0, 1, 1023, 1024, 2 ** 30 - 1, 2 ** 30, 2 ** 59, 5, 5, 5, 5, 5, 5, 5], [1] * 20, [], [0], [2 ** 60 - 1],
             [2 ** 60 - 1, 0, 0, 0, 0, 0, 0, 2 ** 60 - 1], [1023] * 7 + [1024] * 7 + [4095] * 7 + [4096] * 7 + [32767] * 7 + [32768] * 3]
    for w in (3, 10, 12, 15, 20, 30, 45, 60):
        lists.append([int(v) >> (63 - w) for v in rng.integers(0, 2 ** 63, 50)])
    lists.append([int(v) >> int(s) for v, s in zip(rng.integers(0, 2 ** 63, 400), rng.integers(3, 63, 400))])
    for xs in lists:
        ws = list(codec64.encode(xs))
        assert codec64.decodeList(ws) == xs
        enc.append({"xs": [str(x) for x in xs], "ws": [str(w) for w in ws]})
    kat["codec64"] = enc
    kat["codec64_lookup"] = [list(t) for t in codec64._lookup[:12]]
    for bad, name in (([2 ** 61], "first"), ([5, 2 ** 61, 7], "later")):
        try:
            list(codec64.encode(bad))
            r = "ok"
        except Exception as e:
            r = type(e).__name__
        kat["codec64_overflow_" + name] = r

    sp = []
    for _ in range(12):
        nx, ny = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        xs = sorted(set(int(v) for v in rng.integers(0, 100, nx)))
        ys = sorted(set(int(v) for v in rng.integers(0, 100, ny)))
        ent = {"xs": xs, "ys": ys, "split": list(ldist.split(xs, ys))}
        if len(xs) and len(ys) and ent["split"][0] > 0:
            j = cj.jaccard(xs, ys)
            ent["jaccard"] = [j[0], j[1], float(j[2]).hex()]
            for nm in ("brayCurtis", "chord", "hellinger", "jaccard", "kulczynski", "ochiai", "sorensen", "whittaker"):
                ent["m_" + nm] = float(getattr(ldist, nm)(xs, ys, False)).hex()
        sp.append(ent)
    kat["split"] = sp
    kat["beta"] = [{"p": p, "m": m, "n": n, "logIx": float(cj.logIx(p, m, n)).hex(),
                    "q05": float(cj.quantBeta(0.05, m, n)).hex(), "q95": float(cj.quantBeta(0.95, m, n)).hex()}
                   for (p, m, n) in [(0.9, 11, 4), (0.5, 101, 101), (0.95, 1501, 20), (0.2, 2, 900)]]

    # ------------------------------------------------------------------ command-level goldens
    kmerize(5, "kat6.k5", ["kat6.fa"])
    for k in (5, 16, 25, 30, 31):
        kmerize(k, "g1.k%d" % k, ["g1.fa"])
    _, _, exc = run_cmd("kmerize", {"<k>": "32", "<output>": os.path.join(DATA, "g1.k32"),
                                    "<input>": [os.path.join(DATA, "g1.fa")],
                                    "-m": None, "-C": None, "-D": None, "-S": None, "-v": False}, expect_exc="*")
    kat["kmerize_k32_g1"] = exc or "ok"   # k=32: codec64 payload is 60 bits
    if exc is not None and os.path.exists(os.path.join(DATA, "g1.k32")):
        os.remove(os.path.join(DATA, "g1.k32"))
    for k in (8, 25, 31):
        kmerize(k, "r1.k%d" % k, ["r1.fq"])
    kmerize(21, "r2.k21", ["r2.fq"])
    kmerize(21, "r2_spill.k21", ["r2.fq"], mem="1")
    kat["spill_equals_inmemory"] = open(os.path.join(DATA, "r2.k21"), "rb").read() == open(os.path.join(DATA, "r2_spill.k21"), "rb").read()
    os.remove(os.path.join(DATA, "r2_spill.k21"))
    kmerize(25, "mix.k25", ["s0.fa", "r1.fq", "s1.fa"])
    for i in range(5):
        kmerize(25, "s%d.k25" % i, ["s%d.fa" % i])
        kmerize(16, "s%d.k16" % i, ["s%d.fa" % i])
    # empty input: the reference divides by zero (kmerize.py:554-555)
    write(os.path.join(DATA, "empty.fa"), ">nothing\nACGT\n")
    _, _, exc = run_cmd("kmerize", {"<k>": "25", "<output>": os.path.join(DATA, "empty.k25"),
                                    "<input>": [os.path.join(DATA, "empty.fa")],
                                    "-m": None, "-C": None, "-D": None, "-S": None, "-v": False},
                        expect_exc="ZeroDivisionError")
    os.remove(os.path.join(DATA, "empty.k25"))
    kat["kmerize_empty"] = exc

    def merge(out, ins, **kw):
        return run_cmd("merge", {"<output>": os.path.join(DATA, out), "<input>": ins}, **kw)

    merge("m3.k25", ["s0.k25", "s1.k25", "s2.k25"])
    merge("m4.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25"])
    merge("m5.k25", ["s0.k25", "s1.k25", "s2.k25", "s3.k25", "s4.k25"])
    merge("m5dup.k25", ["s0.k25", "s0.k25", "g1.k25", "r1.k25", "s0.k25"])
    merge("m2.k25", ["s0.k25", "s1.k25"])  # quirk: no K/kmers/counts meta (merge.py:173-199)
    _, _, exc = merge("m1.k25", ["s0.k25"], expect_exc="ZeroDivisionError")
    os.remove(os.path.join(DATA, "m1.k25"))
    kat["merge_one_input"] = exc
    _, se, exc = merge("mbad.k25", ["s0.k25", "s1.k25", "s2.k16"], expect_exc="SystemExit(1)")
    kat["merge_mismatched_K"] = {"exc": exc, "stderr": se}
    if os.path.exists(os.path.join(DATA, "mbad.k25")):
        os.remove(os.path.join(DATA, "mbad.k25"))

    def trim(out, inp, c, C="0", **kw):
        return run_cmd("trim", {"<output>": os.path.join(DATA, out), "<input>": inp, "-c": c, "-C": C}, **kw)

    trim("r1_c2.k25", "r1.k25", "2")
    trim("r1_c2_C3.k25", "r1.k25", "2", "3")
    trim("r2_c3.k21", "r2.k21", "3")
    trim("m5_c2.k25", "m5.k25", "2")
    trim("r1_c1000.k25", "r1.k25", "1000")  # everything removed
    _, _, exc = trim("r1_c0.k25", "r1.k25", "0", expect_exc="TypeError")
    kat["trim_c0"] = exc
    if os.path.exists(os.path.join(DATA, "r1_c0.k25")):
        os.remove(os.path.join(DATA, "r1_c0.k25"))

    sets25 = ["s%d.k25" % i for i in range(5)]
    run_cmd("hist", {"<input>": ["g1.k25", "r1.k25", "r1_c2.k25", "m5.k25", "m2.k25"]}, out_txt="hist.txt")
    run_cmd("dist", {"-M": ["*.qual"], "<k>": "25", "<input>": sets25}, out_txt="dist_qual_25.txt")
    run_cmd("dist", {"-M": ["jaccard.qual", "kulczynski.qual"], "<k>": "25", "<input>": sets25[:3] + ["g1.k25"]},
            out_txt="dist_two_25.txt")
    run_cmd("dist", {"-M": ["*.qual"], "<k>": "12", "<input>": sets25}, out_txt="dist_qual_12_of_25.txt")
    run_cmd("dist", {"-M": ["list"], "<k>": "25", "<input>": sets25}, out_txt="dist_list.txt")
    run_cmd("dist", {"-M": ["nosuch"], "<k>": "25", "<input>": sets25}, out_txt="dist_bad.txt")
    _, _, exc = run_cmd("dist", {"-M": ["jaccard.ab"], "<k>": "5", "<input>": ["kat6.k5", "g1.k5"]}, expect_exc="TypeError")
    kat["dist_vec_measure"] = exc
    _, _, exc = run_cmd("dist", {"-M": ["jaccard.qual"], "<k>": "25", "<input>": ["s0.k16", "s1.k16"]}, expect_exc="MismatchedK")
    kat["dist_K_too_big"] = exc
    run_cmd("jaccard", {"-a": False, "-b": False, "-p": None, "<input>": sets25}, out_txt="jaccard_first.txt")
    run_cmd("jaccard", {"-a": True, "-b": False, "-p": None, "<input>": sets25}, out_txt="jaccard_all.txt")
    run_cmd("jaccard", {"-a": True, "-b": False, "-p": "0.9", "<input>": sets25[:3]}, out_txt="jaccard_p.txt")
    run_cmd("jaccard", {"-a": True, "-b": False, "-p": None, "<input>": ["j.fa"]}, out_txt="jaccard_fasta.txt")
    _, se, exc = run_cmd("jaccard", {"-a": False, "-b": False, "-p": None, "<input>": ["s0.k25", "s1.k16"]}, expect_exc="SystemExit(1)")
    kat["jaccard_mismatched_K"] = {"exc": exc, "stderr": se}
    run_cmd("dump", {"<input>": "kat6.k5"}, out_txt="dump_kat6.txt")
    run_cmd("dump", {"<input>": "r1_c2.k25"}, out_txt="dump_r1_c2.txt")
    run_cmd("info", {"<input>": ["kat6.k5", "m3.k25"]}, out_txt="info.txt")

    # ---- zot sample, zot project, zot kmerize -D (docopt gives flags as False/True)

    def sample(out, inp, P=None, S=None, D=False, **kw):
        return run_cmd("sample", {"<output>": os.path.join(DATA, out), "<input>": inp, "-P": P, "-S": S, "-D": D}, **kw)

    sample("r1_P03_S7.k25", "r1.k25", "0.3", "7")
    sample("g1_Pdef.k25", "g1.k25")                       # default p = 0.01, seed 0
    sample("m5_P05_D.k25", "m5.k25", "0.5", None, True)    # -D given: same path
    sample("r1_P1.k25", "r1.k25", "1.0", "123456789")      # p = 1: u < 1 fails only for h & M == M

    def project(ref, out, inp, **kw):
        return run_cmd("project", {"<ref>": ref, "<output>": os.path.join(DATA, out), "<input>": inp}, **kw)

    project("r1_c2.k25", "proj_r1c2_r1.k25", "r1.k25")
    project("s0.k25", "proj_s0_s1.k25", "s1.k25")
    project("s1.k25", "proj_s1_m5.k25", "m5.k25")
    _, se, exc = project("s0.k25", "proj_bad.k25", "s1.k16", expect_exc="SystemExit(1)")
    kat["project_mismatched_K"] = {"exc": exc, "stderr": se}
    if os.path.exists(os.path.join(DATA, "proj_bad.k25")):
        os.remove(os.path.join(DATA, "proj_bad.k25"))

    def kmerizeD(k, out, inputs, D, S=None):
        return run_cmd("kmerize", {"<k>": str(k), "<output>": os.path.join(DATA, out),
                                   "<input>": [os.path.join(DATA, i) for i in inputs],
                                   "-m": None, "-C": None, "-D": D, "-S": S, "-v": False})

    kmerizeD(25, "r1_D03_S5.k25", ["r1.fq"], "0.3", "5")
    kmerizeD(25, "g1_D05.k25", ["g1.fa"], "0.5")
    kmerizeD(16, "g1_D2.k16", ["g1.fa"], "2.0", "9")       # u can exceed 1 (murmer is not masked to 61 bits): d = 2 keeps fewer than all
    def kmerizeC(k, out, inputs, baits):
        return run_cmd("kmerize", {"<k>": str(k), "<output>": os.path.join(DATA, out),
                                   "<input>": [os.path.join(DATA, i) for i in inputs],
                                   "-m": None, "-C": os.path.join(DATA, baits), "-D": None, "-S": None, "-v": False})

    kmerizeC(25, "r1_C.k25", ["r1.fq"], "baits.fa")               # reads that hold a bait 25-mer, whole
    kmerizeC(25, "g1_C.k25", ["g1.fa"], "baits.fa")               # FASTA records: chr1 is captured, the others are not
    kmerizeC(16, "mix_C.k16", ["r1.fq", "g1.fa"], "baits.fa")
    kmerizeC(25, "r1_Cself.k25", ["r1.fq"], "g1.fa")              # baits that share nothing with the reads: empty set

    kat["sub"] = [[s_, p_, x_, bool(basics.sub(s_, p_, x_))] for (s_, p_, x_) in
                  [(0, 0.5, 0), (0, 0.5, 1), (5, 0.3, 0x1234567), (17, 0.01, 2 ** 50 - 1), (9, 2.0, 0xFFFFFFFF), (9, 2.0, 12345)]]

    with open(os.path.join(DATA, "kat.json"), "w") as f:
        json.dump(kat, f, indent=0, sort_keys=True)
    tot = sum(os.path.getsize(os.path.join(DATA, x)) for x in os.listdir(DATA))
    print("golden fixtures written: %d files, %d bytes" % (len(os.listdir(DATA)), tot))


if __name__ == "__main__":
    main()
