"""
zot_oracle -- CPU restatement (Python 3) of the reference's k-mer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (zotmer_b200/) imports this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
there only as the checker or as the timed CPU baseline.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against the
fixtures under tests/golden/data/, which are outputs of the reference's own code run in the build
container (tests/golden/make_golden.py translates the py2 sources in memory; none are copied).

The reference is pure Python 2 (no native code), so the restatement is Python as well; byte
strings are used throughout because py2 `str` is bytes.  oracle/zot_oracle.c holds the same
algorithms in plain C for inputs too large for interpreter loops.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
import json
import math
import os
import struct
import sys

M64 = (1 << 64) - 1

# ---------------------------------------------------------------------------------------------
# zotmer/library/basics.py, bits.py
# ---------------------------------------------------------------------------------------------

# basics.py:42-46 -- ASCII -> 2-bit code, None for everything that is not AaCcGgTtUu
NUC = [None] * 256
for _ch, _v in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"TtUu", 3)):
    for _b in _ch:
        NUC[_b] = _v
NUC = tuple(NUC)

PY2_SPACE = b" \t\n\r\x0b\x0c"  # what py2 str.strip() removes (== py3 bytes.strip())


def kmer(seq):
    """basics.py:48-59: string -> integer, None if a non-nucleotide is present."""
    r = 0
    for ch in seq:
        b = NUC[ch]
        if b is None:
            return None
        r = (r << 2) | b
    return r


def render(k, x):
    """basics.py:61-67."""
    out = []
    for _ in range(k):
        out.append("ACGT"[x & 3])
        x >>= 2
    return "".join(reversed(out))


def rev(x):
    """bits.py:22-31: reverse the 32 bit-pairs of a 64-bit word."""
    x = ((x >> 2) & 0x3333333333333333) | ((x & 0x3333333333333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0F) | ((x & 0x0F0F0F0F0F0F0F0F) << 4)
    x = ((x >> 8) & 0x00FF00FF00FF00FF) | ((x & 0x00FF00FF00FF00FF) << 8)
    x = ((x >> 16) & 0x0000FFFF0000FFFF) | ((x & 0x0000FFFF0000FFFF) << 16)
    x = ((x >> 32) & 0x00000000FFFFFFFF) | ((x & 0x00000000FFFFFFFF) << 32)
    return x


def rc(k, x):
    """basics.py:115-121: reverse complement.  (`~x` on a Python int is negative; `rev` masks it
    back into 64 bits, which the restatement makes explicit.)"""
    return rev(~x & M64) >> (64 - 2 * k)


def murmer(x, s):
    """basics.py:191-229: 64-bit murmur-style mix of k-mer x with seed s."""
    k = (x * 0x87c37b91114253d5) & M64
    k = ((k << 31) | (k >> 33)) & M64
    k = (k * 0x4cf5ad432745937f) & M64
    h = s ^ k
    h = ((h << 27) | (h >> 37)) & M64
    h = (h * 5 + 0x52dce729) & M64
    h ^= h >> 33
    h = (h * 0xff51afd7ed558ccd) & M64
    h ^= h >> 33
    h = (h * 0xc4ceb9fe1a85ec53) & M64
    h ^= h >> 33
    return h


def kmers_list(k, seq, both=False):
    """basics.py:303-347 (`kmersList`; `kmers` :261-301 is its generator twin).

    Sliding window over `seq` (bytes).  A byte outside AaCcGgTtUu restarts the window after it.
    With both=True the result interleaves forward k-mer and its reverse complement."""
    z = len(seq)
    msk = (1 << (2 * k)) - 1
    s = 2 * (k - 1)
    out = []
    i = 0      # window start
    j = 0      # bases currently in the window
    x = 0
    xb = 0
    while i + k <= z:
        while i + j < z and j < k:
            b = NUC[seq[i + j]]
            if b is None:
                i += j + 1
                j = 0
                x = 0
                xb = 0
            else:
                x = (x << 2) | b
                xb = (xb >> 2) | ((3 - b) << s)
                j += 1
        if j == k:
            x &= msk
            out.append(x)
            if both:
                out.append(xb)
            j -= 1
        i += 1
    return out


# ---------------------------------------------------------------------------------------------
# zotmer/library/file.py, reads.py -- sequence input
# ---------------------------------------------------------------------------------------------

def _lines(data):
    """py2 `for l in file` on POSIX: split on b'\\n' only, last unterminated line included."""
    if not data:
        return []
    ls = data.split(b"\n")
    if ls[-1] == b"":
        ls.pop()
    return ls


def read_fasta(data):
    """file.py:19-36: [(name, seq)], lines stripped then joined; text before the first '>' dropped."""
    recs = []
    nm = None
    seq = []
    for l in _lines(data):
        l = l.strip(PY2_SPACE)
        if len(l) and l[0:1] == b">":
            if nm is not None:
                recs.append((nm, b"".join(seq)))
            nm = l[1:].strip(PY2_SPACE)
            seq = []
        else:
            seq.append(l)
    if nm is not None:
        recs.append((nm, b"".join(seq)))
    return recs


def read_fastq(data):
    """file.py:38-52: groups of 4 stripped lines; a trailing partial group is dropped
    (`if grp == 4` at :51 compares a list with an int and is never true)."""
    recs = []
    grp = []
    for l in _lines(data):
        grp.append(l.strip(PY2_SPACE))
        if len(grp) == 4:
            recs.append(tuple(grp))
            grp = []
    return recs


def is_fasta(name):
    """reads.py:11-33: by file-name suffix only, after removing one .gz/.bz2."""
    for suff in (".gz", ".bz2"):
        if name.endswith(suff):
            name = name[:-len(suff)]
            break
    return name.endswith((".fa", ".fasta", ".fas", ".fna"))


def sequences(name, data):
    """reads.py:86-125: the per-record sequence (`rd[1]`) the k-mer extractor sees."""
    if is_fasta(name):
        return [r[1] for r in read_fasta(data)]
    return [r[1] for r in read_fastq(data)]


# ---------------------------------------------------------------------------------------------
# zotmer/library/codec64.py, files.py -- stream codec
# ---------------------------------------------------------------------------------------------

# codec64.py:26-40 -- (width, count) used when n values are pending; never more than 6 per word
LOOKUP = [(0, 64)] * 61
for _i in range(1, 61):
    LOOKUP[_i] = (60 // _i, _i) if 60 % _i == 0 else LOOKUP[_i - 1]
WIDTH_OF_TAG = {}
for _i in range(1, 61):
    WIDTH_OF_TAG[60 // _i] = _i


def encode(xs):
    """codec64.py:82-120: greedy packing of values into 64-bit words (4-bit tag = count, then
    `count` fields of 60//count bits, first value lowest).  A value wider than 60 bits with nothing
    pending raises IndexError exactly as the reference does (:93-99 with _lookup[0] == (0, 64))."""
    out = []
    stk = []
    mw = 0
    for x in xs:
        wx = x.bit_length()
        mwx = max(wx, mw)
        n = len(stk)
        if n == 60 or mwx > LOOKUP[n + 1][0] or n >= LOOKUP[n + 1][1]:
            b, m0 = LOOKUP[n]
            v = 0
            for m in range(m0 - 1, -1, -1):
                v = (v << b) | stk[m]          # IndexError when n == 0 (m0 == 64)
            out.append((v << 4) | m0)
            del stk[0:m0]
            mwx = max([wx] + [y.bit_length() for y in stk])
        stk.append(x)
        mw = mwx
    if stk:
        b, m0 = LOOKUP[len(stk)]
        v = 0
        for m in range(m0 - 1, -1, -1):
            v = (v << b) | stk[m]
        out.append((v << 4) | m0)
    return out


def decode(ws):
    """codec64.py:122-150."""
    out = []
    for w in ws:
        m0 = w & 15
        w >>= 4
        b = WIDTH_OF_TAG[m0]
        msk = (1 << b) - 1
        for _ in range(m0):
            out.append(w & msk)
            w >>= b
    return out


def delta(xs):
    """files.py:85-98."""
    p = 0
    out = []
    for x in xs:
        out.append(x - p)
        p = x
    return out


def undelta(ds):
    """files.py:100-110."""
    x = 0
    out = []
    for d in ds:
        x += d
        out.append(x)
    return out


def words_to_bytes(ws):
    """files.py:65-83 (`struct.pack('Q')`, native == little-endian here); struct.error past 2^64."""
    return struct.pack("<%dQ" % len(ws), *ws)


def bytes_to_words(s):
    """files.py:54-63."""
    assert (len(s) & 7) == 0
    return list(struct.unpack("<%dQ" % (len(s) // 8), s))


# ---------------------------------------------------------------------------------------------
# zotmer/library/container/casket.py, kmers.py -- container
# ---------------------------------------------------------------------------------------------

class CasketWriter(object):
    """casket.py:113-234 (mode 'w') + kmers.py:18-21: blobs, then `__meta__` JSON, then the TOC
    JSON {name: [[offset, length], ...]} and its byte length as a little-endian u64."""

    def __init__(self, path):
        self.f = open(path, "wb")
        self.toc = {}
        self.meta = {}
        self.pos = 0

    def add(self, name, blob):
        self.toc.setdefault(name, []).append((self.pos, len(blob)))
        self.f.write(blob)
        self.pos += len(blob)

    def close(self, with_meta=True):
        if with_meta:
            self.add("__meta__", json.dumps(self.meta).encode("latin-1"))
        t = json.dumps(self.toc).encode("latin-1")
        self.f.write(t)
        self.f.write(struct.pack("<Q", len(t)))
        self.f.close()


class CasketReader(object):
    """casket.py:113-128,182-187,219-225 (mode 'r'): `open` uses the LAST toc entry of a name."""

    def __init__(self, path, with_meta=True):
        with open(path, "rb") as f:
            self.data = f.read()
        (z,) = struct.unpack("<Q", self.data[-8:])
        self.toc = json.loads(self.data[-(8 + z):-8])
        self.meta = json.loads(self.blob("__meta__")) if with_meta else {}

    def blob(self, name):
        p, l = self.toc[name][-1]          # KeyError for a missing entry, as casket.py:185
        return self.data[p:p + l]


def read_kmers(z, nm="kmers"):
    """files.py:152-153 (readKmers -> readDeltas -> undelta(decode(readWords)))."""
    return undelta(decode(bytes_to_words(z.blob(nm))))


def read_counts(z, nm="counts"):
    """files.py:158-159."""
    return decode(bytes_to_words(z.blob(nm)))


def read_kmers_and_counts(z):
    """files.py:219-227 + muxKmersAndCounts :171-193 (asserts equal length)."""
    xs = read_kmers(z)
    cs = read_counts(z)
    assert len(xs) == len(cs)
    return xs, cs


def write_kmers_and_counts(z, xs, cs):
    """files.py:195-217: 'kmers' = delta+codec64, then 'counts' = codec64."""
    z.add("kmers", words_to_bytes(encode(delta(xs))))
    z.add("counts", words_to_bytes(encode(cs)))


# ---------------------------------------------------------------------------------------------
# zotmer/commands/kmerize.py
# ---------------------------------------------------------------------------------------------

def count_sorted(ys):
    """kmerize.py:41-132 with an empty left run: run-length encode a sorted list."""
    zs = []
    ss = []
    j = 0
    n = len(ys)
    while j < n:
        y = ys[j]
        c = 0
        while j < n and ys[j] == y:
            c += 1
            j += 1
        zs.append(y)
        ss.append(c)
    return zs, ss


def kmerize_core(k, inputs):
    """kmerize.py:463-545, in-memory path (the spill path gives byte-identical output -- golden
    `spill_equals_inmemory`).  inputs = [(file name, file bytes)].
    Returns (kmers, counts, hist{count: n} in first-occurrence order, acgt[4] ints, n_records)."""
    buf = []
    acgt = [0, 0, 0, 0]
    nr = 0
    for name, data in inputs:
        for seq in sequences(name, data):
            xs = kmers_list(k, seq, True)          # reads.py:113-114: both strands
            for x in xs:                           # kmerize.py:492-493
                acgt[x & 3] += 1
            buf.extend(xs)                         # :522
            nr += 1
    buf.sort()                                     # misc.py:400-424 radix_sort == ascending sort
    xs, cs = count_sorted(buf)                     # KmerAccumulator2.flush :412-424
    h = {}
    for c in cs:                                   # :544-545
        h[c] = 1 + h.get(c, 0)
    return xs, cs, h, acgt, nr


def cmd_kmerize(k, out, input_paths):
    """kmerize.py:450-562 end to end (options -m/-C/-D/-S/-v at their defaults)."""
    inputs = []
    for p in input_paths:
        with open(p, "rb") as f:
            inputs.append((p, f.read()))
    xs, cs, h, acgt, nr = kmerize_core(k, inputs)
    z = CasketWriter(out)
    try:
        write_kmers_and_counts(z, xs, cs)
        n = float(sum(acgt))
        fr = [c / n for c in acgt]                 # ZeroDivisionError on empty input, :554-555
    except Exception:
        z.f.close()
        raise
    z.meta["K"] = k
    z.meta["kmers"] = "kmers"
    z.meta["counts"] = "counts"
    z.meta["hist"] = h
    z.meta["acgt"] = fr
    z.meta["reads"] = nr
    z.close()


# ---------------------------------------------------------------------------------------------
# zotmer/commands/merge.py
# ---------------------------------------------------------------------------------------------

def merge_core(sets):
    """merge.py:26-86 (pairwise) + :127-163 (mergeNinto): union of sorted (kmer, count) lists with
    counts summed; hist in first-occurrence order along the sorted output; acgt[x&3] += count."""
    acc = {}
    for xs, cs in sets:
        for x, c in zip(xs, cs):
            acc[x] = c + acc.get(x, 0)
    items = sorted(acc.items())
    h = {}
    acgt = [0, 0, 0, 0]
    for x, c in items:
        h[c] = 1 + h.get(c, 0)
        acgt[x & 3] += c
    return [x for x, _ in items], [c for _, c in items], h, acgt


def cmd_merge(out, input_paths):
    """merge.py:165-253.  >= 3 inputs: the working path.  2 inputs: meta holds only hist/acgt and
    acgt is NOT count-weighted (:88-92,191-199).  1 input: ZeroDivisionError (:181-196).
    K is compared as the reference does: the second file of the FIRST pair is never checked
    (:206-207 sets K from z0 only)."""
    n_in = len(input_paths)
    if n_in <= 2:
        zs = [CasketReader(p) for p in input_paths]
        K = zs[0].meta["K"]
        if n_in == 2 and zs[1].meta["K"] != K:
            sys.stderr.write("mismatched K\n")
            sys.exit(1)
        w = CasketWriter(out)
        h = {}
        acgt = [0, 0, 0, 0]
        if n_in == 1:
            xs, cs = read_kmers_and_counts(zs[0])   # hist() generator never consumed: h, acgt stay empty
        else:
            xs, cs, _, _ = merge_core([read_kmers_and_counts(z) for z in zs])
            for x, c in zip(xs, cs):
                h[c] = 1 + h.get(c, 0)
                acgt[x & 3] += 1
        write_kmers_and_counts(w, xs, cs)
        try:
            n = float(sum(acgt))
            fr = [c / n for c in acgt]
        except ZeroDivisionError:
            w.f.close()
            raise
        w.meta["hist"] = h
        w.meta["acgt"] = fr
        w.close()
        return
    K = None
    sets = []
    for i in range(0, n_in, 2):
        grp = input_paths[i:i + 2]
        zs = [CasketReader(p) for p in grp]
        if K is None:
            K = zs[0].meta["K"]
        else:
            for z in zs:
                if z.meta["K"] != K:
                    sys.stderr.write("mismatched K\n")
                    sys.exit(1)
        for z in zs:
            sets.append(read_kmers_and_counts(z))
    xs, cs, h, acgt = merge_core(sets)
    w = CasketWriter(out)
    write_kmers_and_counts(w, xs, cs)
    n = float(sum(acgt))
    fr = [c / n for c in acgt]
    w.meta["K"] = K
    w.meta["kmers"] = "kmers"
    w.meta["counts"] = "counts"
    w.meta["hist"] = h
    w.meta["acgt"] = fr
    w.close()


# ---------------------------------------------------------------------------------------------
# zotmer/commands/trim.py, hist.py, dump.py, info.py
# ---------------------------------------------------------------------------------------------

def trim_core(xs, cs, c, C=None):
    """trim.py:54-62."""
    ox = []
    oc = []
    for x, f in zip(xs, cs):
        if f >= c and (C is None or f <= C):
            ox.append(x)
            oc.append(f)
    return ox, oc


def cmd_trim(out, inp, c, C0=0):
    """trim.py:64-95 with '-C' at its documented default.  c == 0 (cut-off
    inference) is a TypeError in the reference (JSON turned hist keys into str, :26-27,82-84)."""
    z = CasketReader(inp)
    K = z.meta["K"]
    h = z.meta["hist"]
    if c == 0:
        raise TypeError("unsupported operand type(s) for -: 'str' and 'str'")
    C = C0 if C0 > 0 else None
    xs, cs = read_kmers_and_counts(z)
    w = CasketWriter(out)
    w.meta = dict(z.meta)
    del w.meta["kmers"]
    del w.meta["counts"]
    ox, oc = trim_core(xs, cs, c, C)
    write_kmers_and_counts(w, ox, oc)
    w.meta["K"] = K
    w.meta["kmers"] = "kmers"
    w.meta["counts"] = "counts"
    w.meta["hist"] = h                              # the ORIGINAL histogram, :95
    w.close()


def sub(s, p, x):
    """library/basics.py:251-259."""
    return float(murmer(x, s)) / float(0x1FFFFFFFFFFFFFFF) < p


def cmd_kmerize_D(k, out, input_paths, d, S=0):
    """kmerize.py:450-562 with -D d [-S S]: acgt over every k-mer (:492-493), only k-mers with sub(S, d, x) are
    accumulated (:494-506; the two caches only memoise sub())."""
    inputs = []
    for p in input_paths:
        with open(p, "rb") as f:
            inputs.append((p, f.read()))
    xs, cs, h, acgt, nr = kmerize_core(k, inputs)
    keep = [i for i, x in enumerate(xs) if sub(S, d, x)]
    xs = [xs[i] for i in keep]
    cs = [cs[i] for i in keep]
    h = {}
    for c in cs:
        h[c] = 1 + h.get(c, 0)
    z = CasketWriter(out)
    write_kmers_and_counts(z, xs, cs)
    n = float(sum(acgt))
    z.meta["K"] = k
    z.meta["kmers"] = "kmers"
    z.meta["counts"] = "counts"
    z.meta["hist"] = h
    z.meta["acgt"] = [c / n for c in acgt]
    z.meta["reads"] = nr
    z.close()


def cmd_kmerize_C(k, out, input_paths, bait_path):
    """kmerize.py:450-562 with -C BAITS: B = the k-mers (both strands) of every record of the bait FASTA (:478-483);
    acgt over every k-mer of every record (:492-493); a record's k-mers are accumulated, all of them, iff one of them
    is in B (:507-517)."""
    with open(bait_path, "rb") as f:
        bait_data = f.read()
    B = set()
    for seq in sequences("baits.fa", bait_data):       # readFasta whatever the suffix (:481)
        B |= set(kmers_list(k, seq, True))
    buf = []
    acgt = [0, 0, 0, 0]
    nr = 0
    for p in input_paths:
        with open(p, "rb") as f:
            data = f.read()
        for seq in sequences(p, data):
            xs = kmers_list(k, seq, True)
            for x in xs:
                acgt[x & 3] += 1
            if any(x in B for x in xs):
                buf.extend(xs)
            nr += 1
    buf.sort()
    xs, cs = count_sorted(buf)
    h = {}
    for c in cs:
        h[c] = 1 + h.get(c, 0)
    z = CasketWriter(out)
    write_kmers_and_counts(z, xs, cs)
    n = float(sum(acgt))
    z.meta["K"] = k
    z.meta["kmers"] = "kmers"
    z.meta["counts"] = "counts"
    z.meta["hist"] = h
    z.meta["acgt"] = [c / n for c in acgt]
    z.meta["reads"] = nr
    z.close()


def sample_core(xs, cs, p, S):
    """commands/sample.py:27-34 sampleD (the path docopt always selects, :53)."""
    M = 0xFFFFFFFFFF
    ox, oc = [], []
    for x, c in zip(xs, cs):
        if float(murmer(x, S) & M) / float(M) < p:
            ox.append(x)
            oc.append(c)
    return ox, oc


def cmd_sample(out, inp, p=0.01, S=0):
    """commands/sample.py:36-67."""
    z0 = CasketReader(inp)
    K = z0.meta["K"]
    xs, cs = read_kmers_and_counts(z0)
    z = CasketWriter(out)
    z.meta = dict(z0.meta)
    del z.meta["kmers"]
    del z.meta["counts"]
    ox, oc = sample_core(xs, cs, p, S)
    h = {}
    for c in oc:
        h[c] = 1 + h.get(c, 0)
    write_kmers_and_counts(z, ox, oc)
    z.meta["K"] = K
    z.meta["kmers"] = "kmers"
    z.meta["counts"] = "counts"
    z.meta["hist"] = h
    z.close()


def cmd_project(ref, out, inp):
    """commands/project.py:42-70 (inputs with counts: project2 :30-40)."""
    zr = CasketReader(ref)
    K = zr.meta["K"]
    rs = set(read_kmers(zr))
    z0 = CasketReader(inp)
    if z0.meta["K"] != K:
        raise SystemExit(1)
    xs, cs = read_kmers_and_counts(z0)
    z = CasketWriter(out)
    z.meta["K"] = K
    ox = [x for x in xs if x in rs]
    oc = [c for x, c in zip(xs, cs) if x in rs]
    write_kmers_and_counts(z, ox, oc)
    z.meta["kmers"] = "kmers"
    z.meta["counts"] = "counts"
    z.meta["hist"] = z0.meta["hist"]
    z.close()


def cmd_hist(input_paths):
    """hist.py:14-24."""
    out = []
    for inp in input_paths:
        z = CasketReader(inp)
        if "hist" in z.meta:
            for f, c in sorted((int(f), c) for f, c in z.meta["hist"].items()):
                out.append("%s\t%d\t%d\n" % (inp, f, c))
    return "".join(out)


def cmd_dump(inp):
    """dump.py:13-29."""
    z = CasketReader(inp)
    K = z.meta["K"]
    xs, cs = read_kmers_and_counts(z)
    return "".join("%s\t%d\n" % (render(K, x), c) for x, c in zip(xs, cs))


# ---------------------------------------------------------------------------------------------
# zotmer/library/dist.py, commands/dist.py, commands/jaccard.py
# ---------------------------------------------------------------------------------------------

def split(xs, ys):
    """library/dist.py:241-265: (|X n Y|, |X \\ Y|, |Y \\ X|) of two sorted duplicate-free lists."""
    i = j = b = dx = dy = 0
    xz, yz = len(xs), len(ys)
    while i < xz and j < yz:
        x, y = xs[i], ys[j]
        if x < y:
            dx += 1
            i += 1
        elif x > y:
            dy += 1
            j += 1
        else:
            b += 1
            i += 1
            j += 1
    return b, dx + xz - i, dy + yz - j


def _bray(a, b, c):
    return float(b + c) / float(2 * a + b + c)           # library/dist.py:40-41, 209-210


def _chord(a, b, c):
    return math.sqrt(2 * (1 - a / math.sqrt((a + b) * (a + c))))   # :67-68, 93-94


def _jacc(a, b, c):
    return float(b + c) / float(a + b + c)               # :112-113


def _kulc(a, b, c):
    a, b, c = float(a), float(b), float(c)               # :168-172
    return 1 - 0.5 * (a / (a + b) + a / (a + c))


def _ochi(a, b, c):
    return 1 - a / math.sqrt((a + b) * (a + c))          # :190-191


def _whit(a, b, c):
    a, b, c = float(a), float(b), float(c)               # :235-239
    return 0.5 * (b / (a + b) + c / (a + c) + abs(a / (a + b) - a / (a + c)))


# commands/dist.py:59-92 -- the working (`vec=False`) measures
QUAL = {
    "bray.curtis.qual": _bray, "chord.qual": _chord, "hellinger.qual": _chord, "jaccard.qual": _jacc,
    "kulczynski.qual": _kulc, "ochiai.qual": _ochi, "sorensen.qual": _bray, "whittaker.qual": _whit,
}


def prep(K, path):
    """commands/dist.py:29-49 (`vec=False` branch): project to K-mers and drop adjacent duplicates."""
    z = CasketReader(path)
    fK = z.meta["K"]
    if fK < K:
        raise ValueError("incompatible values of K: %d & %d" % (K, fK))
    S = 2 * (fK - K)
    v = []
    for x in read_kmers(z):
        y = x >> S
        if not v or v[-1] != y:
            v.append(y)
    return v


def cmd_dist(measures, K, paths):
    """commands/dist.py:94-168 for a list of exact `.qual` measure names."""
    ms = sorted(set(measures))
    out = ["\t".join(["lhs.name", "rhs.name"] + ms) + "\n"]
    fmt = "\t".join(["%s", "%s"] + ["%g"] * len(ms)) + "\n"
    for i in range(len(paths)):
        lhs = prep(K, paths[i])
        for j in range(i + 1, len(paths)):
            rhs = prep(K, paths[j])
            vs = [paths[i], paths[j]]
            for m in ms:
                vs.append(QUAL[m](*split(lhs, rhs)))     # one split() per measure, as :161-167
            out.append(fmt % tuple(vs))
    return "".join(out)


def jaccard(xs, ys):
    """commands/jaccard.py:31-54: (|X n Y|, |X u Y|, index)."""
    b, dx, dy = split(xs, ys)
    return b, b + dx + dy, float(b) / float(b + dx + dy)


def cmd_jaccard(paths, all_pairs=False):
    """commands/jaccard.py:144-166 (k-mer-set inputs, no -p)."""
    out = []
    Z = len(paths) if all_pairs else 1
    for i in range(Z):
        z0 = CasketReader(paths[i])
        xs = read_kmers(z0)
        for j in range(i + 1, len(paths)):
            z1 = CasketReader(paths[j])
            ys = read_kmers(z1)
            if z0.meta["K"] != z1.meta["K"]:
                sys.stderr.write("mismatched K: %s\n" % paths[j])
                sys.exit(1)
            isec, union, d = jaccard(xs, ys)
            out.append("%s\t%s\t%d\t%d\t%d\t%d\t%f\n" % (paths[i], paths[j], len(xs), len(ys), isec, union, d))
    return "".join(out)
