"""
Deterministic synthetic inputs of the shapes BASELINE.json names, generated with numpy
so that 150 Mbases take a couple of seconds.  Used by bench.py, tools/ and the full-size GPU tests.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    COMP[_a] = _b


def genome(n, seed=17, repeats=True):
    """i.i.d. ACGT with a few inserted repeat families so that the count histogram has mass above 2."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = ACGT[rng.integers(0, 4, n)]
    if repeats and n >= 200000:
        for (length, copies) in ((5000, 7), (1200, 20)):
            unit = ACGT[rng.integers(0, 4, length)]
            for c in range(copies):
                p = int(rng.integers(0, n - length))
                g[p:p + length] = unit if c % 2 == 0 else COMP[unit[::-1]]
    return g


def fasta_bytes(g, name=b"chr1", width=80):
    n = len(g)
    rows = (n + width - 1) // width
    pad = rows * width - n
    body = np.concatenate([g, np.full(pad, ord("\n"), np.uint8)]).reshape(rows, width)
    out = np.concatenate([body, np.full((rows, 1), ord("\n"), np.uint8)], axis=1).reshape(-1)
    if pad:
        out = out[:len(out) - pad - 1]
        out = np.concatenate([out, np.frombuffer(b"\n", np.uint8)])
    return b">" + name + b"\n" + out.tobytes()


def fastq_array(g, nreads, L=150, seed=18, err=0.005, pn=0.0002):
    """-> uint8 array [nreads, reclen] of 4-line FASTQ records with fixed-width names."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = rng.integers(0, len(g) - L, nreads)
    idx = pos[:, None] + np.arange(L)[None, :]
    s = g[idx]
    rev = rng.random(nreads) < 0.5
    s[rev] = COMP[s[rev][:, ::-1]]
    e = rng.random(s.shape) < err
    s[e] = ACGT[(np.searchsorted(ACGT, s[e]) + rng.integers(1, 4, int(e.sum()))) & 3]
    nn = rng.random(s.shape) < pn
    s[nn] = ord("N")
    ids = np.arange(nreads)
    digits = np.stack([(ids // 10 ** d) % 10 for d in range(8, -1, -1)], axis=1).astype(np.uint8) + ord("0")
    reclen = 1 + 9 + 1 + L + 1 + 2 + L + 1
    rec = np.empty((nreads, reclen), np.uint8)
    o = 0
    rec[:, o] = ord("@"); o += 1
    rec[:, o:o + 9] = digits; o += 9
    rec[:, o] = ord("\n"); o += 1
    rec[:, o:o + L] = s; o += L
    rec[:, o] = ord("\n"); o += 1
    rec[:, o] = ord("+"); rec[:, o + 1] = ord("\n"); o += 2
    rec[:, o:o + L] = ord("I"); o += L
    rec[:, o] = ord("\n")
    return rec


def mutate(g, rate, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    h = g.copy()
    e = rng.random(len(h)) < rate
    h[e] = ACGT[(np.searchsorted(ACGT, h[e]) + rng.integers(1, 4, int(e.sum()))) & 3]
    return h


def bgzf_bytes(data, level=6, block=65280, eof=True, strategy=0):
    """`data` as bgzip writes it (BGZF, SAM spec 4.1): gzip members of <= 65280 bytes of input, each with a 'BC' extra
    field holding its total size - 1, closed by the 28-byte empty member."""
    import struct
    import zlib
    out = []
    pieces = [data[i:i + block] for i in range(0, len(data), block)]
    if eof:
        pieces.append(b"")
    for piece in pieces:
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        body = c.compress(bytes(piece)) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        assert bsize <= 65536
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + body +
                   struct.pack("<II", zlib.crc32(bytes(piece)) & 0xffffffff, len(piece)))
    return b"".join(out)
