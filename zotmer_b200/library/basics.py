"""Host-side k-mer helpers (mirrors zotmer/library/basics.py:42-121,191-229; bits.py:22-31)."""
import numpy as np

_nuc = {'A': 0, 'a': 0, 'C': 1, 'c': 1, 'G': 2, 'g': 2, 'T': 3, 't': 3, 'U': 3, 'u': 3}
M64 = (1 << 64) - 1


def kmer(seq):
    "basics.py:48-59"
    r = 0
    for ch in seq:
        b = _nuc.get(ch if isinstance(ch, str) else chr(ch))
        if b is None:
            return None
        r = (r << 2) | b
    return r


def render(k, x):
    "basics.py:61-67"
    x = int(x)
    r = []
    for i in range(k):
        r.append("ACGT"[x & 3])
        x >>= 2
    return ''.join(r[::-1])


def renderMany(k, xs):
    """render() for a whole uint64 array -> list of str (vectorised; used by `zot dump`)."""
    xs = np.ascontiguousarray(xs, dtype=np.uint64)
    shifts = (np.arange(k - 1, -1, -1, dtype=np.uint64) * np.uint64(2))[None, :]
    codes = ((xs[:, None] >> shifts) & np.uint64(3)).astype(np.uint8)
    letters = np.frombuffer(b"ACGT", np.uint8)[codes]
    return [row.tobytes().decode('ascii') for row in letters]


def rev(x):
    "bits.py:22-31"
    x = ((x >> 2) & 0x3333333333333333) | ((x & 0x3333333333333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0F) | ((x & 0x0F0F0F0F0F0F0F0F) << 4)
    x = ((x >> 8) & 0x00FF00FF00FF00FF) | ((x & 0x00FF00FF00FF00FF) << 8)
    x = ((x >> 16) & 0x0000FFFF0000FFFF) | ((x & 0x0000FFFF0000FFFF) << 16)
    x = ((x >> 32) & 0x00000000FFFFFFFF) | ((x & 0x00000000FFFFFFFF) << 32)
    return x


def rc(k, x):
    "basics.py:115-121"
    return rev(~x & M64) >> (64 - 2 * k)


def murmer(x, s):
    "basics.py:191-229"
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
