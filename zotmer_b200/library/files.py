"""
Stream files inside a casket: 'kmers' = delta + codec64, 'counts' = codec64 (mirrors
zotmer/library/files.py:54-227).  The reference codes value by value in Python (54 % of its
kmerize time); here whole streams go through libzot_b200 (zb_encode/zb_decode, zb_set_encode,
zb_set_from_streams) and k-mer sets stay on the device between decode and use.
"""
import numpy as np

from zotmer_b200 import _native


def readWords(f):
    """files.py:54-63: the blob as little-endian u64 words (AssertionError on a ragged blob)."""
    s = f.view() if hasattr(f, 'view') else None     # casket entries: a view of the mapped file instead of a copy
    if s is None:
        s = f.read()
    assert (len(s) & 7) == 0
    return np.frombuffer(s, dtype='<u8')


def writeWords(f, ws):
    """files.py:65-83"""
    f.write(memoryview(np.ascontiguousarray(ws, dtype='<u8')).cast('B'))   # no intermediate bytes copy
    return len(ws)


def _names(nm):
    return ('kmers', 'counts') if nm is None else (nm + '-kmers', nm + '-counts')


def readKmers(z, nm='kmers', device=0):
    """files.py:152-153 -> uint64 array"""
    return _native.decode_stream(readWords(z.open(nm)), delta=True, device=device)


def readCounts(z, nm='counts', device=0):
    """files.py:158-159 -> uint64 array"""
    return _native.decode_stream(readWords(z.open(nm)), delta=False, device=device)


def readKmersAndCounts(z, nm=None, device=0):
    """files.py:219-227 -> (kmers, counts) arrays; AssertionError when the lengths differ (files.py:182)"""
    xNm, cNm = _names(nm)
    xs = readKmers(z, xNm, device)
    cs = readCounts(z, cNm, device)
    assert len(xs) == len(cs)
    return xs, cs


def stageKmerSet(z, nm=None, counts=True, device=0):
    """start moving the word streams of container `z` to the device (library I/O threads: pread -> pinned ring -> H2D)
    -> token for finishKmerSet.  `z` must stay open until then."""
    xNm, cNm = _names(nm)
    fd = z.fo.fileno()
    toks = []
    for name in ((xNm, cNm) if counts else (xNm,)):
        offset, length = z.toc[name][-1]          # KeyError for a name that is not there (casket.open)
        assert (length & 7) == 0                  # files.py:58
        toks.append(_native.stage_fd(fd, offset, length, device))
    return toks


def finishKmerSet(toks, device=0):
    """decode the staged streams -> _native.KmerSet"""
    return _native.KmerSet.from_staged(toks[0], toks[1] if len(toks) > 1 else None, device=device)


def readKmerSet(z, nm=None, counts=True, device=0):
    """Device-resident form of readKmersAndCounts / readKmers: -> _native.KmerSet"""
    return finishKmerSet(stageKmerSet(z, nm, counts, device), device)


def readKmerSetFiles(paths, counts=True, device=0, ahead=2):
    """the sets of several k-mer set files, in order: yields (KmerSet, metadata dict).  While one file is decoded the
    streams of the next `ahead` files are already on their way to the device -- the reference decodes file after file
    (and `zot dist` re-decodes file j for every pair, dist.py:153-159)."""
    from zotmer_b200.library.kmers import kmers
    pending = []
    it = iter(paths)

    def start():
        for fn in it:
            z = kmers(fn, 'r')
            try:
                pending.append((z, stageKmerSet(z, None, counts, device)))
            except BaseException:
                z.close()
                raise
            return True
        return False

    try:
        while len(pending) < ahead + 1 and start():
            pass
        while pending:
            z, toks = pending.pop(0)
            try:
                s = finishKmerSet(toks, device)
                meta = dict(z.meta)
            finally:
                z.close()
            start()
            yield s, meta
    finally:
        for (z, toks) in pending:
            for t in toks:
                t.free()
            z.close()


def writeKmerSet(z, kset, nm=None):
    """files.py:195-217 (writeKmersAndCounts / writeKmersAndCounts2) from a device-resident set."""
    xNm, cNm = _names(nm)
    # codec64 + delta on the device; the packed words then go device -> pinned ring -> pwrite() on the library's I/O
    # threads, both streams at once, straight to their places in the container (no host copy of the streams in Python)
    w = kset.encode_dev()
    try:
        nk, nc = w.sizes()
        z._writable()
        z.fo.flush()
        at = z._end()
        w.write_fd(z.fo.fileno(), at, at + 8 * nk)
        z._record(xNm, at, 8 * nk)
        z._record(cNm, at + 8 * nk, 8 * nc)
    finally:
        w.free()


def writeKmersAndCounts2(z, xs, cs, nm=None, device=0):
    """files.py:209-217 from host arrays."""
    xNm, cNm = _names(nm)
    with z.add_stream(xNm) as f:
        writeWords(f, _native.encode_stream(xs, delta=True, device=device))
    with z.add_stream(cNm) as f:
        writeWords(f, _native.encode_stream(cs, delta=False, device=device))


writeKmersAndCounts = writeKmersAndCounts2
