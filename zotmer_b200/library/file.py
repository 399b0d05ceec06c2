"""
Sequence-file access for the host layer (mirrors zotmer/library/file.py:79-160).

The reference parses FASTA/FASTQ text in Python (readFasta/readFastq, file.py:19-52); here the raw
bytes go to the device and libzot_b200 parses them, so this module only opens files (plain, '-' =
stdin, .gz/.bz2 inflated in-process -- BGZF members in parallel -- with the result `gunzip -c` / `bunzip2 -c` of
file.py:93-102 would give) and keeps the
temp-file helpers.  `readFasta` remains for the one place that needs record NAMES on the host
(`zot jaccard` in single-FASTA mode).
"""
import os
import subprocess
import sys
import uuid

PY2_SPACE = b" \t\n\r\x0b\x0c"


def _bgzf_blocks(data):
    """[(start, end)] of the members of a BGZF file (bgzip: every gzip member carries its own compressed size in a
    'BC' extra field, so the members can be found without inflating anything), or None if `data` is not BGZF."""
    blocks = []
    p, n = 0, len(data)
    while p < n:
        if n - p < 18 or data[p:p + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = data[p + 10] | (data[p + 11] << 8)
        q, end, bsize = p + 12, p + 12 + xlen, None
        while q + 4 <= end:
            slen = data[q + 2] | (data[q + 3] << 8)
            if data[q:q + 2] == b"BC" and slen == 2:
                bsize = (data[q + 4] | (data[q + 5] << 8)) + 1
            q += 4 + slen
        if bsize is None or p + bsize > n:
            return None
        blocks.append((p, p + bsize, 12 + xlen))
        p += bsize
    return blocks


def gunzipBytes(data, threads=None):
    """`gunzip -c` of a whole file held in memory (file.py:93-97 pipes through the gunzip binary).  All members are
    inflated and concatenated; a BGZF file (bgzip) is inflated member by member on a thread pool -- zlib releases the
    GIL -- so that a compressed FASTQ does not trickle in at one core's inflate rate.  Like the
    reference, which never looks at gunzip's exit status, a truncated or corrupt tail just ends the data."""
    import zlib
    blocks = _bgzf_blocks(data) if len(data) >= (1 << 20) else None
    if blocks is not None and len(blocks) > 1:
        from concurrent.futures import ThreadPoolExecutor
        view = memoryview(data)

        def inflate(b):
            return zlib.decompress(view[b[0] + b[2]:b[1] - 8], -15)
        try:
            with ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 1)) as ex:
                return b"".join(ex.map(inflate, blocks, chunksize=64))
        except zlib.error:
            pass   # not what it claimed to be: member by member below
    out = []
    rest = data
    while len(rest) >= 10 and rest[:2] == b"\x1f\x8b":
        d = zlib.decompressobj(31)
        try:
            out.append(d.decompress(rest))
        except zlib.error:
            break
        if not d.eof:
            break
        rest = d.unused_data
    return b"".join(out)


def bunzip2Bytes(data):
    """`bunzip2 -c` of a whole file held in memory (file.py:98-102), concatenated streams included."""
    import bz2
    out = []
    rest = data
    while rest[:3] == b"BZh":
        d = bz2.BZ2Decompressor()
        try:
            out.append(d.decompress(rest))
        except (OSError, ValueError):
            break
        if not d.eof:
            break
        rest = d.unused_data
    return b"".join(out)


def readBytes(fn):
    """Whole content of a sequence file as bytes (openFile(fn).read() of the reference); .gz / .bz2 are decompressed
    in-process instead of through a `gunzip -c` / `bunzip2 -c` child and its pipe."""
    if fn == "-":
        return sys.stdin.buffer.read()
    with open(fn, "rb") as f:
        data = f.read()
    if fn.endswith(".gz"):
        return gunzipBytes(data)
    if fn.endswith(".bz2"):
        return bunzip2Bytes(data)
    return data


def mapBytes(fn):
    """readBytes for the device parser: a plain regular file is mapped read-only instead of being read into a fresh
    bytes object (no copy and no page faults of a second 315 MB buffer: the H2D copy reads the page cache itself --
    SURVEY.md 8f row 3).  The result supports len(), rfind() and the buffer protocol, which is all `pieces` and
    `Kmerizer.feed` need; anything else (stdin, compressed, empty, not mappable) comes back as bytes."""
    if fn == "-" or fn.endswith(".gz") or fn.endswith(".bz2"):
        return readBytes(fn)
    import mmap
    with open(fn, "rb") as f:
        try:
            size = os.fstat(f.fileno()).st_size
            if size == 0:
                return b""
            # up to 2 GiB the page tables are filled in one go; a larger file is faulted in piece by piece as it is fed
            populate = getattr(mmap, "MAP_POPULATE", 0) if size <= (2 << 30) else 0
            return mmap.mmap(f.fileno(), 0, flags=mmap.MAP_SHARED | populate, prot=mmap.PROT_READ)
        except (OSError, ValueError):
            return f.read()


def readFasta(data):
    """(name, sequence) pairs of FASTA text -- file.py:19-36 semantics on bytes."""
    nm = None
    seq = []
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for l in lines:
        l = l.strip(PY2_SPACE)
        if len(l) and l[0:1] == b'>':
            if nm is not None:
                yield (nm, b''.join(seq))
            nm = l[1:].strip(PY2_SPACE)
            seq = []
        else:
            seq.append(l)
    if nm is not None:
        yield (nm, b''.join(seq))


_tmpfiles = []


class _AutoRemover:
    def __init__(self):
        _tmpfiles.append(set([]))

    def __enter__(self):
        return None

    def __exit__(self, _t, _v, _tb):
        assert len(_tmpfiles) > 0
        for fn in _tmpfiles.pop():
            if os.path.isfile(fn):
                os.remove(fn)


def autoremove():
    """file.py:138-147"""
    return _AutoRemover()


def tmpfile(suffix=''):
    """file.py:149-160"""
    fn = os.getenv('TMPDIR', '/tmp') + '/' + str(uuid.uuid4()) + suffix
    if len(_tmpfiles):
        _tmpfiles[-1].add(fn)
    return fn
