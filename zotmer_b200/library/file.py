"""
Sequence-file access for the host layer (mirrors zotmer/library/file.py:79-160).

The reference parses FASTA/FASTQ text in Python (readFasta/readFastq, file.py:19-52); here the raw
bytes go to the device and libzot_b200 parses them, so this module only opens files (plain, '-' =
stdin, .gz/.bz2 through `gunzip -c` / `bunzip2 -c` exactly as file.py:93-102) and keeps the
temp-file helpers.  `readFasta` remains for the one place that needs record NAMES on the host
(`zot jaccard` in single-FASTA mode).
"""
import os
import subprocess
import sys
import uuid

PY2_SPACE = b" \t\n\r\x0b\x0c"


def readBytes(fn):
    """Whole content of a sequence file as bytes (openFile(fn).read() of the reference)."""
    if fn == "-":
        return sys.stdin.buffer.read()
    if fn.endswith(".gz") and os.path.exists(fn):
        return subprocess.run(['gunzip', '-c', fn], stdout=subprocess.PIPE, check=False).stdout
    if fn.endswith(".bz2") and os.path.exists(fn):
        return subprocess.run(['bunzip2', '-c', fn], stdout=subprocess.PIPE, check=False).stdout
    with open(fn, "rb") as f:
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
