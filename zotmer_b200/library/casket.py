"""
The casket container (mirrors zotmer/library/container/casket.py:52-234, byte for byte on disk):

    blob ... blob | TOC as JSON {name: [[offset, length], ...]} | u64 little-endian length of the TOC

`open(name)` reads the LAST entry of a name (casket.py:185).  JSON is written with the default
separators and in insertion order (the PyPy2 / Python 3 layout).
"""
import json
import os
import struct


class MultipleOpenFiles(Exception):
    def __init__(self):
        super(MultipleOpenFiles, self).__init__('cannot add to archive while streaming object open')


_block_size_ = 1024 * 1024


class CasketReader(object):
    def __init__(self, fo, p, l):
        self.fo, self.p, self.l, self.o = fo, p, l, 0

    def read(self, z=None):
        if z is None:
            z = self.l - self.o
        z = min(z, self.l - self.o)
        if z == 0:
            return b''
        self.fo.seek(self.p + self.o, os.SEEK_SET)
        w = self.fo.read(z)
        assert len(w) == z
        self.o += z
        return w


class CasketStreamWriter(object):
    def __init__(self, ar, afn):
        self.ar, self.afn = ar, afn
        self.ar.fo.seek(0, os.SEEK_END)
        self.p = self.ar.fo.tell()
        self.l = 0
        self.closed = False

    def write(self, dat):
        self.l += len(dat)
        self.ar.fo.write(dat)

    def close(self):
        assert not self.closed
        self.ar.updateToc(self.afn, self.p, self.l)
        self.ar.fip = None
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, t, v, tb):
        if t is not None:
            return False
        if not self.closed:
            self.close()
        return True


class casket(object):
    def __init__(self, fn, mode='r'):
        self.fn = fn
        self.mode = mode
        self.toc = {}
        self.fip = None
        if mode == 'r':
            self.fo = open(fn, 'rb')
            self._readToc()
            self.stale = False
        elif mode == 'w':
            self.fo = open(fn, 'wb')
            self.stale = True
        else:
            raise ValueError(mode)

    def list(self):
        return [(nm, ys[-1][1]) for (nm, ys) in sorted(self.toc.items())]

    def add_file(self, afn, fn):
        assert self.mode == 'w'
        if self.fip is not None:
            raise MultipleOpenFiles
        self.fo.seek(0, os.SEEK_END)
        p = self.fo.tell()
        l = 0
        with open(fn, 'rb') as f:
            w = f.read(_block_size_)
            while len(w) > 0:
                l += len(w)
                self.fo.write(w)
                w = f.read(_block_size_)
        self.updateToc(afn, p, l)

    def add_content(self, afn, data):
        assert self.mode == 'w'
        if self.fip is not None:
            raise MultipleOpenFiles
        if isinstance(data, str):
            data = data.encode('latin-1')
        self.fo.seek(0, os.SEEK_END)
        p = self.fo.tell()
        self.fo.write(data)
        self.updateToc(afn, p, len(data))

    def add_stream(self, afn):
        assert self.mode == 'w'
        if self.fip is not None:
            raise MultipleOpenFiles
        self.fip = CasketStreamWriter(self, afn)
        return self.fip

    def open(self, afn):
        assert self.mode == 'r'
        (p, l) = self.toc[afn][-1]
        return CasketReader(self.fo, p, l)

    def updateToc(self, afn, p, l):
        self.toc.setdefault(afn, []).append((p, l))
        self.stale = True

    def flush(self):
        pass

    def close(self):
        if self.fo is None:
            return
        if self.stale:
            self.flush()
            self._writeToc()
        self.fo.close()
        self.fo = None
        self.stale = False

    def abandon(self):
        """Close WITHOUT a table of contents: what the reference leaves behind when a command
        raises inside `with casket(...)` (casket.py:213-217 returns before close())."""
        if self.fo is not None:
            self.fo.close()
            self.fo = None

    def __enter__(self):
        return self

    def __exit__(self, t, v, tb):
        if t is not None:
            self.abandon()
            return False
        self.close()
        return True

    def _readToc(self):
        self.fo.seek(-8, os.SEEK_END)
        z = struct.unpack('<Q', self.fo.read(8))[0]
        self.fo.seek(-(8 + z), os.SEEK_END)
        self.toc = json.loads(self.fo.read(z))

    def _writeToc(self):
        w = json.dumps(self.toc).encode('latin-1')
        self.fo.seek(0, os.SEEK_END)
        self.fo.write(w)
        self.fo.write(struct.pack('<Q', len(w)))
        self.fo.flush()
