# The "casket" container of the k-mer set files (on-disk format of zotmer/library/container/casket.py:52-234, byte for
# byte):
#
#     entry bytes ... entry bytes | table of contents, JSON {name: [[offset, length], ...]} | u64 LE: length of that JSON
#
# A name may occur several times (every write appends a version); readers see the LAST version (casket.py:185).  The
# table is written when a container opened for writing is closed, in insertion order and with json's default
# separators (what the reference's files look like under PyPy2 / any Python >= 3.7).  If a command dies inside
# `with casket(...)`, the file is left without a table, as the reference leaves it (casket.py:213-217).
import json
import struct

TAIL = struct.Struct('<Q')


class MultipleOpenFiles(Exception):
    """an entry was started while a streamed entry of the same container was still open"""

    def __init__(self):
        Exception.__init__(self, 'cannot add to archive while streaming object open')


class Entry(object):
    """read access to one stored entry: read() / read(n) like a file"""

    def __init__(self, fileobj, offset, length, owner=None):
        self._f = fileobj
        self._at = offset
        self._left = length
        self._owner = owner      # the casket: keeps ONE mapping of the file for all its entries

    def read(self, n=None):
        take = self._left if n is None else min(n, self._left)
        if take == 0:
            return b''
        self._f.seek(self._at)
        data = self._f.read(take)
        assert len(data) == take
        self._at += take
        self._left -= take
        return data

    def view(self):
        """the unread rest of the entry as a read-only buffer over a mapping of the file -- no copy into a bytes object
        (the word streams of a k-mer set go straight from the page cache to the device); None when the file cannot be
        mapped, the caller then read()s"""
        import mmap
        if self._left == 0:
            return b''
        m = getattr(self._owner, '_map', None)
        if m is None:
            try:
                import os
                big = os.fstat(self._f.fileno()).st_size > (2 << 30)
                m = mmap.mmap(self._f.fileno(), 0, flags=mmap.MAP_SHARED | (0 if big else getattr(mmap, 'MAP_POPULATE', 0)),
                              prot=mmap.PROT_READ)
            except (OSError, ValueError, AttributeError):
                return None
            if self._owner is not None:
                self._owner._map = m
        out = memoryview(m)[self._at:self._at + self._left]
        assert len(out) == self._left
        self._at += self._left
        self._left = 0
        return out


class EntryWriter(object):
    """an entry written piece by piece (`with z.add_stream(name) as f: f.write(...)`); it is entered into the table
    when it is closed, and nothing else may be added to the container until then"""

    def __init__(self, owner, name):
        self._owner = owner
        self._name = name
        self._start = owner._end()
        self._size = 0
        self._open = True

    def write(self, data):
        self._owner.fo.write(data)
        self._size += len(data)

    def close(self):
        assert self._open
        self._open = False
        self._owner._streaming = None
        self._owner._record(self._name, self._start, self._size)

    def __enter__(self):
        return self

    def __exit__(self, etype, value, tb):
        if etype is None and self._open:
            self.close()
        return etype is None


class casket(object):
    def __init__(self, fn, mode='r'):
        if mode not in ('r', 'w'):
            raise ValueError(mode)
        self.fn = fn
        self.mode = mode
        self.toc = {}
        self._streaming = None
        self._map = None
        self.fo = open(fn, mode + 'b')
        self._dirty = (mode == 'w')
        if mode == 'r':
            self.fo.seek(-TAIL.size, 2)
            (toc_len,) = TAIL.unpack(self.fo.read(TAIL.size))
            self.fo.seek(-(TAIL.size + toc_len), 2)
            self.toc = json.loads(self.fo.read(toc_len))

    # ---- reading
    def open(self, name):
        assert self.mode == 'r'
        offset, length = self.toc[name][-1]          # KeyError for a name that is not there
        return Entry(self.fo, offset, length, self)

    def list(self):
        """[(name, length of its latest version)] in name order"""
        return [(name, self.toc[name][-1][1]) for name in sorted(self.toc)]

    # ---- writing
    def _end(self):
        self.fo.seek(0, 2)
        return self.fo.tell()

    def _record(self, name, offset, length):
        self.toc.setdefault(name, []).append((offset, length))
        self._dirty = True

    def _writable(self):
        assert self.mode == 'w'
        if self._streaming is not None:
            raise MultipleOpenFiles

    def add_content(self, name, data):
        self._writable()
        if isinstance(data, str):
            data = data.encode('latin-1')
        at = self._end()
        self.fo.write(data)
        self._record(name, at, len(data))

    def add_file(self, name, path, block=1 << 20):
        self._writable()
        at = self._end()
        size = 0
        with open(path, 'rb') as src:
            for piece in iter(lambda: src.read(block), b''):
                self.fo.write(piece)
                size += len(piece)
        self._record(name, at, size)

    def add_stream(self, name):
        self._writable()
        self._streaming = EntryWriter(self, name)
        return self._streaming

    # ---- closing
    def close(self):
        if self.fo is None:
            return
        if self._dirty:
            toc = json.dumps(self.toc).encode('latin-1')
            self._end()
            self.fo.write(toc)
            self.fo.write(TAIL.pack(len(toc)))
        self.fo.close()
        self.fo = None
        self._dirty = False

    def abandon(self):
        """close WITHOUT writing a table of contents"""
        if self.fo is not None:
            self.fo.close()
            self.fo = None

    def __enter__(self):
        return self

    def __exit__(self, etype, value, tb):
        if etype is None:
            self.close()
        else:
            self.abandon()
        return etype is None
