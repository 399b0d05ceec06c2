"""
Input-type rules (mirrors zotmer/library/reads.py:11-33) and record-aligned splitting of large
inputs for the device parser.
"""
compressionSuffixes = ['.gz', '.bz2']

MAX_PIECE = 1 << 30  # bytes per zb_kmerize_feed call (the C ABI accepts < 2^31)


def stripCompressionSuffix(nm):
    for suff in compressionSuffixes:
        if nm.endswith(suff):
            return nm[:-len(suff)]
    return nm


def isFasta(nm):
    bnm = stripCompressionSuffix(nm)
    return bnm.endswith((".fa", ".fasta", ".fas", ".fna"))


def pieces(data, is_fasta, max_piece=MAX_PIECE):
    """Split file bytes into pieces that each parse exactly like a whole file (record aligned).

    FASTA: cut in front of a b'\\n>' header line.  FASTQ: cut after a newline whose line number is a
    multiple of 4.  A single record larger than max_piece cannot be split."""
    n = len(data)
    if n <= max_piece:
        yield data
        return
    mv = memoryview(data)
    start = 0
    lines_before = 0
    while n - start > max_piece:
        end = start + max_piece
        if is_fasta:
            cut = data.rfind(b"\n>", start + 1, end)
            if cut <= start:
                raise ValueError("FASTA record larger than %d bytes cannot be fed in pieces" % max_piece)
            cut += 1
        else:
            # newline number q (1-based, counted from `start`) ends a record when q % 4 == 0: count the newlines of the
            # window, then step back over the q % 4 newlines of the incomplete last record
            from zotmer_b200 import _native
            total = _native.host_count_byte(mv[start:end], 10)      # on the library's I/O threads
            if total < 4:
                raise ValueError("FASTQ record larger than %d bytes cannot be fed in pieces" % max_piece)
            at = end
            for _ in range(total % 4 + 1):
                at = data.rfind(b"\n", start, at)
            cut = at + 1
        yield mv[start:cut]
        start = cut
    if start < n:
        yield mv[start:]


_WS = b" \t\n\r\x0b\x0c"      # what py2's strip() removes (file.py:24,30)


def _isHeaderAt(data, at, lo):
    """data[at] == '>': is it the first non-blank byte of its line (readFasta's test for a header, file.py:24-26)?"""
    j = at - 1
    while j >= lo and data[j:j + 1] != b"\n":
        if data[j:j + 1] not in (b" ", b"\t", b"\r", b"\x0b", b"\x0c"):
            return False
        j -= 1
    return True


def _headerIn(data, lo, hi):
    """does data[lo:hi] hold a header line?"""
    at = data.find(b">", lo, hi)
    while at >= 0:
        if _isHeaderAt(data, at, lo):
            return True
        at = data.find(b">", at + 1, hi)
    return False


def _seqTail(region, m, starts_line):
    """the last m characters of the sequence readFasta has joined so far when it has read `region` (a stretch of a
    record's lines; starts_line: it begins at the start of a line).  None if the region ends in a header line."""
    lines = region.split(b"\n")
    seq = []
    header = False
    for i, l in enumerate(lines):
        t = l.strip(_WS) if (i > 0 or starts_line) else l.rstrip(_WS)
        if i == len(lines) - 1 and l and l[-1:] not in (b" ", b"\t", b"\r", b"\x0b", b"\x0c"):
            t = l.lstrip(_WS) if (i > 0 or starts_line) else l       # the line goes on behind the cut: nothing trails yet
        if (i > 0 or starts_line) and t[:1] == b">":
            seq = []
            header = True
            continue
        header = False
        seq.append(t)
    if header:
        return None
    joined = b"".join(seq[-(m + 2):]) if m else b""
    if len(joined) < m:
        joined = b"".join(seq)
    return joined[max(0, len(joined) - m):] if m else b""


def splitPieces(data, is_fasta, k, max_piece=MAX_PIECE):
    """`pieces` that never gives up on a long FASTA record: -> (prefix, piece, fake).  A record that fills a whole window
    is cut inside (after a line, or inside a line that is itself longer than the window); the next piece then starts
    with `prefix` = an empty header line + the last k - 1 characters of the sequence so far, so that exactly the
    windows that span the cut are seen there (the reference streams a record of any length, file.py:19-36).  `fake` = 1
    for such a piece: it adds a record that is not one."""
    n = len(data)
    if n <= max_piece or not is_fasta:
        for p in pieces(data, is_fasta, max_piece):
            yield (b"", p, 0)
        return
    mv = memoryview(data)
    start = 0
    prefix, fake = b"", 0
    in_record = False          # a header line lies before `start`
    look = max(4096, 64 * k)
    while n - start > max_piece:
        end = start + max_piece
        cut = data.rfind(b"\n>", start + 1, end)
        if cut > start:
            cut += 1
            yield (prefix, mv[start:cut], fake)
            prefix, fake, start, in_record = b"", 0, cut, True
            continue
        nl = data.rfind(b"\n", start + 1, end)
        if nl > start:
            p, midline = nl + 1, False
        else:
            p, midline = end, True
            # not next to white space (it would become leading / trailing and vanish) and not in front of a '>' (it would
            # become the first byte of a line: a header)
            while p > start + 2 and (data[p - 1:p] in (b" ", b"\t", b"\r", b"\x0b", b"\x0c") or
                                     data[p:p + 1] in (b" ", b"\t", b"\r", b"\x0b", b"\x0c", b">")):
                p -= 1
        in_record = in_record or _headerIn(data, start, p)
        yield (prefix, mv[start:p], fake)
        if in_record:
            lo = max(start, p - look)
            region = bytes(mv[lo:p])
            starts_line = lo == start or data[lo - 1:lo] == b"\n"
            if lo == start and prefix:
                region, starts_line = prefix + region, True
            tail = _seqTail(region, k - 1, starts_line)
            if tail is None:
                # the cut follows a header line: that header opens an empty record in the piece just given out, and
                # the sequence needs a header of its own here
                prefix, fake = b">\n", 1
            else:
                if tail[:1] == b">":        # a '>' from inside a line must not open the line here (any letter that is
                    tail = b"-" + tail      # no base does in front of it: no window starts there)
                prefix, fake = b">\n" + tail + (b"" if midline else b"\n"), 1
        else:
            prefix, fake = b"", 0
        start = p
    if start < n:
        yield (prefix, mv[start:], fake)


def stagedPieces(inputs, device=0, verbose=False, max_piece=MAX_PIECE, k=None):
    """(staged piece, is_fasta) for every record-aligned piece of every input file, in order; the copy of the NEXT piece
    to the device has already been started (library I/O threads, pinned ring) when a piece is handed out, so reading /
    copying piece i + 1 overlaps the parsing and extraction of piece i.  A plain file that fits one piece is read by the
    I/O threads themselves (pread into pinned chunks: no mapping, no Python bytes object); anything else (stdin, .gz /
    .bz2, files beyond one piece) is staged from host memory.  k: the k-mer length, for callers that count k-mers -- a
    FASTA record longer than a piece is then cut inside instead of refused; such a piece counts one record too many
    (`fake_records` of the staged piece)."""
    import os
    import sys
    from zotmer_b200 import _native
    from zotmer_b200.library.file import mapBytes

    def jobs():
        for fn in inputs:
            held = None
            if isinstance(fn, tuple):      # (name, content already in memory): an input that can only be read once
                fn, held = fn
            fa = isFasta(fn)
            plain = held is None and fn != '-' and not fn.endswith(('.gz', '.bz2')) and os.path.isfile(fn)
            size = os.path.getsize(fn) if plain else -1
            if verbose:
                print('reading %s (%s)' % (fn, 'FASTA' if fa else 'FASTQ'), file=sys.stderr)
            if plain and 0 < size <= max_piece:
                yield ('fd', fn, size, fa)
            elif held is None and fn.endswith('.gz') and os.path.isfile(fn) and _bgzf(fn) is not None:
                yield ('bgzf', _bgzf(fn), 0, fa)      # block-compressed: inflated on the device, group by group
            else:
                data = held if held is not None else mapBytes(fn)
                if k is not None:
                    # a FASTA record longer than a piece is cut inside (splitPieces): the piece behind the cut starts with
                    # an empty header line and the k - 1 characters in front of the cut
                    for (prefix, piece, fake) in splitPieces(data, fa, k, max_piece):
                        if len(piece):
                            yield ('mem', (prefix, piece, fake), len(piece), fa)
                else:
                    for piece in pieces(data, fa, max_piece):
                        if len(piece):
                            yield ('mem', piece, len(piece), fa)

    def start(job):
        kind, src, n, fa = job
        if kind == 'fd':
            f = open(src, 'rb')
            st = _native.stage_fd(f.fileno(), 0, n, device)
            st.keep = f            # the descriptor stays open until the piece has been fed
            return st, fa
        if isinstance(src, tuple):
            (prefix, piece, fake) = src
            st = _native.stage_input(piece, device)
            if prefix:
                keep = st.keep
                st = _native.stage_concat(prefix, st, device)     # waits for the copy of `piece`
                st.keep = keep
            st.fake_records = fake
            return st, fa
        return _native.stage_input(src, device), fa

    it = jobs()
    cur = None
    for job in it:
        if job[0] == 'bgzf':
            if cur is not None:
                yield cur
                cur = None
            for st in bgzfPieces(job[1], job[3], device, max_piece):
                yield st, job[3]
            continue
        nxt = start(job)
        if cur is not None:
            yield cur
        cur = nxt
    if cur is not None:
        yield cur


BGZF_GROUP = 256 << 20   # bytes of text inflated per launch


def bgzfPieces(comp, is_fasta, device=0, max_piece=MAX_PIECE, group=None):
    """Staged pieces of a BGZF file (`comp`: the compressed bytes -- a mapping of the .gz file).  The members are inflated
    on the device, `group` bytes of text at a time (library zb_stage_bgzf: the compressed bytes cross PCIe, one warp per
    member); a piece ends at the last record boundary of its group (found on the device, zb_staged_cut) and the
    incomplete record behind it is carried to the front of the next piece -- what `pieces` does for text in host memory.
    The inflate of group i + 1 has been issued when piece i is handed out."""
    import numpy as np
    from zotmer_b200 import _native
    group = group or min(BGZF_GROUP, max_piece)
    a = np.frombuffer(comp, dtype=np.uint8)
    n, off = len(a), 0

    def damaged(rest, carry):
        """a member that does not inflate: what `gunzip -c` would still have written before it gave up -- the reference
        never looks at gunzip's exit status (file.py:93-97) -- inflated on the host, behind the carried bytes"""
        from zotmer_b200.library.file import gunzipBytes
        text = carry + gunzipBytes(bytes(rest))
        return _native.stage_input(text, device) if text else None

    try:
        st, used = _native.stage_bgzf(a, device, group)
    except AssertionError:
        last = damaged(a, b"")
        if last is not None:
            yield last
        return
    off += used
    while off < n:
        cut = st.cut(is_fasta)
        carried = len(st) - cut
        # cut == 0: one record fills the whole piece so far -- it grows by another group (up to the parser's limit)
        try:
            nxt, used = _native.stage_bgzf(a[off:], device, carried + group, st, cut)
        except AssertionError:
            last = damaged(a[off:], st.fetch_range(cut, carried))
            if cut:
                st.set_len(cut)
                yield st
            else:
                st.free()
            if last is not None:
                yield last
            return
        off += used
        if cut:
            st.set_len(cut)
            yield st
        else:
            st.free()
        st = nxt
    yield st


_bgzf_cache = {}


def _bgzf(fn):
    """the mapped bytes of `fn` if it is a BGZF file, else None"""
    import os
    key = (fn, os.path.getmtime(fn), os.path.getsize(fn))
    if key not in _bgzf_cache:
        import mmap
        from zotmer_b200 import _native
        _bgzf_cache.clear()
        res = None
        if key[2] >= 28:
            with open(fn, 'rb') as f:
                m = mmap.mmap(f.fileno(), 0, prot=mmap.PROT_READ)
            if bytes(m[:4]) == b"\x1f\x8b\x08\x04" and _native.bgzf_probe(m) is not None:
                res = m
        _bgzf_cache[key] = res
    return _bgzf_cache[key]
