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
            # window (64 MiB at a time -- no index array of a gigabyte of text), then step back over the q % 4 newlines
            # of the incomplete last record
            import numpy as np
            total = 0
            for o in range(start, end, 1 << 26):
                total += int(np.count_nonzero(np.frombuffer(mv[o:min(o + (1 << 26), end)], dtype=np.uint8) == 10))
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
