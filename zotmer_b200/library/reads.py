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
            import numpy as np
            a = np.frombuffer(mv[start:end], dtype=np.uint8)
            nl = np.flatnonzero(a == 10)
            # newline number q (1-based, counted from `start`) ends a record when q % 4 == 0
            usable = (len(nl) // 4) * 4
            if usable == 0:
                raise ValueError("FASTQ record larger than %d bytes cannot be fed in pieces" % max_piece)
            cut = start + int(nl[usable - 1]) + 1
        yield mv[start:cut]
        start = cut
    if start < n:
        yield mv[start:]
