# `zot jaccard` (zotmer/commands/jaccard.py:85-167): sizes, intersection, union and Jaccard index of pairs of k-mer sets
# -- the first input against all the others, or all pairs with -a; with -p P also a 90 % interval and the log10
# probability that the index is below P.  A single FASTA input is treated as one set (K = 25, both strands) per record
# (:110-142).  The reference walks two sorted arrays per pair (jaccard(), :31-54); here the sets live on the device and
# (both, only left, only right) come from zb_allpairs_abc / zb_pairs_abc; the statistics are host floats
# (library/stats.py).  Kept: every line is flushed as it is printed; a pair of files with different K stops the run
# with "mismatched K: <file>" and exit status 1 after the pairs before it were printed.
import math
import sys

import numpy as np

from zotmer_b200 import _native
from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library.file import readBytes, readFasta
from zotmer_b200.library.files import readKmerSet, readKmerSetFiles
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.stats import betaQuantile, logBetaSeries

__doc__ = usage.JACCARD
FASTA_K = 25


def isFasta(name):
    """commands/jaccard.py:90-99: the suffix rule of THIS command (only .gz is looked through)"""
    stem = name[:-3] if name.endswith('.gz') else name
    return stem.endswith(('.fa', '.fasta', '.fas', '.fna'))


def logIx(x, m, n):
    return logBetaSeries(x, m, n)


def quantBeta(q, m, n):
    return betaQuantile(q, m, n)


def describe(left, right, nleft, nright, both, either, p):
    """one output line"""
    index = float(both) / float(either)
    fields = '%s\t%s\t%d\t%d\t%d\t%d\t%f' % (left, right, nleft, nright, both, either, index)
    if p is None:
        return fields
    hits, misses = both + 1, (either - both) + 1
    below = logBetaSeries(p, hits, misses) / math.log(10)
    q05 = betaQuantile(0.05, hits, misses)
    q95 = betaQuantile(0.95, hits, misses)
    return fields + '\t-%f\t+%f\t%f' % (index - q05, q95 - index, below)


def cardinalities(sets, pairs):
    """(both, only left, only right) per listed pair; the all-pairs kernel when the list is the whole upper triangle"""
    n = len(sets)
    if n > 2 and len(pairs) == n * (n - 1) // 2:
        return _native.allpairs_abc(sets)          # same row-major pair order
    left = np.array([i for (i, _) in pairs], dtype=np.uint32)
    right = np.array([j for (_, j) in pairs], dtype=np.uint32)
    return _native.pairs_abc(sets, left, right)


def report(names, sets, pairs, p):
    abc = cardinalities(sets, pairs)
    for row, (i, j) in enumerate(pairs):
        both = int(abc[row, 0])
        either = both + int(abc[row, 1]) + int(abc[row, 2])
        print(describe(names[i], names[j], len(sets[i]), len(sets[j]), both, either, p))
        sys.stdout.flush()


def recordSets(path):
    """one (name, both-strand 25-mer set) per record of a FASTA file"""
    names, sets = [], []
    for (header, seq) in readFasta(readBytes(path)):
        km = _native.Kmerizer(FASTA_K)
        km.feed(b'>r\n' + seq + b'\n', True)
        (kset, _) = km.finish()
        km.close()
        names.append(header.split()[0].decode('latin-1'))
        sets.append(kset)
    return names, sets


def main(argv):
    opts = docopt.docopt(__doc__, argv)
    paths = opts['<input>']
    p = float(opts['-p']) if opts['-p'] is not None else None

    if len(paths) == 1 and isFasta(paths[0]):
        names, sets = recordSets(paths[0])
        print(len(sets))
        firsts = len(sets) if opts['-a'] else 1
        report(names, sets, [(i, j) for i in range(firsts) for j in range(i + 1, len(sets))], p)
        return

    ks, sets = [], []
    for (xs, meta) in readKmerSetFiles(paths, counts=False):     # the next files are read and copied meanwhile
        ks.append(meta['K'])
        sets.append(xs)
    firsts = len(paths) if opts['-a'] else 1
    pairs, offender = [], None
    for i in range(firsts):
        for j in range(i + 1, len(paths)):
            if ks[i] != ks[j]:
                offender = paths[j]
                break
            pairs.append((i, j))
        if offender is not None:
            break
    report(paths, sets, pairs, p)
    if offender is not None:
        print('mismatched K:', offender, file=sys.stderr)
        sys.exit(1)


if __name__ == '__main__':
    main(sys.argv[1:])
