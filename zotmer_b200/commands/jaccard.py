"""
Usage:
    zot jaccard [-abp P] <input>...

Compute Jaccard indexes between k-mer sets. By default, indexes are
computed only between the first k-mer set and all the remaining
k-mer sets. If the -a option is given, all pairwise indexes are
computed.  If the -p P option is given, a Null hypothesis test is
performed for the hypothesis that the underlying Jaccard Index is
less than P. This is particularly useful if subsets of k-mers are
being used (NB, if the k-mer sets are large, the statistics can be
very expensive to compute).

Options:
    -a          print all pairwise distances
    -p P        Jaccard distance thresshhold for p-value computation
"""
# Drop-in for zotmer/commands/jaccard.py:101-167.  The two-pointer jaccard() (:31-54) is
# zb_pairs_abc on the device; the beta-quantile statistics (:56-83) stay in host Python.
import math
import sys

import numpy as np

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import _native
from zotmer_b200.library.file import readBytes, readFasta
from zotmer_b200.library.files import readKmerSet
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.reads import stripCompressionSuffix
from zotmer_b200.library.stats import logAdd, logChoose


def logIx(x, m, n):
    "jaccard.py:56-70"
    lx = math.log(x)
    j = m
    v = logChoose(n + j - 1, j)
    s = v + j * lx
    while True:
        j += 1
        v += math.log((n + j - 1.0) / j)
        t = v + j * lx
        u = logAdd(s, t)
        if u == s:
            break
        s = u
    return n * math.log1p(-x) + s


def quantBeta(q, m, n):
    "jaccard.py:72-83"
    lq = math.log(q)
    l = 1e-10
    h = 1 - 1e-10
    while (h - l) > 1e-7:
        x = (h + l) / 2.0
        lp = logIx(x, m, n)
        if lp < lq:
            l = x
        else:
            h = x
    return l


def isFasta(nm):
    "jaccard.py:90-99 (only .gz is stripped here)"
    bnm = nm[:-3] if nm.endswith('.gz') else nm
    return bnm.endswith((".fa", ".fasta", ".fas", ".fna"))


def _line(xnm, ynm, xz, yz, isec, union, p):
    d = float(isec) / float(union)
    if p is None:
        return '%s\t%s\t%d\t%d\t%d\t%d\t%f' % (xnm, ynm, xz, yz, isec, union, d)
    pv = logIx(p, isec + 1, (union - isec) + 1) / math.log(10)
    q05 = quantBeta(0.05, isec + 1, (union - isec) + 1)
    q95 = quantBeta(0.95, isec + 1, (union - isec) + 1)
    return '%s\t%s\t%d\t%d\t%d\t%d\t%f\t-%f\t+%f\t%f' % (xnm, ynm, xz, yz, isec, union, d, d - q05, q95 - d, pv)


def _report(names, sets, pairs, p):
    I = np.array([i for (i, j) in pairs], dtype=np.uint32)
    J = np.array([j for (i, j) in pairs], dtype=np.uint32)
    n = len(sets)
    if n > 2 and len(pairs) == n * (n - 1) // 2:
        # -a over every input: the tiled all-pairs kernel (same row-major pair order)
        abc = _native.allpairs_abc(sets)
    else:
        abc = _native.pairs_abc(sets, I, J)
    sizes = [len(s) for s in sets]
    for q in range(len(pairs)):
        (i, j) = pairs[q]
        isec = int(abc[q, 0])
        union = isec + int(abc[q, 1]) + int(abc[q, 2])
        print(_line(names[i], names[j], sizes[i], sizes[j], isec, union, p))
        sys.stdout.flush()


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    fns = opts['<input>']

    p = None
    if opts['-p'] is not None:
        p = float(opts['-p'])

    if len(fns) == 1 and isFasta(fns[0]):
        # one k-mer set (K=25, both strands) per FASTA record -- jaccard.py:110-142
        K = 25
        names = []
        sets = []
        for (nm, seq) in readFasta(readBytes(fns[0])):
            km = _native.Kmerizer(K)
            km.feed(b'>r\n' + seq + b'\n', True)
            (s, _) = km.finish()
            km.close()
            names.append(nm.split()[0].decode('latin-1'))
            sets.append(s)
        Z = 1
        if opts['-a']:
            Z = len(sets)
        print(len(sets))
        pairs = [(i, j) for i in range(Z) for j in range(i + 1, len(sets))]
        _report(names, sets, pairs, p)
        return

    Z = 1
    if opts['-a']:
        Z = len(fns)

    Ks = []
    sets = []
    for fn in fns:
        with kmers(fn, 'r') as z:
            Ks.append(z.meta['K'])
            sets.append(readKmerSet(z, counts=False))
    pairs = []
    bad = None
    for i in range(Z):
        for j in range(i + 1, len(fns)):
            if Ks[i] != Ks[j]:
                bad = fns[j]
                break
            pairs.append((i, j))
        if bad is not None:
            break
    _report(fns, sets, pairs, p)
    if bad is not None:
        print('mismatched K:', bad, file=sys.stderr)
        sys.exit(1)


if __name__ == '__main__':
    main(sys.argv[1:])
