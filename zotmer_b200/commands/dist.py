"""
Usage:
    zot dist [-M measure]... <k> <input>...

Options:
    -M measure  use "measure" for the distance between k-mer frequency sets.
                Use "-M list" to get a list of available measures.
"""
# Drop-in for zotmer/commands/dist.py:94-168.  The reference decodes file j again for every pair and
# runs the two-pointer split() once per measure per pair (:145-168, library/dist.py:241-265); here
# every file is decoded once, projected to K (zb_project = Measure.prep :29-49) and kept on the
# device, all (i<j) pairs go through zb_pairs_abc in one batch, and the measures are evaluated on the
# host from (a,b,c) with the reference's float formulas.
import fnmatch
import sys

import numpy as np

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import _native
import zotmer_b200.library.dist as dist
from zotmer_b200.library.exceptions import MismatchedK
from zotmer_b200.library.files import readKmerSet
from zotmer_b200.library.kmers import kmers

measures = {}


class Measure:
    def __init__(self, name, desc, vec, func):
        self.name = name
        self.desc = desc
        self.vec = vec
        self.func = func

    def prep(self, K, fn, device=0):
        """commands/dist.py:29-49 (set form): project to K-mers, drop adjacent duplicates -> KmerSet"""
        with kmers(fn, 'r') as z:
            fK = z.meta['K']
            if fK < K:
                raise MismatchedK(K, fK)
            if self.vec:
                # the reference unpacks (x, c) from a stream of plain ints here (:34-41)
                raise TypeError("cannot unpack non-iterable int object")
            xs = readKmerSet(z, counts=False, device=device)
            S = 2 * (fK - K)
            v = xs.project(S)
            xs.free()
            return v

    def measure(self, a, b, c):
        return self.func(a, b, c)


def addMeasure(name, desc, vec, func):
    measures[name] = Measure(name, desc, vec, func)


addMeasure('bray.curtis.quant', 'Quantative Bray.Curtis distance', True, None)
addMeasure('bray.curtis.qual', 'Qualitative Bray.Curtis distance', False, dist.brayCurtis)
addMeasure('chord.quant', 'Quantative Chord distance', True, None)
addMeasure('chord.qual', 'Qualitative Chord distance', False, dist.chord)
addMeasure('hellinger.quant', 'Quantative Hellinger distance', True, None)
addMeasure('hellinger.qual', 'Qualitative Hellinger distance', False, dist.hellinger)
addMeasure('jaccard.ab', 'Abundance.based Jaccard distance', True, None)
addMeasure('jaccard.qual', 'Qualitative Jaccard distance', False, dist.jaccard)
addMeasure('jensen.shannon', 'Jensen.Shannon distance', True, None)
addMeasure('kulczynski.quant', 'Quantative Kulczynski distance', True, None)
addMeasure('kulczynski.qual', 'Qualitative Kulczynski distance', False, dist.kulczynski)
addMeasure('ochiai.ab', 'Abundance.based Ochiai distance', True, None)
addMeasure('ochiai.qual', 'Qualitative Ochiai distance', False, dist.ochiai)
addMeasure('sorensen.ab', 'Abundance.based Sorensen distance', True, None)
addMeasure('sorensen.qual', 'Qualitative Sorensen distance', False, dist.sorensen)
addMeasure('whittaker.quant', 'Quantative Whittaker distance', True, None)
addMeasure('whittaker.qual', 'Qualitative Whittaker distance', False, dist.whittaker)


def allPairsABC(sets):
    """(a,b,c) for every i<j in row-major order -> uint64 [npairs, 3]"""
    N = len(sets)
    (I, J) = np.triu_indices(N, 1)
    return (I, J, _native.allpairs_abc(sets))


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    if "list" in opts['-M']:
        msg = []
        for m in sorted(measures.keys()):
            msg.append(m + '\t' + measures[m].desc)
        print('\n'.join(msg))
        return

    allMs = sorted(measures.keys())

    seen = set([])
    bad = False
    for mo in opts['-M']:
        found = False
        for m in allMs:
            if fnmatch.fnmatch(m, mo):
                seen.add(m)
                found = True
        if not found:
            print('warning: measure \'%s\' not found. Use -M list to see all measures.' % (mo,), file=sys.stderr)
            bad = True
    ms = sorted(seen)

    if len(ms) == 0 or bad:
        return

    K = int(opts['<k>'])

    fns = opts['<input>']
    N = len(fns)

    hdr = ['lhs.name', 'rhs.name']
    fmt = ['%s', '%s']
    vecNeeded = None
    setNeeded = None
    for m in ms:
        hdr.append(m)
        fmt.append('%g')
        if measures[m].vec:
            vecNeeded = m
        else:
            setNeeded = m
    fmt = '\t'.join(fmt)

    print('\t'.join(hdr))
    sys.stdout.flush()
    if N == 0:
        return
    if vecNeeded is not None:
        measures[vecNeeded].prep(K, fns[0])     # TypeError, as the reference's first prep call
    sets = [measures[setNeeded].prep(K, fn) for fn in fns]
    (I, J, abc) = allPairsABC(sets)
    out = []
    for p in range(len(I)):
        (a, b, c) = (int(abc[p, 0]), int(abc[p, 1]), int(abc[p, 2]))
        vs = [fns[I[p]], fns[J[p]]]
        for m in ms:
            vs.append(measures[m].measure(a, b, c))
        out.append(fmt % tuple(vs))
        if len(out) >= 4096:
            print('\n'.join(out))
            out = []
    if out:
        print('\n'.join(out))
    for s in sets:
        s.free()


if __name__ == '__main__':
    main(sys.argv[1:])
