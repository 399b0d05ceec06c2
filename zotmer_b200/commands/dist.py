# `zot dist` (zotmer/commands/dist.py:94-168): a table of distances between every pair of k-mer sets, one column per
# selected measure.  The reference decodes file j again for every pair (i, j) and walks the two sorted arrays once per
# measure per pair (library/dist.py:241-265 split()).  Here every file is decoded once on the device and projected to
# K-mers (zb_project = Measure.prep, :29-49); ALL pairs get their (a, b, c) = (both, only left, only right) from one
# zb_allpairs_abc call; the measures are the reference's float formulas on the host (library/dist.py), so the %g
# columns agree to the last digit.
# Kept from the reference: "-M list"; unknown patterns warn and print nothing; quantitative ("vec") measures crash with
# a TypeError after the header line, because the reference unpacks (x, c) from a stream of plain ints (:34-41); a file
# whose K is below the requested K raises MismatchedK.
import fnmatch
import sys

from zotmer_b200 import _native
from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
import zotmer_b200.library.dist as formulas
from zotmer_b200.library.exceptions import MismatchedK
from zotmer_b200.library.files import readKmerSet, readKmerSetFiles
from zotmer_b200.library.kmers import kmers

__doc__ = usage.DIST

# name -> (description, function of (a, b, c) or None for the quantitative measures, which need counts)
MEASURES = dict((name, (desc, fn)) for (name, desc, fn) in [
    ('bray.curtis.quant', 'Quantative Bray.Curtis distance', None),
    ('bray.curtis.qual', 'Qualitative Bray.Curtis distance', formulas.brayCurtis),
    ('chord.quant', 'Quantative Chord distance', None),
    ('chord.qual', 'Qualitative Chord distance', formulas.chord),
    ('hellinger.quant', 'Quantative Hellinger distance', None),
    ('hellinger.qual', 'Qualitative Hellinger distance', formulas.hellinger),
    ('jaccard.ab', 'Abundance.based Jaccard distance', None),
    ('jaccard.qual', 'Qualitative Jaccard distance', formulas.jaccard),
    ('jensen.shannon', 'Jensen.Shannon distance', None),
    ('kulczynski.quant', 'Quantative Kulczynski distance', None),
    ('kulczynski.qual', 'Qualitative Kulczynski distance', formulas.kulczynski),
    ('ochiai.ab', 'Abundance.based Ochiai distance', None),
    ('ochiai.qual', 'Qualitative Ochiai distance', formulas.ochiai),
    ('sorensen.ab', 'Abundance.based Sorensen distance', None),
    ('sorensen.qual', 'Qualitative Sorensen distance', formulas.sorensen),
    ('whittaker.quant', 'Quantative Whittaker distance', None),
    ('whittaker.qual', 'Qualitative Whittaker distance', formulas.whittaker),
])


def select(patterns):
    """measure names matching the -M patterns, sorted; None when a pattern matches nothing (after the warning)"""
    names = sorted(MEASURES)
    chosen = set()
    clean = True
    for pat in patterns:
        hits = fnmatch.filter(names, pat)
        if not hits:
            print("warning: measure '%s' not found. Use -M list to see all measures." % (pat,), file=sys.stderr)
            clean = False
        chosen.update(hits)
    return sorted(chosen) if clean else None


def projected(K, path, device=0):
    """the file's k-mers cut down to their first K bases, duplicates dropped, on the device (Measure.prep, set form)"""
    with kmers(path, 'r') as z:
        fileK = z.meta['K']
        if fileK < K:
            raise MismatchedK(K, fileK)
        whole = readKmerSet(z, counts=False, device=device)
    cut = whole.project(2 * (fileK - K))
    whole.free()
    return cut


def rows(paths, abc, columns):
    """the table body: pairs in row-major order of the upper triangle, which is zb_allpairs_abc's order"""
    p = 0
    for i in range(len(paths)):
        for j in range(i + 1, len(paths)):
            a, b, c = (int(v) for v in abc[p])
            p += 1
            yield '\t'.join([paths[i], paths[j]] + ['%g' % MEASURES[m][1](a, b, c) for m in columns])


def main(argv):
    opts = docopt.docopt(__doc__, argv)
    if 'list' in opts['-M']:
        print('\n'.join(name + '\t' + MEASURES[name][0] for name in sorted(MEASURES)))
        return
    columns = select(opts['-M'])
    if not columns:
        return
    K = int(opts['<k>'])
    paths = opts['<input>']
    print('\t'.join(['lhs.name', 'rhs.name'] + columns))
    sys.stdout.flush()
    if not paths:
        return
    if any(MEASURES[m][1] is None for m in columns):
        with kmers(paths[0], 'r') as z:          # the reference gets as far as opening the first file (and its K check)
            if z.meta['K'] < K:
                raise MismatchedK(K, z.meta['K'])
        raise TypeError("cannot unpack non-iterable int object")
    sets = []
    for (whole, meta) in readKmerSetFiles(paths, counts=False):  # the next files are read and copied meanwhile
        if meta['K'] < K:
            raise MismatchedK(K, meta['K'])
        sets.append(whole.project(2 * (meta['K'] - K)))          # Measure.prep, set form (dist.py:36-49)
        whole.free()
    abc = _native.allpairs_abc(sets)
    pending = []
    for line in rows(paths, abc, columns):
        pending.append(line)
        if len(pending) == 4096:
            print('\n'.join(pending))
            pending = []
    if pending:
        print('\n'.join(pending))
    for s in sets:
        s.free()


if __name__ == '__main__':
    main(sys.argv[1:])
