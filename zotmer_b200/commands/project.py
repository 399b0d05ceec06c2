"""
Usage:
    zot project <ref> <output> <input>

Project one or more inputs on to a reference set. For each k-mer in <ref>,
a whitespace separated 0 or a 1 is printed indicating whether that k-mer
was present in the input, with a separate line for each input k-mer set.
"""
# Drop-in for zotmer/commands/project.py:42-70: the output holds the entries of <input> whose k-mer occurs in
# <ref> (project1 / project2, :18-40 = a sorted intersection that keeps the input's counts) -- zb_restrict on
# the device.  As in the reference the histogram is copied from the input, not recomputed (:68).
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.files import readKmerSet, writeKmerSet, writeWords


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    with kmers(opts['<ref>'], 'r') as z:
        K = z.meta['K']
        xs = readKmerSet(z, counts=False)

    with kmers(opts['<input>'], 'r') as z0:
        K0 = z0.meta['K']
        if K0 != K:
            print("mismatched K (%d)" % (K0, ), file=sys.stderr)
            sys.exit(1)

        with kmers(opts['<output>'], 'w') as z:
            z.meta['K'] = K
            if 'counts' in z0.meta:
                ys = readKmerSet(z0)
                zs = ys.restrict(xs)
                writeKmerSet(z, zs)
                z.meta['kmers'] = 'kmers'
                z.meta['counts'] = 'counts'
            else:
                ys = readKmerSet(z0, counts=False)
                zs = ys.restrict(xs)
                (kw, _) = zs.encode()
                with z.add_stream('kmers') as f:
                    writeWords(f, kw)
                z.meta['kmers'] = 'kmers'
            z.meta['hist'] = z0.meta['hist']
            ys.free()
            zs.free()
    xs.free()


if __name__ == '__main__':
    main(sys.argv[1:])
