"""
Usage:
    zot dump <input>
"""
# Drop-in for zotmer/commands/dump.py:13-29.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.basics import renderMany
from zotmer_b200.library.files import readKmers, readKmersAndCounts
from zotmer_b200.library.kmers import kmers


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    inp = opts['<input>']
    with kmers(inp, 'r') as z:
        K = z.meta['K']
        if 'kmers' not in z.meta:
            print('cannot dump "%s" as it contains no k-mers' % (inp,), file=sys.stderr)
            return
        out = sys.stdout
        B = 1 << 16
        if 'counts' in z.meta:
            (xs, cs) = readKmersAndCounts(z)
            for i in range(0, len(xs), B):
                out.write(''.join('%s\t%d\n' % (s, c) for (s, c) in zip(renderMany(K, xs[i:i + B]), cs[i:i + B].tolist())))
        else:
            xs = readKmers(z)
            for i in range(0, len(xs), B):
                out.write(''.join(s + '\n' for s in renderMany(K, xs[i:i + B])))


if __name__ == '__main__':
    main(sys.argv[1:])
