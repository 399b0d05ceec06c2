"""
Usage:
    zot hist <input>...

Options:
    -u              update the input container to include the histogram
"""
# Drop-in for zotmer/commands/hist.py:14-24 (metadata only; the histogram itself is computed by
# zb_set_stats inside kmerize/merge).
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.kmers import kmers


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    for inp in opts['<input>']:
        with kmers(inp, 'r') as z:
            if 'hist' in z.meta:
                h = sorted((int(f), c) for (f, c) in z.meta['hist'].items())
                for (f, c) in h:
                    print('%s\t%d\t%d' % (inp, f, c))


if __name__ == '__main__':
    main(sys.argv[1:])
