# `zot dump` (zotmer/commands/dump.py:13-29): the k-mers of a container as text, "ACGT...<tab>count" per line when the
# container has counts, the bare k-mer otherwise.  The streams are decoded on the device (library/files.py) and rendered
# 65,536 k-mers at a time with numpy (library/basics.py renderMany) instead of one Python call per k-mer.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library.basics import renderMany
from zotmer_b200.library.files import readKmers, readKmersAndCounts
from zotmer_b200.library.kmers import kmers

__doc__ = usage.DUMP
CHUNK = 1 << 16


def textChunks(K, xs, cs=None):
    """the dump, CHUNK k-mers per string"""
    for lo in range(0, len(xs), CHUNK):
        words = renderMany(K, xs[lo:lo + CHUNK])
        if cs is None:
            yield ''.join(w + '\n' for w in words)
        else:
            yield ''.join('%s\t%d\n' % wc for wc in zip(words, cs[lo:lo + CHUNK].tolist()))


def main(argv):
    path = docopt.docopt(__doc__, argv)['<input>']
    with kmers(path, 'r') as z:
        K = z.meta['K']                      # KeyError for a container without K, as the reference
        if 'kmers' not in z.meta:
            print('cannot dump "%s" as it contains no k-mers' % (path,), file=sys.stderr)
            return
        if 'counts' in z.meta:
            xs, cs = readKmersAndCounts(z)
        else:
            xs, cs = readKmers(z), None
    for text in textChunks(K, xs, cs):
        sys.stdout.write(text)


if __name__ == '__main__':
    main(sys.argv[1:])
