# `zot hist` (zotmer/commands/hist.py:14-24): print the count histogram stored in each container's metadata, one
# "file <tab> count <tab> number of k-mers" line per count in ascending order.  Nothing is scanned: the histogram was
# computed on the device (zb_set_stats) when kmerize / merge wrote the file.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library.kmers import kmers

__doc__ = usage.HIST


def histRows(path):
    """[(count, number of k-mers)] ascending by count; [] for a container without a histogram"""
    with kmers(path, 'r') as z:
        stored = z.meta.get('hist')
    if stored is None:
        return []
    return sorted((int(count), n) for (count, n) in stored.items())


def main(argv):
    for path in docopt.docopt(__doc__, argv)['<input>']:
        sys.stdout.write(''.join('%s\t%d\t%d\n' % (path, count, n) for (count, n) in histRows(path)))


if __name__ == '__main__':
    main(sys.argv[1:])
