"""
Usage:
    zot info <input>...
"""
# Drop-in for zotmer/commands/info.py:11-19.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.kmers import kmers


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    for inp in opts['<input>']:
        with kmers(inp, 'r') as z:
            for (k, v) in sorted(z.meta.items()):
                print(k, v)


if __name__ == '__main__':
    main(sys.argv[1:])
