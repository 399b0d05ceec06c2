# `zot info` (zotmer/commands/info.py:11-19): the metadata of each container, one "key value" line per entry in key order.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library.setio import readMeta

__doc__ = usage.INFO


def main(argv):
    for path in docopt.docopt(__doc__, argv)['<input>']:
        meta = readMeta(path)
        for key in sorted(meta):
            print(key, meta[key])


if __name__ == '__main__':
    main(sys.argv[1:])
