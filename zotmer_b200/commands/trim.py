# `zot trim` (zotmer/commands/trim.py:54-95): keep the k-mers whose count is at least -c (and at most -C when -C > 0);
# the filter itself is zb_trim on the device.  Reference behaviours kept: the histogram in the output is the INPUT's
# (:91); `-c 0` asks the reference to infer a cut-off from the histogram, which dies with a TypeError because JSON turned
# the histogram keys into strings (:26-27, :82-84) -- same here, before anything is written.  One deliberate difference:
# under real docopt the reference raises KeyError('-C') on every run (its usage pattern lacks -C);
# docopt_mini knows the described option, so the command works.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library import setio

__doc__ = usage.TRIM


def bounds(opts):
    """(lowest count kept, highest count kept or 0 for no upper bound)"""
    lo = int(opts['-c']) if opts['-c'] is not None else 0
    hi = int(opts['-C']) if opts['-C'] is not None else 0
    return lo, (hi if hi > 0 else 0)


def main(argv):
    opts = docopt.docopt(__doc__, argv)
    lo, hi = bounds(opts)
    meta = setio.readMeta(opts['<input>'])
    meta['K'], meta['hist']                   # KeyError for a container without them, as the reference (:77-78)
    if lo == 0:
        raise TypeError("unsupported operand type(s) for -: 'str' and 'str'")
    full, _ = setio.readSetFile(opts['<input>'])
    kept = full.trim(lo, hi)
    full.free()
    setio.writeSetFile(opts['<output>'], kept, setio.carriedMeta(meta))
    kept.free()


if __name__ == '__main__':
    main(sys.argv[1:])
