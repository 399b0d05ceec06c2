"""
Usage:
    zot trim [-c CUTOFF] <output> <input>

Options:
    -c CUTOFF   discard k-mers with frequency less than CUTOFF. A
                cutoff of 0 (the default) indicates that cutoff
                inference should be used. [default: 0]
    -C CUTOFF   discard k-mers with frequency greater than CUTOFF.
                A cutoff of 0 (the default) indicates that the
                cutoff value should be effectively infinite.
                [default: 0]
"""
# Drop-in for zotmer/commands/trim.py:64-95; the filter generator (:54-62) is zb_trim on the device.
# Under real docopt the reference raises KeyError('-C') on every run (-C is missing from its usage
# pattern); here -C takes its documented default so the command works.  `-c 0`
# (cut-off inference) is a TypeError in the reference (:26-27,82-84: JSON made the hist keys
# strings) and stays one.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.files import readKmerSet, writeKmerSet


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    inp = opts['<input>']
    out = opts['<output>']

    c = 0
    if opts['-c'] is not None:
        c = int(opts['-c'])

    C = None
    if opts['-C'] is not None:
        C0 = int(opts['-C'])
        if C0 > 0:
            C = C0

    with kmers(inp, 'r') as z:
        K = z.meta['K']
        h = z.meta['hist']
        if c == 0:
            raise TypeError("unsupported operand type(s) for -: 'str' and 'str'")
        xs = readKmerSet(z)
        with kmers(out, 'w') as w:
            w.meta = z.meta.copy()
            del w.meta['kmers']
            del w.meta['counts']
            writeKmerSet(w, xs.trim(c, 0 if C is None else C))
            w.meta['K'] = K
            w.meta['kmers'] = 'kmers'
            w.meta['counts'] = 'counts'
            w.meta['hist'] = h


if __name__ == '__main__':
    main(sys.argv[1:])
