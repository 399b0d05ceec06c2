"""
Usage:
    zot sample [-DS SEED] [-P PROBABILITY] <output> <input>

Options:
    -D              use deterministic sampling
    -P PROBABILITY  the proportion of samples to include in the output.
                    default: 0.01
    -S SEED         use the given seed for the sampling
"""
# Drop-in for zotmer/commands/sample.py:36-67.  The reference tests `opts['-D'] is None` (:53), which is never
# true for a docopt flag (flags are False/True), so it ALWAYS takes the deterministic path sampleD (:27-34):
# keep (x, c) iff float(murmer(x, S) & 0xFFFFFFFFFF) / float(0xFFFFFFFFFF) < p.  That filter is zb_sample(mode 0)
# on the device; the histogram of the kept counts (first-occurrence order, :31) comes from zb_set_stats.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.files import readKmerSet, writeKmerSet


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    p = 0.01
    if opts['-P'] is not None:
        p = float(opts['-P'])
    inp = opts['<input>']
    out = opts['<output>']
    with kmers(out, 'w') as z:
        with kmers(inp, 'r') as z0:
            K = z0.meta['K']
            z.meta = z0.meta.copy()
            del z.meta['kmers']
            del z.meta['counts']
            xs = readKmerSet(z0)
            S = 0
            if opts['-S']:
                S = int(opts['-S'])
            ys = xs.sample(p, S, 0)
            writeKmerSet(z, ys)
            h = {}
            for (c, f) in ys.stats()['hist']:
                h[c] = f
            xs.free()
            ys.free()
        z.meta['K'] = K
        z.meta['kmers'] = 'kmers'
        z.meta['counts'] = 'counts'
        z.meta['hist'] = h


if __name__ == '__main__':
    main(sys.argv[1:])
