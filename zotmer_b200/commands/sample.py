# `zot sample` (zotmer/commands/sample.py:36-67): keep a pseudo-random fraction -P of the k-mers.  The reference tests
# `opts['-D'] is None` (:53), which is never true for a docopt flag, so it ALWAYS samples deterministically (sampleD,
# :27-34): (x, c) stays iff float(murmer(x, S) & 0xFFFFFFFFFF) / float(0xFFFFFFFFFF) < p -- zb_sample on the device,
# evaluated in IEEE double like the reference.  The histogram of the output is recomputed from the kept counts, in
# first-occurrence order (:31).
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library import setio

__doc__ = usage.SAMPLE


def main(argv):
    opts = docopt.docopt(__doc__, argv)
    fraction = float(opts['-P']) if opts['-P'] is not None else 0.01
    seed = int(opts['-S']) if opts['-S'] else 0
    full, meta = setio.readSetFile(opts['<input>'])
    meta['K']                                 # KeyError for a container without K (:49)
    kept = full.sample(fraction, seed, 0)
    full.free()
    out = setio.carriedMeta(meta)
    out['hist'] = dict(kept.stats()['hist'])
    setio.writeSetFile(opts['<output>'], kept, out)
    kept.free()


if __name__ == '__main__':
    main(sys.argv[1:])
