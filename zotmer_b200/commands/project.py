# `zot project` (zotmer/commands/project.py:18-70): the entries of <input> whose k-mer also occurs in <ref>, with the
# input's counts (project1 / project2 walk the two sorted streams; here zb_restrict on the device).  Kept from the
# reference: K of the two files must agree ("mismatched K (n)" on stderr, exit status 1, nothing written); an input
# without counts gives an output without counts; the histogram is copied from the input, not recomputed (:68).
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library import setio

__doc__ = usage.PROJECT


def main(argv):
    opts = docopt.docopt(__doc__, argv)
    ref_meta = setio.readMeta(opts['<ref>'])
    in_meta = setio.readMeta(opts['<input>'])
    K = ref_meta['K']
    if in_meta['K'] != K:
        print("mismatched K (%d)" % (in_meta['K'],), file=sys.stderr)
        sys.exit(1)
    counted = 'counts' in in_meta
    wanted, _ = setio.readSetFile(opts['<ref>'], counts=False)
    given, _ = setio.readSetFile(opts['<input>'], counts=counted)
    common = given.restrict(wanted)
    given.free()
    wanted.free()
    out = {'K': K, 'kmers': 'kmers'}
    if counted:
        out['counts'] = 'counts'
    out['hist'] = in_meta['hist']
    setio.writeSetFile(opts['<output>'], common, out, counts=counted)
    common.free()


if __name__ == '__main__':
    main(sys.argv[1:])
