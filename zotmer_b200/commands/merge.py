# Drop-in for zotmer/commands/merge.py:165-253.  The pairwise generator merge (:26-86) and the
# heap-of-radix-blocks N-way merge (:127-163) are replaced by zb_merge (merge-path + reduce-by-key
# on the device); hist/acgt come from zb_set_stats.  The reference's behaviour for 1 and 2 inputs is
# kept as it is: 1 input -> ZeroDivisionError after the streams were written;
# 2 inputs -> meta holds only hist/acgt and acgt is NOT count-weighted (:88-92,173-199).
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import _native
from zotmer_b200.library.kmers import kmers
from zotmer_b200.library.files import readKmerSet, readKmerSetFiles, writeKmerSet
from zotmer_b200 import usage

__doc__ = usage.MERGE


def _histDict(st):
    h = {}
    for (c, f) in st['hist']:
        h[c] = f
    return h


def main(argv):
    opts = docopt.docopt(__doc__, argv)

    out = opts['<output>']
    inputs = opts['<input>']

    if len(inputs) <= 2:
        with kmers(out, 'w') as z:
            h = {}
            acgt = [0, 0, 0, 0]
            if len(inputs) == 1:
                with kmers(inputs[0], 'r') as z0:
                    xs = readKmerSet(z0)
                    writeKmerSet(z, xs)      # the hist() generator is never consumed (:181-183)
            else:
                with kmers(inputs[0], 'r') as z0, kmers(inputs[1], 'r') as z1:
                    K = z0.meta['K']
                    K1 = z1.meta['K']
                    if K1 != K:
                        print("mismatched K", file=sys.stderr)
                        sys.exit(1)
                    zs = _native.merge([readKmerSet(z0), readKmerSet(z1)])
                    st = zs.stats()
                    h = _histDict(st)
                    acgt = st['acgt_plain']  # `acgt[z[0]&3] += 1` (:88-92)
                    writeKmerSet(z, zs)
            n = float(sum(acgt))
            acgt = [c / n for c in acgt]
            z.meta['hist'] = h
            z.meta['acgt'] = acgt
        return

    K = None
    sets = []
    # the files come in one after the other with the next ones already being read and copied (files.readKmerSetFiles);
    # K is checked in the reference's order: pair by pair, the second file of the FIRST pair never (:203-207)
    for i, (xs, meta) in enumerate(readKmerSetFiles(inputs)):
        sets.append(xs)
        if i == 0:
            K = meta['K']
        elif i >= 2 and meta['K'] != K:
            print("mismatched K", file=sys.stderr)
            sys.exit(1)

    assert K is not None

    with kmers(out, 'w') as z:
        merged = _native.merge(sets)
        st = merged.stats()
        writeKmerSet(z, merged)
        acgt = st['acgt_weighted']           # `acgt[x&3] += c` (:159)
        n = float(sum(acgt))
        acgt = [c / n for c in acgt]
        z.meta['K'] = K
        z.meta['kmers'] = 'kmers'
        z.meta['counts'] = 'counts'
        z.meta['hist'] = _histDict(st)
        z.meta['acgt'] = acgt


if __name__ == '__main__':
    main(sys.argv[1:])
