# Drop-in for zotmer/commands/kmerize.py:450-562.  The per-record Python loop (reads() ->
# kmersList -> acgt tally -> KmerAccumulator2 -> radix_sort -> merge, :490-545) is replaced by
# zb_kmerize_* : raw file bytes go to the GPU, which parses, extracts both strands, sorts and counts.
# The spill machinery (-m, :527-553) is not reproduced: its output is byte-identical to the
# in-memory path (golden `spill_equals_inmemory`); -m is accepted and ignored.
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import _native
from zotmer_b200.library.file import mapBytes
from zotmer_b200.library.files import writeKmerSet
from zotmer_b200.library.reads import isFasta, pieces, stagedPieces
import zotmer_b200.library.kmers as zotk
from zotmer_b200 import usage

__doc__ = usage.KMERIZE


def kmerizeFiles(K, inputs, device=0, verbose=False, baits=None):
    """-> (KmerSet with both strands, number of records).  kmerize.py:463-539.
    baits: KmerSet of the bait k-mers (-C): only records holding one of them contribute (kmerize.py:507-517)."""
    km = _native.Kmerizer(K, device)
    try:
        if baits is not None:
            km.set_baits(baits)
        fake = 0
        # capture mode keeps or drops whole records: a record is then never cut inside
        for (staged, fa) in stagedPieces(inputs, device, verbose, k=K if baits is None else None):
            fake += getattr(staged, 'fake_records', 0)
            km.feed_staged(staged, fa)
        (s, nr) = km.finish()
        return s, nr - fake
    finally:
        km.close()


def baitSet(K, fn, device=0):
    """kmerize.py:478-483: B = set of the k-mers (both strands) of every record of the FASTA file `fn` (readFasta
    whatever its suffix)."""
    km = _native.Kmerizer(K, device)
    try:
        data = mapBytes(fn)
        for piece in pieces(data, True):
            km.feed(piece, True)
        (b, _) = km.finish()
        return b
    finally:
        km.close()


def main(argv):
    import os
    import time
    trace = [] if os.environ.get('ZB_CLI_TRACE') else None

    def mark(what):
        if trace is not None:
            trace.append((what, time.perf_counter()))

    mark('start')
    opts = docopt.docopt(__doc__, argv)

    verbose = opts['-v']
    K = int(opts['<k>'])
    out = opts['<output>']

    inputs = opts['<input>']
    if opts['-C'] is not None and opts['-D'] is None:
        # capture reads the inputs twice (below); what cannot be read twice (stdin) or is costly to produce twice
        # (decompressed text) is held in memory after the first read -- the reference reads its input once
        from zotmer_b200.library.file import readBytes
        inputs = [(fn, readBytes(fn)) if (fn == '-' or fn.endswith(('.gz', '.bz2'))) else fn for fn in inputs]
    from zotmer_b200.library import devices
    devs = devices.deviceList()
    multi = len(devs) > 1     # ZB_GPUS=N: the set then lives on N devices as consecutive key ranges (library/devices.py)

    def kmerizeAll(baits_fn=None):
        if multi:
            return devices.kmerizeFilesMulti(K, inputs, devs, verbose=verbose, baits_fn=baits_fn)
        baits = baitSet(K, baits_fn, devs[0]) if baits_fn is not None else None
        try:
            (one, n_) = kmerizeFiles(K, inputs, device=devs[0], verbose=verbose, baits=baits)
        finally:
            if baits is not None:
                baits.free()
        return [one], n_

    def statsAll(rs):
        return devices.statsMulti(rs) if multi else rs[0].stats()

    (ranges, nr) = kmerizeAll()
    mark('read + kmerize + count')
    st = statsAll(ranges)
    mark('stats')      # acgt counts EVERY k-mer, before any sub-sampling / capture (kmerize.py:492-493)

    if opts['-C'] is not None and opts['-D'] is None:
        # kmerize.py:507-517 (the -D branch comes first in the reference's if / elif chain): whole records are kept
        # when one of their k-mers is a bait.  acgt still covers every record (above), hence the second pass.
        for r_ in ranges:
            r_.free()
        (ranges, nr) = kmerizeAll(opts['-C'])
        st['hist'] = statsAll(ranges)['hist']

    with zotk.kmers(out, 'w') as z:
        if opts['-D'] is not None:
            # kmerize.py:494-506: a k-mer is kept iff sub(S, d, x) (basics.py:251-259) -- a function of the k-mer
            # alone (each strand is tested on its own), so filtering the counted set is the same as filtering
            # the k-mers before counting them
            d = float(opts['-D'])
            S = 0
            if opts['-S'] is not None:
                S = int(opts['-S'])
            kept = [r_.sample(d, S, 1) for r_ in ranges]
            for r_ in ranges:
                r_.free()
            ranges = kept
            st['hist'] = statsAll(ranges)['hist']
        h = {}
        for (c, f) in st['hist']:       # first-occurrence order == the reference's dict order (:544-545)
            h[c] = f
        mark('open output')
        if multi:
            devices.writeRangesMulti(z, ranges)
        else:
            writeKmerSet(z, ranges[0])
        mark('encode + write streams')
        acgt = st['acgt_weighted']
        n = float(sum(acgt))
        acgt = [c / n for c in acgt]    # ZeroDivisionError on input without k-mers, as :554-555
        z.meta['K'] = K
        z.meta['kmers'] = 'kmers'
        z.meta['counts'] = 'counts'
        z.meta['hist'] = h
        z.meta['acgt'] = acgt
        z.meta['reads'] = nr
    mark('meta + table + close')
    for r_ in ranges:
        r_.free()
    if trace is not None:
        print('zot kmerize phases (ms): ' + ', '.join('%s %.1f' % (trace[i][0], (trace[i][1] - trace[i - 1][1]) * 1e3)
                                                      for i in range(1, len(trace))), file=sys.stderr)


if __name__ == '__main__':
    main(sys.argv[1:])
