# Whole-file helpers of the set commands (trim, sample, project): a k-mer set file in, a device-resident set + its
# metadata out, and back.  The JSON metadata is written in dictionary order, so the helpers that build it state the
# key order the reference's files have (tests compare file bytes).
from zotmer_b200.library.files import readKmerSet, writeKmerSet, writeWords
from zotmer_b200.library.kmers import kmers


def readMeta(path):
    with kmers(path, 'r') as z:
        return dict(z.meta)


def readSetFile(path, counts=True, device=0):
    """-> (KmerSet on the device, metadata dict)"""
    with kmers(path, 'r') as z:
        return readKmerSet(z, counts=counts, device=device), dict(z.meta)


def carriedMeta(meta):
    """the metadata a filter command passes on: everything of the input in its order, the stream names re-appended
    (the reference deletes 'kmers' / 'counts' from a copy and sets them again after writing the streams)"""
    out = dict((k, v) for (k, v) in meta.items() if k not in ('kmers', 'counts'))
    out['kmers'] = 'kmers'
    out['counts'] = 'counts'
    return out


def writeSetFile(path, kset, meta, counts=True):
    """streams of `kset` ('kmers' [+ 'counts']) followed by the metadata"""
    with kmers(path, 'w') as w:
        if counts:
            writeKmerSet(w, kset)
        else:
            kw, _ = kset.encode()
            with w.add_stream('kmers') as f:
                writeWords(f, kw)
        w.meta = meta
