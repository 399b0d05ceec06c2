"""
Several GPUs behind ONE `zot` process (ZB_GPUS=N, or all visible devices with ZB_GPUS=all).

The reference is one process, one thread (zotmer/cli.py:45-59); so are the commands here, except that `zot kmerize`
may spread its work over N devices: one host thread per device (the C ABI releases the GIL, and the library keeps one
context -- stream, allocator -- per (device, host thread)), NVLink peer mappings between the devices
(zb_peer_enable), no NCCL and no second process.  The steps, in the order kmerizeFilesMulti runs them:

  1. the input files are cut into record-aligned pieces and dealt out to the devices in ROUNDS of N pieces; a device
     stages its piece through the pinned ring, parses it and extracts the canonical k-mers (csrc/parse.cu, extract.cu).
     A FASTA record longer than a device's share (a chromosome) is cut inside (reads.splitPieces) and spread over the
     devices.  A block-compressed file (BGZF) is dealt out as runs of whole gzip members instead: a device inflates its run itself
     (csrc/inflate.cu) and the incomplete record behind a run's last record boundary travels to the device that holds
     the next run (_stageBgzfGroup);
  2. per round, the hash-range exchange of bench.py / multigpu.py, without collectives: the per-owner counts of all
     devices meet in a Python list (a threading.Barrier is the "all-gather"), every owner allocates its receive buffer
     at the exact size, route_p2p_kernel stores each key into its owner's buffer over NVLink, a second barrier says
     that all stores have landed; the owner adopts the buffer and counts it (sort + count, csrc/segsort.cu);
  3. every device finishes its hash-owned share (mirror + merge: the both-strand set of its k-mers);
  4. one sorted file needs key RANGES, not hash shares: splitters are quantiles of device 0's share (a uniform sample
     of all distinct k-mers, since ownership is a hash), every device cuts its sorted share at the splitters
     (zb_set_lower_bound), device r copies range r of every share out of its peers' memory (zb_set_from_device on a
     peer pointer) and merges the N sorted runs (zb_merge) -- range r of the final set, no gather to one GPU;
  5. stats per range, combined on the host (histogram keys in order of first occurrence along the ranges);
  6. the range-partitioned codec64 encode (zb_set_encode_plan / _emit): every range is encoded where it lives, the
     six-state maps of the ranges are chained on the host, and all devices write their words into the one output file
     at once (zb_words_write_fd).  The file is byte-identical to the single-GPU one (tests/test_gpu_multi.py).
"""
import os
import threading

import numpy as np

from zotmer_b200 import _native
from zotmer_b200.library.reads import isFasta, pieces, MAX_PIECE


def deviceList():
    """devices the commands may use: ZB_GPUS=N -> 0..N-1, ZB_GPUS=all -> every visible device, unset -> device 0"""
    want = os.environ.get("ZB_GPUS", "1").strip().lower()
    have = _native.device_count()
    n = have if want in ("all", "0") else int(want)
    if n < 1 or n > have:
        raise SystemExit("ZB_GPUS=%s: this machine has %d CUDA device(s)" % (want, have))
    return list(range(n))


class _Group(object):
    """N device threads that meet at barriers and share a blackboard"""

    def __init__(self, devs):
        self.devs = devs
        self.n = len(devs)
        self.barrier = threading.Barrier(self.n)
        self.board = {}
        self.errors = []
        for a in devs:
            for b in devs:
                if a != b:
                    _native.peer_enable(a, b)

    def meet(self):
        self.barrier.wait()

    def run(self, fn):
        """fn(rank) on one thread per device -> list of results; the first exception of any thread is raised"""
        out = [None] * self.n

        def body(r):
            try:
                out[r] = fn(r)
            except BaseException as e:      # noqa: B902 -- must release the others from their barriers
                self.errors.append(e)
                self.barrier.abort()

        ths = [threading.Thread(target=body, args=(r,)) for r in range(self.n)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if self.errors:
            first = [e for e in self.errors if not isinstance(e, threading.BrokenBarrierError)]
            raise (first or self.errors)[0]
        return out


class _BgzfGroup(object):
    """a run of whole members of a block-compressed file: inflated on the device that gets it (library zb_stage_bgzf);
    its text starts with the rest of the record the previous group ended in (`first`: nothing to wait for) and the
    record it ends in is completed by the next group (`last`: nothing to hand on)"""

    def __init__(self, comp, job, first, last):
        self.comp, self.job, self.first, self.last = comp, job, first, last
        self.tail = None                    # bytes behind this group's last record boundary, for the next group
        self.done = threading.Event()

    def __len__(self):
        return len(self.comp)


def _bgzfJobs(fn, fa, n, jobs):
    """groups of members of the BGZF file `fn` appended to `jobs`; False if it is not BGZF"""
    import os
    from zotmer_b200.library.reads import _bgzf, BGZF_GROUP
    if not (fn.endswith('.gz') and os.path.isfile(fn)):
        return False
    comp = _bgzf(fn)
    if comp is None:
        return False
    a = np.frombuffer(comp, dtype=np.uint8)
    (_, text) = _native.bgzf_probe(a)
    if text == 0:
        return True
    share = max(1 << 20, min(BGZF_GROUP, -(-text // n)))
    starts = _native.bgzf_groups(a, share) + [len(a)]
    for i in range(len(starts) - 1):
        jobs.append((_BgzfGroup(a[starts[i]:starts[i + 1]], len(jobs), i == 0, i == len(starts) - 2), fa))
    return True


def _stageBgzfGroup(grp, prev, fa, dev, errors):
    """the staged piece of a group (None: it holds no record boundary): inflate, put the previous group's leftover in
    front, cut at the last record boundary and hand the bytes behind it on.  `prev`: the group before it in the file."""
    nat = _native
    st, _ = nat.stage_bgzf(grp.comp, dev)
    if not grp.first:
        while not prev.done.wait(0.05):
            if errors:
                st.free()
                raise RuntimeError("another device failed")
        if prev.tail:
            st = nat.stage_concat(prev.tail, st, dev)
        prev.tail = None
    try:
        if grp.last:
            return st
        cut = st.cut(fa)
        grp.tail = st.fetch_range(cut, len(st) - cut)
        if cut == 0:
            st.free()
            return None
        st.set_len(cut)
        return st
    finally:
        grp.done.set()


def _pieceRounds(inputs, n, verbose=False, k=None):
    """record-aligned pieces of all inputs, about 1/n of a file each (at most MAX_PIECE), grouped n at a time.
    k: the k-mer length when a FASTA record longer than a share may be cut inside (reads.splitPieces: a chromosome is then
    spread over the devices like any other input); such a piece is a (prefix, piece, extra records) triple."""
    import sys
    from zotmer_b200.library.file import mapBytes
    jobs = []
    for fn in inputs:
        held = None
        if isinstance(fn, tuple):
            fn, held = fn
        fa = isFasta(fn)
        if held is None and _bgzfJobs(fn, fa, n, jobs):
            if verbose:
                print('reading %s (block-compressed, %s): inflated on the devices' % (fn, 'FASTA' if fa else 'FASTQ'), file=sys.stderr)
            continue
        data = held if held is not None else mapBytes(fn)
        if verbose:
            print('reading %s (%d bytes, %s)' % (fn, len(data), 'FASTA' if fa else 'FASTQ'), file=sys.stderr)
        if len(data) == 0:
            continue
        share = max(1 << 20, min(MAX_PIECE, -(-len(data) // n)))
        if k is not None and fa:
            from zotmer_b200.library.reads import splitPieces
            for t in splitPieces(data, fa, k, share):
                if len(t[1]):
                    jobs.append((t if (t[0] or t[2]) else t[1], fa))
            continue
        try:
            cut = list(pieces(data, fa, share))
        except ValueError:      # a record longer than a share (a chromosome): it goes to one device whole
            cut = list(pieces(data, fa, MAX_PIECE))
        for p in cut:
            if len(p):
                jobs.append((p, fa))
    return [jobs[i:i + n] for i in range(0, len(jobs), n)]


def mergeHists(hists):
    """count histograms of consecutive key ranges -> the histogram of the whole set, keys in order of first occurrence"""
    out = {}
    for h in hists:
        for (c, f) in h:
            out[c] = out.get(c, 0) + f
    return list(out.items())


def kmerizeFilesMulti(K, inputs, devs, verbose=False, baits_fn=None):
    """-> (list of per-device KmerSets = consecutive key ranges of the both-strand counted set, number of records).
    baits_fn: FASTA file of bait sequences (`-C`); every device kmerizes it for itself."""
    g = _Group(devs)
    n = g.n
    # capture mode keeps or drops whole records: a record is then never cut inside
    rounds = _pieceRounds(inputs, n, verbose, k=K if baits_fn is None else None)
    flat = [job for rd in rounds for job in rd]     # a BGZF group finds the group before it here
    fake = sum(job[0][2] for job in flat if isinstance(job[0], tuple))   # headers that splitPieces put in front of cut records
    nat = _native

    def work(r):
        dev = devs[r]
        baits = None
        if baits_fn is not None:
            from zotmer_b200.commands.kmerize import baitSet
            baits = baitSet(K, baits_fn, dev)
        km = nat.Kmerizer(K, dev)
        km.set_owners(n)          # extraction tallies the keys per owner: no counting pass before the exchange
        if baits is not None:
            km.set_baits(baits)
        bufs = []

        def stage(ri):
            """the piece of this device in round ri on its way to the device"""
            src = rounds[ri][r][0]
            if isinstance(src, _BgzfGroup):
                prev = flat[src.job - 1][0] if not src.first else None
                return _stageBgzfGroup(src, prev, rounds[ri][r][1], dev, g.errors)
            if isinstance(src, tuple):             # the piece behind a cut inside a record: its prefix goes in front
                st = nat.stage_input(src[1], dev)
                return nat.stage_concat(src[0], st, dev) if src[0] else st
            return nat.stage_input(src, dev)

        try:
            staged = None
            if rounds and r < len(rounds[0]):
                staged = stage(0)
            for ri, rd in enumerate(rounds):
                if staged is not None:
                    km.feed_staged(staged, rd[r][1])
                    staged = None
                if ri + 1 < len(rounds) and r < len(rounds[ri + 1]):
                    staged = stage(ri + 1)     # copies (or inflates) while this round is exchanged
                # ---- exchange: counts meet on the board, owners allocate, everybody routes, owners adopt
                g.board[("cnt", r)] = km.bucket_counts(n)
                g.meet()
                M = np.array([g.board[("cnt", s)] for s in range(n)], dtype=np.int64)     # [src][owner]
                nrecv = int(M[:, r].sum())
                ptr, _ = nat.ipc_alloc(max(nrecv, 1) * 8 + 256, dev)
                bufs.append(ptr)
                g.board[("buf", r)] = ptr
                g.meet()
                offs = [int(M[:r, o].sum()) for o in range(n)]
                km.route_p2p([g.board[("buf", o)] + 8 * offs[o] for o in range(n)])
                g.meet()                                  # every device's stores have completed
                km.adopt_canonical_dev(ptr, nrecv)
                km.flush()                                # counted now: the buffer can go
                nat.ipc_free(bufs.pop(), dev)
            share, nr = km.finish()
        finally:
            km.close()
            for p in bufs:
                nat.ipc_free(p, dev)
            if baits is not None:
                baits.free()
        # ---- key ranges instead of hash shares
        if r == 0:
            m = len(share)
            pos = [(m * q) // n for q in range(1, n)]
            g.board["split"] = np.array([int(share.slice(p, p + 1).fetch(counts=False)[0]) if m else 0 for p in pos], np.uint64)
        g.meet()
        split = g.board["split"]
        idx = [0] + [int(x) for x in share.lower_bound(split)] + [len(share)]
        for q in range(1, len(idx)):                      # equal splitters / an empty share: keep the cuts ordered
            idx[q] = max(idx[q], idx[q - 1])
        g.board[("share", r)] = (share.dev_ptrs(), idx)
        g.meet()
        parts = []
        for s in range(n):
            (kp, cp), ix = g.board[("share", s)]
            b, e = ix[r], ix[r + 1]
            parts.append(nat.KmerSet.from_device(kp + 8 * b, cp + 4 * b, e - b, dev))     # peer memory -> my memory
        g.meet()                                          # everybody has copied: the shares can go
        share.free()
        mine = nat.merge(parts) if n > 1 else parts[0]
        for p in parts:
            if p is not mine:
                p.free()
        return mine, nr

    res = g.run(work)
    return [x[0] for x in res], sum(x[1] for x in res) - fake


def statsMulti(ranges):
    """zb_set_stats of consecutive ranges, combined: acgt tallies add up; the histogram keeps first-occurrence order"""
    sts = [s.stats() for s in ranges]
    out = {"acgt_weighted": [sum(st["acgt_weighted"][q] for st in sts) for q in range(4)],
           "acgt_plain": [sum(st["acgt_plain"][q] for st in sts) for q in range(4)],
           "total": sum(st["total"] for st in sts), "hist": mergeHists([st["hist"] for st in sts])}
    return out


def writeRangesMulti(z, ranges, nm=None):
    """the 'kmers' and 'counts' streams of the set whose consecutive key ranges live on several devices, written into the
    open container `z` exactly as files.writeKmerSet writes a single-device set (files.py:209-217)."""
    from zotmer_b200.library.files import _names
    xNm, cNm = _names(nm)
    n = len(ranges)
    sizes = [len(s) for s in ranges]
    # halo of every range: the last k-mer in front of it, the first (up to) five entries behind it
    heads = []
    for s in ranges:
        m = min(5, len(s))
        if m:
            h = s.slice(0, m)
            heads.append(h.fetch())
            h.free()
        else:
            heads.append((np.zeros(0, np.uint64), np.zeros(0, np.uint32)))
    lasts = []
    for s in ranges:
        if len(s):
            t = s.slice(len(s) - 1, len(s))
            lasts.append(int(t.fetch(counts=False)[0]))
            t.free()
        else:
            lasts.append(None)
    plans, kmaps, cmaps = [], [], []
    prev = 0
    for r in range(n):
        nk = np.concatenate([heads[q][0] for q in range(r + 1, n)] + [np.zeros(0, np.uint64)])[:5]
        nc = np.concatenate([heads[q][1] for q in range(r + 1, n)] + [np.zeros(0, np.uint32)])[:5]
        p, km, cm = ranges[r].encode_plan(prev, nk, nc)
        plans.append(p); kmaps.append(km); cmaps.append(cm)
        if lasts[r] is not None:
            prev = lasts[r]
    kentry, koff, ktot = _native.chain_ranges(kmaps)
    centry, coff, ctot = _native.chain_ranges(cmaps)
    z._writable()
    z.fo.flush()
    at = z._end()
    fd = z.fo.fileno()
    errors = []

    def emit(r):
        try:
            w = plans[r].emit(kentry[r], centry[r])
            try:
                w.write_fd(fd, at + 8 * koff[r], at + 8 * ktot + 8 * coff[r])
            finally:
                w.free()
        except BaseException as e:      # noqa: B902
            errors.append(e)

    ths = [threading.Thread(target=emit, args=(r,)) for r in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errors:
        raise errors[0]
    z._record(xNm, at, 8 * ktot)
    z._record(cNm, at + 8 * ktot, 8 * ctot)
    return sum(sizes)
