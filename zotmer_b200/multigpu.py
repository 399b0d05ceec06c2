"""
Multi-GPU kmerize+count: hash-range routing of canonical k-mers (one process per GPU, torch.distributed).

The reference has no parallelism at all; this is the B200 layer BASELINE.json's north star asks for:
reads are sharded across ranks, every rank extracts the canonical k-mers of its shard, each k-mer is sent
to its OWNER rank -- owner(x) = floor(mix64(x) * nranks / 2^64), the high bits of an invertible 64-bit mix,
so ownership is uniform even on low-complexity sequence -- with ONE all-to-all over NVLink (sizes first),
and every rank sorts and counts its disjoint share locally.  x and rc(x) land on different owners, which
is fine: the canonical key is what is routed, mirroring happens after counting on the owner.

All-pairs distances (`zot dist`, `zot jaccard -a`) shard without any exchange: the upper triangle of the
set x set matrix is cut into work units (pairs of blocks of 32 sets x 8 key-range shards, csrc/allpairs.cu), every rank computes a
contiguous, pair-count-balanced range of tiles over its own copy of the sets, and the partial matrices (zero
outside a rank's tiles) are added up -- the "final gather" of the north star.

`owner_of` is the host (numpy) statement of the device function in csrc/extract.cu; tests check one
against the other.  `exchange_tensors` is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    """murmur3 fmix64 on a uint64 array (same constants as csrc/extract.cu:mix64)"""
    x = np.ascontiguousarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def owner_of(keys, nranks):
    """floor(mix64(key) * nranks / 2^64) -> int array in [0, nranks)"""
    h = mix64(keys)
    hi = (h >> np.uint64(32)).astype(np.uint64)
    lo = (h & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    n = np.uint64(nranks)
    # (hi*2^32 + lo) * n >> 64  without 128-bit ints
    t = (lo * n) >> np.uint64(32)
    return (((hi * n) + t) >> np.uint64(32)).astype(np.int64)


def bucket_host(keys, nranks):
    """host stand-in for zb_kmerize_take_bucketed_dev: keys grouped by owner + bucket sizes"""
    ow = owner_of(keys, nranks)
    order = np.argsort(ow, kind="stable")
    counts = np.bincount(ow, minlength=nranks).astype(np.int64)
    return np.ascontiguousarray(keys, dtype=np.uint64)[order], [int(c) for c in counts]


def exchange_tensors(dist, send, send_counts, make_recv):
    """all-to-all of variable-sized int64 tensors: sizes first, then the payload.
    send: 1-D int64 tensor grouped by destination rank; send_counts: python ints per destination;
    make_recv(n) -> tensor of n int64 on the right device.  Returns (recv tensor, recv_counts)."""
    import torch
    cin = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
    cout = torch.empty_like(cin)
    dist.all_to_all_single(cout, cin)
    recv_counts = [int(x) for x in cout.tolist()]
    nrecv = sum(recv_counts)
    recv = make_recv(nrecv)
    dist.all_to_all_single(recv[:nrecv], send[:sum(send_counts)], recv_counts, list(send_counts))
    return recv, recv_counts


def exchange_pending(nat, km, ctx):
    """GPU path: route every pending canonical k-mer of `km` to its owner and hand the received keys
    back to the kmerizer.  ctx: dict(world, rank, dev, send, recv, a2a_ms, a2a_bytes)."""
    import torch
    import torch.distributed as dist
    world, dev = ctx["world"], ctx["dev"]
    n = km.pending()
    if ctx["send"].numel() < max(n, 1):
        ctx["send"] = torch.empty(int(n * 1.1) + 16, dtype=torch.int64, device="cuda:%d" % dev)
    send = ctx["send"]
    counts = km.take_bucketed_dev(world, send.data_ptr())

    def make_recv(m):
        if ctx["recv"].numel() < max(m, 1):
            ctx["recv"] = torch.empty(int(m * 1.1) + 16, dtype=torch.int64, device="cuda:%d" % dev)
        return ctx["recv"]

    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    recv, recv_counts = exchange_tensors(dist, send, counts, make_recv)
    t1.record()
    torch.cuda.synchronize(dev)
    ctx["a2a_ms"].append(t0.elapsed_time(t1))
    ctx["a2a_bytes"].append(8 * (n - counts[ctx["rank"]]))
    km.add_canonical_dev(recv.data_ptr(), sum(recv_counts))


# ------------------------------------------------------------------------------------------------
# fused exchange over NVLink peer memory
def p2p_offsets(count_matrix, rank):
    """count_matrix[src][dst] = keys rank src sends to owner dst.  -> (where in every owner's receive buffer this
    rank's run starts [per dst], how many keys this rank receives in total).  Runs are laid out in source order."""
    M = np.asarray(count_matrix, dtype=np.int64)
    return [int(M[:rank, p].sum()) for p in range(M.shape[1])], int(M[:, rank].sum())


class P2PExchange(object):
    """Routing and transfer in ONE kernel (csrc/extract.cu route_p2p_kernel): every rank maps every other rank's
    receive buffers (CUDA IPC) and the bucketing kernel stores each key straight into its owner's buffer over
    NVLink.  NCCL carries only the tiny count matrix (so that ranks agree on their slots) and the all-reduce that says
    "all stores have landed" (and whether anybody overflowed).

    Steps are numbered: exchange(km, seq) uses receive buffer seq % nbuf, and exchanges are issued in the order of their
    numbers on every rank -- a host thread whose step's turn has not come waits (`_turn`), so several steps can be in
    flight per rank (one host thread each: the H2D copy / extraction of step s + 1 and the sort of step s - 1 overlap
    the exchange of step s) while the collectives still match up across ranks.  Why nbuf = steps in flight + 1 is
    enough: peers store into buffer s % nbuf after the closing all-reduce of step s - 1, which this rank entered only
    after the thread that runs step s - 1 had finished ITS previous step, s - 1 - inflight, completely (sorted); the
    buffer's previous user, step s - nbuf, is no later than that.  A single kmerizer that is streamed through many
    exchanges (tools/human_scale.py, bench.py's human leg) consumes the previously adopted buffer before it enters the
    closing all-reduce (`consume=True`), so two buffers are enough there as well.

    reserve (ZB_P2P_RESERVE=1 or reserve=True): no count matrix at all -- every receive buffer ends in a cursor word,
    and a thread block of the routing kernel reserves its run in the owner's buffer with one system-scope atomic add on
    that word (over NVLink for a remote owner).  That drops the owner-count pass over the keys, the all-gather and two
    host round trips from every step; after the closing all-reduce the owner reads its own cursor = keys received.  The
    runs land in timing order, which the owner's sort makes irrelevant."""

    def __init__(self, nat, dist, rank, world, dev, capacity_keys, reserve=None, nbuf=2):
        import os
        import threading
        import torch
        self.nat, self.dist, self.rank, self.world, self.dev = nat, dist, rank, world, dev
        self.capacity = int(capacity_keys)
        self.reserve = (os.environ.get("ZB_P2P_RESERVE", "0") == "1") if reserve is None else bool(reserve)
        self.nbuf = max(2, int(nbuf))
        self.step = 0                  # next sequence number to be exchanged
        self._turn = threading.Condition()
        self.bufs = []
        self.route_ms = []
        self.remote_bytes = []
        self.cnt = torch.empty(world, dtype=torch.int64, device="cuda:%d" % dev)
        self.allc = torch.empty(world * world, dtype=torch.int64, device="cuda:%d" % dev)
        self.flag = torch.zeros(1, dtype=torch.int64, device="cuda:%d" % dev)
        for _ in range(self.nbuf):
            ptr, handle = nat.ipc_alloc(self.capacity * 8 + 256, dev)     # the cursor word sits behind the keys
            _as_tensor(ptr + self.capacity * 8, 1, torch.int64, dev).zero_()
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            ptrs = [ptr if r == rank else nat.ipc_open(handles[r], dev) for r in range(world)]
            self.bufs.append((ptr, ptrs))
        torch.cuda.synchronize(dev)
        dist.barrier()

    def prepare(self, km):
        """call on a fresh kmerizer BEFORE it is fed: its extraction then tallies the keys per owner, and the exchange needs
        no counting pass over them"""
        km.set_owners(self.world)
        return km

    def _landed(self, failed):
        """closing collective of a step: returns once every rank's stores (and reservations) have completed; raises on
        EVERY rank when any rank overflowed a receive buffer (a rank that raised alone would leave the others waiting)"""
        self.flag.fill_(1 if failed else 0)
        self.dist.all_reduce(self.flag, op=self.dist.ReduceOp.MAX)
        if int(self._host(self.flag, 1)[0]):
            raise RuntimeError("P2PExchange: a receive buffer of %d keys overflowed on some rank" % self.capacity)

    def _host(self, tensor, n):
        """n int64 of a device tensor that a collective on torch's stream has just produced -> numpy.  Not `.cpu()` /
        `.item()`: those are copy-engine transfers and queue behind the bulk H2D / D2H copies that the other steps in
        flight have running (64 MiB pieces: milliseconds per exchange, measured at N = 2)."""
        import torch
        torch.cuda.current_stream(self.dev).synchronize()
        return self.nat.dev_read_small(tensor.data_ptr(), n, np.int64, self.dev)

    def _exchange_reserve(self, km, seq, consume):
        import time
        import torch
        own, ptrs = self.bufs[seq % self.nbuf]
        t0 = time.perf_counter()
        failed = False
        sent = [0] * self.world
        try:
            sent = km.route_p2p_reserve(ptrs, [p + self.capacity * 8 for p in ptrs], self.capacity)
        except IndexError:           # ZB_E_RANGE: a reservation passed the capacity (nothing was stored out of bounds)
            failed = True
        self.route_ms.append((time.perf_counter() - t0) * 1e3)
        self.remote_bytes.append(8 * (sum(sent) - sent[self.rank]))
        if consume:
            km.flush()               # the buffer adopted in the previous step is sorted before peers may move on
        cur = _as_tensor(own + self.capacity * 8, 1, torch.int64, self.dev)
        self._landed(failed)         # every rank's stores and reservations have completed
        nrecv = int(self._host(cur, 1)[0])
        cur.zero_()                  # peers touch this buffer's cursor again nbuf steps from now
        torch.cuda.current_stream(self.dev).synchronize()
        if nrecv > self.capacity:    # cannot happen after a clean _landed; never read past the buffer
            raise RuntimeError("P2PExchange: cursor %d beyond the capacity %d" % (nrecv, self.capacity))
        km.adopt_canonical_dev(own, nrecv)

    def _exchange_counts(self, km, seq, consume):
        import time
        import torch
        counts = km.bucket_counts(self.world)
        self.nat.dev_write_small(self.cnt.data_ptr(), np.array(counts, dtype=np.int64), self.dev)
        self.dist.all_gather_into_tensor(self.allc, self.cnt)
        M = self._host(self.allc, self.world * self.world).reshape(self.world, self.world)
        offs, nrecv = p2p_offsets(M, self.rank)
        if int(M.sum(axis=0).max()) > self.capacity:      # the same matrix on every rank: everybody raises
            raise RuntimeError("P2PExchange: a rank would receive %d keys, capacity is %d" % (int(M.sum(axis=0).max()), self.capacity))
        own, ptrs = self.bufs[seq % self.nbuf]
        t0 = time.perf_counter()
        dst = [ptrs[p] + 8 * offs[p] for p in range(self.world)]
        if consume:
            # one kmerizer streamed through many exchanges: this batch's keys travel over NVLink (second stream) while the
            # batch the previous exchange delivered is sorted and counted (first stream)
            km.route_p2p_begin(dst)
            km.flush()
            km.route_p2p_end()
        else:
            km.route_p2p(dst)
        self.route_ms.append((time.perf_counter() - t0) * 1e3)
        self.remote_bytes.append(8 * (sum(counts) - counts[self.rank]))
        self._landed(False)          # every rank's stores have completed
        km.adopt_canonical_dev(own, nrecv)    # sorted in place by km.finish() / the next flush

    def exchange(self, km, seq=None, consume=True):
        """route km's pending canonical k-mers to their owners and adopt what this rank received.  seq: the step's
        number (None: the next one -- a single-threaded caller).  consume=False: the caller promises that a step's
        buffer is sorted (km.finish()) before step seq + nbuf - 1 is exchanged -- one kmerizer per step, at most
        nbuf - 1 steps in flight."""
        import torch
        with self._turn:
            if seq is None:
                seq = self.step
            while self.step != seq:
                self._turn.wait()
        torch.cuda.set_device(self.dev)      # collectives below run on this host thread's current device
        try:
            if self.reserve:
                self._exchange_reserve(km, seq, consume)
            else:
                self._exchange_counts(km, seq, consume)
        finally:
            with self._turn:
                self.step = seq + 1
                self._turn.notify_all()

    def close(self):
        for (own, ptrs) in self.bufs:
            for r, p in enumerate(ptrs):
                if r != self.rank:
                    self.nat.ipc_close(p, self.dev)
        self.dist.barrier()
        for (own, _) in self.bufs:
            self.nat.ipc_free(own, self.dev)
        self.bufs = []


def gather_counted_set(nat, kset, dist, rank, world, dev, root=0):
    """One sorted file from hash-owned shares: every rank sends its counted both-strand share (sorted, disjoint
    from all others) to `root`, which merges them with zb_merge (merge-path tree) into THE sorted set -- the input
    of the file writer, so the output is byte-identical to a single-GPU `zot kmerize` (codec64's greedy word
    boundaries depend on the whole sequence, so the encode has to see one array).  Returns the merged KmerSet on
    root, None elsewhere.  Volume: 12 B per distinct k-mer, far less than the key exchange at 30x coverage.
    (Sets beyond one GPU's memory need the range-partitioned variant, SURVEY.md 8e(ii): not built.)"""
    import torch
    n = len(kset)
    sizes = [None] * world
    dist.all_gather_object(sizes, n)
    kp, cp = kset.dev_ptrs()
    dv = "cuda:%d" % dev
    # wrap the set's device arrays as tensors without copying (torch only does the plumbing)
    k_local = _as_tensor(kp, n, torch.int64, dev)
    c_local = _as_tensor(cp, n, torch.int32, dev)
    if rank == root:
        shares = [kset]
        reqs = []
        bufs = []
        for r in range(world):
            if r == root:
                continue
            kb = torch.empty(max(sizes[r], 1), dtype=torch.int64, device=dv)
            cb = torch.empty(max(sizes[r], 1), dtype=torch.int32, device=dv)
            if sizes[r]:
                reqs.append(dist.irecv(kb[:sizes[r]], src=r))
                reqs.append(dist.irecv(cb[:sizes[r]], src=r))
            bufs.append((r, kb, cb))
        for q in reqs:
            q.wait()
        torch.cuda.synchronize(dev)
        for (r, kb, cb) in bufs:
            shares.append(nat.KmerSet.from_device(kb.data_ptr(), cb.data_ptr(), sizes[r], dev))
        merged = nat.merge(shares) if len(shares) > 1 else kset
        for sh in shares[1:]:
            sh.free()
        return merged
    if n:
        dist.send(k_local, dst=root)
        dist.send(c_local, dst=root)
    torch.cuda.synchronize(dev)
    return None


def merge_sharded(nat, sets, dist, rank, world, dev, root=0):
    """`zot merge` over several GPUs: the key space is cut into `world` ranges at splitters taken from
    the quantiles of the inputs, rank r merges range r of EVERY input (no exchange during the merge: every rank
    holds the inputs, e.g. decoded from the same files), and the ranges -- sorted and disjoint by construction --
    are concatenated on `root`.  Returns the merged KmerSet on root, None elsewhere; identical to nat.merge(sets)."""
    import torch
    # splitters: equally spaced order statistics of the largest input (identical on every rank)
    big = max(sets, key=len)
    nb = len(big)
    if world > 1 and nb >= world:
        pos = np.array([(nb * r) // world for r in range(1, world)], dtype=np.int64)
        kp, _ = big.dev_ptrs()
        keys = _as_tensor(kp, nb, torch.int64, dev)
        split = keys[torch.from_numpy(pos).to(keys.device)].cpu().numpy().view(np.uint64)
    else:
        split = np.zeros(0, np.uint64)
    lo = None if rank == 0 or len(split) == 0 else split[rank - 1]
    hi = None if rank == world - 1 or len(split) == 0 else split[rank]
    if len(split) == 0 and rank != 0:
        lo = hi = None
    parts = []
    for s in sets:
        n = len(s)
        if len(split) == 0:
            b, e = (0, n) if rank == 0 else (0, 0)
        else:
            probes = [x for x in (lo, hi) if x is not None]
            idx = s.lower_bound(np.array(probes, np.uint64)) if probes else []
            b = 0 if lo is None else int(idx[0])
            e = n if hi is None else int(idx[-1])
        parts.append(s.slice(b, e))
    mine = nat.merge(parts) if len(parts) > 1 else parts[0]
    for p in parts:
        if p is not mine:
            p.free()
    # concatenate the ranges on root, in rank order
    n = len(mine)
    sizes = [None] * world
    dist.all_gather_object(sizes, n)
    dv = "cuda:%d" % dev
    kp, cp = mine.dev_ptrs()
    k_local = _as_tensor(kp, n, torch.int64, dev)
    c_local = _as_tensor(cp, n, torch.int32, dev)
    out = None
    if rank == root:
        total = sum(sizes)
        kb = torch.empty(max(total, 1), dtype=torch.int64, device=dv)
        cb = torch.empty(max(total, 1), dtype=torch.int32, device=dv)
        off = 0
        reqs = []
        for r in range(world):
            if r == root:
                kb[off:off + n].copy_(k_local)
                cb[off:off + n].copy_(c_local)
            elif sizes[r]:
                reqs.append(dist.irecv(kb[off:off + sizes[r]], src=r))
                reqs.append(dist.irecv(cb[off:off + sizes[r]], src=r))
            off += sizes[r]
        for q in reqs:
            q.wait()
        torch.cuda.synchronize(dev)
        out = nat.KmerSet.from_device(kb.data_ptr(), cb.data_ptr(), total, dev)
    elif n:
        dist.send(k_local, dst=root)
        dist.send(c_local, dst=root)
    torch.cuda.synchronize(dev)
    mine.free()
    return out


def _as_tensor(ptr, n, dtype, dev):
    """a torch view of n elements of library-owned device memory (no copy)"""
    import torch

    class _Arr(object):
        pass
    if n == 0:
        return torch.empty(0, dtype=dtype, device="cuda:%d" % dev)
    a = _Arr()
    itemsize = torch.empty(0, dtype=dtype).element_size()
    typestr = {8: "<i8", 4: "<i4"}[itemsize]
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device="cuda:%d" % dev)


# ------------------------------------------------------------------------------------------------
# all-pairs distance matrix: work units = (pair of blocks of 32 sets) x (one of 8 key-range shards), sharded over ranks
AP_S = 32   # sets per block, csrc/allpairs.cu AB_S
AP_KS = 8   # key-range shards per tile, csrc/allpairs.cu AB_KS


def tile_blocks(nsets, u):
    """work unit -> (bi, bj), bi <= bj: its tile u // AP_KS.  More than two blocks: row-major over the upper triangle
    of blocks, diagonal included; up to two blocks (<= 64 sets): ONE tile (0, nblk - 1) that covers every pair
    (host statement of csrc/allpairs.cu tile_to_blocks)"""
    t = u // AP_KS
    nblk = -(-nsets // AP_S)
    if nblk <= 2:
        return 0, nblk - 1
    r, start = 0, 0
    while start + (nblk - r) <= t:
        start += nblk - r
        r += 1
    return r, r + (t - start)


def tile_pairs(nsets, u):
    """the set pairs (i < j) a work unit contributes to (it holds their counts over ITS key-range shard u % AP_KS)"""
    nblk = -(-nsets // AP_S)
    if nblk <= 2:
        return [(i, j) for i in range(nsets) for j in range(i + 1, nsets)]
    bi, bj = tile_blocks(nsets, u)
    out = []
    for i in range(bi * AP_S, min((bi + 1) * AP_S, nsets)):
        for j in range(max(bj * AP_S, i + 1), min((bj + 1) * AP_S, nsets)):
            out.append((i, j))
    return out


def key_shard(x, key_bits):
    """key-range shard of k-mer x when the largest k-mer of the collection has key_bits bits: floor(x AP_KS / 2^key_bits)"""
    return (int(x) * AP_KS) >> key_bits


def n_tiles(nsets):
    nblk = -(-nsets // AP_S)
    return (1 if nblk <= 2 else nblk * (nblk + 1) // 2) * AP_KS


def tile_ranges(nsets, world):
    """contiguous tile ranges [(begin, end)] per rank with about the same number of set pairs each"""
    nt = n_tiles(nsets)
    w = np.array([len(tile_pairs(nsets, t)) for t in range(nt)], dtype=np.int64) if nt < 20000 else np.full(nt, 1024)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(nt)
    cuts = [min(max(c, 0), nt) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def unit_share(nsets, rank, world):
    """(begin, end, stride) of rank's work units: every world-th unit, starting at `rank`.  A rank so holds key-range
    shards of EVERY tile (with 8 ranks: shard `rank` of each), and the ranks stay level however unevenly the tiles cost --
    contiguous ranges (tile_ranges) left the ranks holding the more diverged sets of 1,000 35 % behind the others
    (profiles/r01_allpairs.md).  More ranks than units: the surplus ranks get nothing."""
    return rank, n_tiles(nsets), max(1, world)


def share_units(nsets, rank, world):
    """the units of unit_share, listed"""
    b, e, st = unit_share(nsets, rank, world)
    return list(range(b, e, st))


def pair_index(nsets, i, j):
    """row-major index of pair (i < j) in the upper triangle"""
    return i * (2 * nsets - i - 1) // 2 + (j - i - 1)


def allpairs_sharded(compute_tiles, nsets, dist, rank, world, device="cpu"):
    """every rank computes its share of the work units with compute_tiles(begin, end, stride) -> uint64 [npairs, 3]
    (zeros outside its units; zotmer_b200._native.allpairs_abc on a GPU), then one all-reduce adds the shards up.
    Returns the full matrix on every rank."""
    import torch
    b, e, st = unit_share(nsets, rank, world)
    part = compute_tiles(b, e, st)
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.int64).copy()).to(device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64)
