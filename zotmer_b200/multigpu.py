"""
Multi-GPU kmerize+count: hash-range routing of canonical k-mers (one process per GPU, torch.distributed).

The reference has no parallelism at all; this is the B200 layer BASELINE.json's north star asks for:
reads are sharded across ranks, every rank extracts the canonical k-mers of its shard, each k-mer is sent
to its OWNER rank -- owner(x) = floor(mix64(x) * nranks / 2^64), the high bits of an invertible 64-bit mix,
so ownership is uniform even on low-complexity sequence -- with ONE all-to-all over NVLink (sizes first),
and every rank sorts and counts its disjoint share locally.  x and rc(x) land on different owners, which
is fine: the canonical key is what is routed, mirroring happens after counting on the owner.

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
