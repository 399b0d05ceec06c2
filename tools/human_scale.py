"""
BASELINE.json config[4]: human-scale synthetic genome at 30x 150 bp reads, kmerize+count k=31, canonical k-mers routed to
their hash-range owner over NVLink, N = 1/2/4/8 B200 (weak scaling: 375 Mbp of genome and 75,000,000 reads per GPU, so
N = 8 is the 3 Gbp / 600 M reads / 90 Gbases instance).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/human_scale.py [--scale S]

A 200 GB FASTQ cannot be staged through gpurun, so the reads are generated on the device, batch by batch,
as base codes (torch: data generation is plumbing, it is NOT inside the timed sections) and fed through
zb_kmerize_feed_codes_dev; everything timed is the library: extraction, exchange (route_p2p over peer memory), sort +
count of the received keys, merge into the rank's running counted set, then mirror / finish and trim -c 2.
--scale S divides genome and reads by S (S = 1000 is the bit-exactly checked down-scale of tests/test_gpu_multi.py's
k = 31 case; here the result is checked through invariants: sum of counts = 2 x windows, strict order per rank,
sum(hist c * freq) = sum of counts, per-rank shares disjoint by ownership).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from zotmer_b200 import _native as nat  # noqa: E402
from zotmer_b200 import multigpu  # noqa: E402

K = 31
L = 150
GENOME_PER_RANK = 375000000
READS_PER_RANK = 75000000
ERR = 0.001


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--batch-reads", type=int, default=3000000)   # 3 M x 151 codes < 2^29: one pending batch, never a local flush
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    dv = "cuda:%d" % dev
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dv))

    G = GENOME_PER_RANK * world // args.scale
    nreads = READS_PER_RANK // args.scale
    B = min(args.batch_reads, nreads)

    # ---- the genome (the same on every rank): i.i.d. ACGT + 5 % of its length in repeat families
    gen = torch.Generator(device=dv)
    gen.manual_seed(5)
    genome = torch.randint(0, 4, (G,), dtype=torch.uint8, device=dv, generator=gen)
    rep_len = max(200, min(6000, G // 1000))
    fam = torch.randint(0, 4, (20, rep_len), dtype=torch.uint8, device=dv, generator=gen)
    ncopies = int(0.05 * G / rep_len)
    where = torch.randperm(G // rep_len - 1, device=dv, generator=gen)[:ncopies] * rep_len   # disjoint: one writer per base
    which = torch.randint(0, 20, (ncopies,), device=dv, generator=gen)
    for c0 in range(0, ncopies, 4096):
        w = where[c0:c0 + 4096]
        idx = (w[:, None] + torch.arange(rep_len, device=dv)[None, :]).reshape(-1)
        genome[idx] = fam[which[c0:c0 + 4096]].reshape(-1)
    del where, which, fam
    torch.cuda.synchronize(dev)

    p2p = None
    if world > 1:
        cap = int(B * (L - K + 1) * 1.15) + (1 << 20)
        p2p = multigpu.P2PExchange(nat, dist, rank, world, dev, cap)

    def barrier():
        torch.cuda.synchronize(dev)
        nat.device_sync(dev)
        if world > 1:
            dist.barrier()

    rgen = torch.Generator(device=dv)
    ar = torch.arange(L, device=dv, dtype=torch.int64)

    def make_batch(b, seed):
        """b reads as base codes [b, L + 1] (last column = 4, the record break): uniform start, strand 50/50,
        substitution errors at ERR per base"""
        rgen.manual_seed(seed)
        pos = torch.randint(0, G - L, (b,), device=dv, generator=rgen)
        codes = torch.empty((b, L + 1), dtype=torch.uint8, device=dv)
        rd = genome[(pos[:, None] + ar[None, :]).reshape(-1)].reshape(b, L)
        rev = torch.rand(b, device=dv, generator=rgen) < 0.5
        rd = torch.where(rev[:, None], 3 - rd.flip(1), rd)
        e = torch.rand((b, L), device=dv, generator=rgen) < ERR
        sub = torch.randint(1, 4, (b, L), dtype=torch.uint8, device=dv, generator=rgen)
        rd = torch.where(e, (rd + sub) & 3, rd)
        codes[:, :L] = rd
        codes[:, L] = 4
        return codes

    # ---- warm-up: one batch through a throw-away kmerizer (device allocator, peer mappings, kernel attributes)
    kw = nat.Kmerizer(K, dev)
    for it in range(2):
        codes = make_batch(B, 7 + it)
        kw.feed_codes_dev(codes.data_ptr(), codes.numel(), B)
        if p2p is not None:
            p2p.exchange(kw)
        nat.device_sync(dev)
        del codes
    sw, _ = kw.finish()
    kw.close()
    sw.trim(2).free()
    sw.free()
    barrier()

    km = nat.Kmerizer(K, dev)
    timed = 0.0
    fed = 0
    windows = 0
    nb = 0
    while fed < nreads:
        b = min(B, nreads - fed)
        # ---- generate one batch of reads on the device   (NOT timed)
        codes = make_batch(b, 1000003 * (rank + 1) + nb)
        barrier()
        # ---- timed: extraction, exchange, sort + count of what the previous exchange delivered
        t0 = time.perf_counter()
        km.feed_codes_dev(codes.data_ptr(), codes.numel(), b)
        if p2p is not None:
            p2p.exchange(km)
        nat.device_sync(dev)
        timed += time.perf_counter() - t0
        del codes
        fed += b
        windows += b * (L - K + 1)
        nb += 1
        if rank == 0 and (nb % 5 == 0 or fed == nreads):
            print("rank 0: %d / %d reads, %.3f s in the library so far" % (fed, nreads, timed), file=sys.stderr, flush=True)
    barrier()
    torch.cuda.empty_cache()                   # the generator's scratch goes back to the driver before the big allocations
    free_b, total_b = torch.cuda.mem_get_info(dev)
    t_feed = timed
    t0 = time.perf_counter()
    s, nr = km.finish()
    km.close()
    nat.device_sync(dev)
    t1 = time.perf_counter()
    t = s.trim(2)
    nat.device_sync(dev)
    t2 = time.perf_counter()
    timed += t2 - t0
    print("rank %d: feed + exchange + count %.3f s, finish (last count, mirror, merge) %.3f s, trim %.3f s; %.1f GB free before finish" % (
        rank, t_feed, t1 - t0, t2 - t1, free_b / 1e9), file=sys.stderr, flush=True)
    barrier()

    # ---- invariants
    st = s.stats()
    kp, cp = s.dev_ptrs()
    n = len(s)
    ok_sorted = True
    step = 1 << 27
    kt = multigpu._as_tensor(kp, n, torch.int64, dev)
    for a in range(0, n - 1, step):            # keys < 2^62: signed comparison is fine
        bnd = min(n - 1, a + step)
        ok_sorted = ok_sorted and bool((kt[a + 1:bnd + 1] > kt[a:bnd]).all())
    hist_total = sum(int(c) * int(f) for c, f in st["hist"])
    vals = torch.tensor([float(st["total"]), float(windows), float(n), float(len(t)), timed, float(nr)], dtype=torch.float64, device=dv)
    tmax = torch.tensor([timed], dtype=torch.float64, device=dv)
    flags = torch.tensor([1.0 if (ok_sorted and hist_total == st["total"]) else 0.0], dtype=torch.float64, device=dv)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    total, win, ndist, ntrim, _, nrec = [float(x) for x in vals.tolist()]
    assert flags.item() == 1.0, "a rank's share is not strictly ascending or its histogram does not add up"
    assert total == 2.0 * win, "sum of counts %r != 2 x windows %r" % (total, win)
    assert nrec == nreads * world
    if rank == 0:
        bases = float(nreads) * L * world
        out = {"config": "config[4]: %.3f Gbp genome, %d reads x %d bp (%.1f Gbases), k=%d, %d GPU(s)%s" % (
                   G / 1e9, nreads * world, L, bases / 1e9, K, world, "" if args.scale == 1 else " [1/%d scale]" % args.scale),
               "n_gpus": world, "seconds_in_library_max_over_ranks": tmax.item(),
               "Gbases_per_s": bases / tmax.item() / 1e9, "distinct_kmers_both_strands": int(ndist), "after_trim_c2": int(ntrim),
               "sum_of_counts": int(total), "batches_per_rank": nb, "batch_reads": B,
               "checked": "sum of counts = 2 x windows; every rank's share strictly ascending; sum(hist c * freq) = sum of counts"}
        print(json.dumps(out))
    s.free()
    t.free()
    if p2p is not None:
        p2p.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
