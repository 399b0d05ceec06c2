#!/usr/bin/env python
"""
bench.py -- measures BASELINE.json's metric "kmerize+count Gbases/s at 1/2/4/8 B200; pairwise Jaccard set-pairs/s".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (`value`, `e2e`): BASELINE.json configs[1] -- synthetic 30x 150 bp Illumina-like FASTQ of a 5 Mbp genome
(1,000,000 reads = 150 Mbases per GPU), kmerize+count at k=25, then zot trim at min-count 2.  One "step" = one pass of
the hot path over the whole read set of a rank, doing what `zot kmerize` + `zot trim` do up to the file write (and what
the reference arm does on the CPU): parse, extract both strands, sort, count, count histogram + acgt tallies, codec64
(+ delta) encode of the counted set, trim, codec64 encode of the trimmed set -- all through the C ABI.
  value : FASTQ text already resident in HBM (zb_kmerize_feed_dev), results left in HBM; CUDA events.  Measured twice over
          the same K steps: one after the other on one stream (`one_in_flight`: the region the per-stage times and the
          `roofline` of the kernels come from), and with several steps in flight, one host thread and one stream each --
          the host round trips of a step and, at N > 1, its exchange over NVLink then run under the kernels of the others.
          `value` is the faster of the two (`steps_in_flight` says which); every step does the whole work either way.
  e2e   : FASTQ text in pinned HOST memory (zb_kmerize_feed), the encoded word streams of the trimmed set fetched into
          pinned host memory (zb_words_fetch) -- the bytes `zot trim` would write -- H2D and D2H inside the timed region;
          several steps in flight (one host thread each), at N > 1 as well: exchanges are issued in step order.
N > 1 (weak scaling): every rank kmerizes its own 1M-read shard, routes each canonical k-mer to its owner rank (high
bits of a 64-bit mix) with one fused routing kernel that stores into the owner's buffer over NVLink peer memory, and
sorts / counts / trims its disjoint share.  Before anything is timed the ranks run a small multi-GPU kmerize (k = 25 and
k = 31) and compare the union of their shares with the oracle, bit for bit (`mgpu_parity`).
  e2e_bgzf : the e2e step fed from the same reads as a bgzip'd file (BGZF) in pinned host memory: the compressed bytes
          cross PCIe and are inflated on the device (informational; `e2e` and the reference arm read plain text).
Further objects of the JSON line (each skippable): `human` = configs[4]'s per-GPU shape (375 Mbp of genome, 75 M reads,
k = 31; reads generated on the device); `pairs` = configs[3] (all pairs of 1,000 sets, 499,500 pairs); `cli` = the
`zot kmerize` command itself, FASTQ file in -> k-mer set file out, wall clock.
--impl reference: the reference's own algorithm (pure Python, single thread, as the reference is) timed on a bounded
sample of the same workload on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K = 25
READ_LEN = 150
READS_PER_RANK = int(os.environ.get("ZB_BENCH_READS", 1000000))
GENOME = 5000000
METRIC = "kmerize+count Gbases/s"
UNIT = "Gbases/s"
# algorithmic bytes per input base of kmerize+count at k=25 on this workload
# (1 + 2*W*8*(3+2P) + 12*d with W=0.84, P=7, d~0.3) and per key per radix pass (8 read + 8 write)
BYTES_PER_BASE = 233.0
BYTES_PER_KEY_PASS = 16.0
# configs[4] (k=31, 8 passes, W=0.8, d~0.13): SURVEY.md 8d
HUMAN_BYTES_PER_BASE = 246.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons sampled in-process through NVML (the library nvidia-smi uses) every
    20 ms while the timed region runs.  Spawning nvidia-smi itself every 200 ms stalls CUDA API calls
    for milliseconds at a time (measured: 104 ms/step with it, 20 ms/step without)."""

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.sm, self.reasons, self.max_sm = [], set(), None
        self.stop_flag = threading.Event()
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(self.gpu_index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception as e:
                self.err = repr(e)
                break
            self.stop_flag.wait(0.02)

    def stop(self):
        self.stop_flag.set()
        if self.t is not None:
            self.t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples: %s" % self.err]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": "NVML in-process, 20 ms period, during the timed steps (device-resident and e2e regions)"}


def physical_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            pass
    return local_index


_GPU_CPUS = {}


def gpu_cpus(gpu_index):
    """the CPUs next to a GPU (NVML's affinity mask) that this process may run on, or None"""
    if gpu_index not in _GPU_CPUS:
        pick = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(gpu_index))
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = set(i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1)
            allowed = os.sched_getaffinity(0)
            if cpus & allowed and (cpus & allowed) != allowed:
                pick = cpus & allowed
        except Exception:
            pick = None
        _GPU_CPUS[gpu_index] = pick
    return _GPU_CPUS[gpu_index]


class near_gpu(object):
    """Context manager: while pinned host buffers are allocated and first touched, the calling thread runs on the CPUs
    next to its GPU (NVML's affinity mask), so that the pages land in that socket's memory -- with 8 ranks on one host
    the H2D rate per rank halved in round 1 (25 GB/s instead of 52).  The mask is restored afterwards; nothing else is
    pinned down.  Does nothing when NVML or the mask is unavailable.  One instance per `with` (threads use their own)."""

    def __init__(self, gpu_index):
        self.pick = gpu_cpus(gpu_index)
        self.saved = None

    def __enter__(self):
        if self.pick:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.pick)
            except Exception:
                self.saved = None
        return self

    def __exit__(self, *a):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def make_reads(rank, nreads, world=1):
    """rank's shard of the read set.  Weak scaling keeps the COVERAGE fixed as well as the reads per GPU: N ranks
    sample a genome of N x 5 Mbp (config[4]'s shape -- the genome grows with the machine), so every owner rank
    sees the same k-mer statistics (30x, ~10 M distinct genomic k-mers) as the single-GPU run."""
    from tools import synth
    g = synth.genome(GENOME * world, seed=17)
    return synth.fastq_array(g, nreads, L=READ_LEN, seed=18 + 1000 * rank).reshape(-1)


# ------------------------------------------------------------------------------------------------
def reference_step(zo, fq):
    """what the reference does for `zot kmerize` + `zot trim` up to the file write (oracle port)"""
    xs, cs, h, acgt, nr = zo.kmerize_core(K, [("reads.fq", fq)])
    zo.words_to_bytes(zo.encode(zo.delta(xs)))       # the reference writes the set (54 % of its time)
    zo.words_to_bytes(zo.encode(cs))
    tx, tc = zo.trim_core(xs, cs, 2, None)
    zo.words_to_bytes(zo.encode(zo.delta(tx)))
    zo.words_to_bytes(zo.encode(tc))
    return len(xs)


def run_reference(args, rank):
    """The reference's own CPU path (pure Python, single thread): oracle port timed on a bounded sample."""
    if rank != 0:
        return
    from oracle import zot_oracle as zo
    sample_reads = int(os.environ.get("ZB_REF_SAMPLE_READS", 4000))
    fq = make_reads(0, READS_PER_RANK)[:sample_reads * 315].tobytes()
    bases = sample_reads * READ_LEN
    for _ in range(args.warmup):
        reference_step(zo, fq)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(zo, fq)
    dt = (time.perf_counter() - t0) / args.steps
    val = bases / dt / 1e9
    sample = "first %d of the %d reads (%d bases) per step; CPython %s, 1 thread" % (
        sample_reads, READS_PER_RANK, bases, sys.version.split()[0])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n):
    return {"workload": "config[1]: synthetic 30x 150bp FASTQ of a %d Mbp genome, %d reads per GPU, kmerize+count k=25 "
                        "(count histogram, acgt, codec64 encode of the set), then trim min-count 2 (+ encode)" % (
                            GENOME * n // 1000000, READS_PER_RANK),
            "k": K, "reads_per_gpu": READS_PER_RANK, "read_len": READ_LEN, "bases_per_gpu": READS_PER_RANK * READ_LEN,
            "genome_bp": GENOME * n,
            "parallelism": "1 GPU" if n == 1 else "%d GPUs: %d reads per GPU from a %d Mbp genome (30x), canonical k-mers "
                                                  "routed to their hash-range owner over NVLink" % (n, READS_PER_RANK, GENOME * n // 1000000),
            "l2": "inputs (315 MB of text, 1 GB of keys per GPU) are larger than the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
def finish_step(km, fetch=None):
    """the rest of a step once the rank's keys are in the kmerizer: count, stats, encode, trim, encode
    -> (n distinct, n after trim, stats, words of the full set, words of the trimmed set)"""
    s, nr = km.finish()
    km.close()
    st = s.stats()            # count histogram (first-occurrence order) + acgt: part of every `zot kmerize`
    w = s.encode_dev()        # the two streams `zot kmerize` writes
    t = s.trim(2)
    tw = t.encode_dev()       # the two streams `zot trim` writes
    out = (len(s), len(t), st, w.sizes(), tw.sizes())
    if fetch is not None:
        # the pipeline's product leaves the device: the streams of the trimmed set (what `zot trim` writes).  The full
        # set's streams are encoded as well (the reference arm encodes all four) but stay in HBM: with 8 ranks on one
        # host the copies in and out share ~200 GB/s of host memory bandwidth, and 203 MB of intermediate result per
        # step would cost more than the 315 MB of input (measured at N = 2: D2H 8.1 ms/step against H2D 7.1)
        tw.fetch(fetch[0], fetch[1])
    for x in (w, tw, s, t):
        x.free()
    return out


def step_device(nat, dev, d_ptr, nbytes, p2p):
    """one step with inputs resident in HBM"""
    km = nat.Kmerizer(K, dev)
    if p2p is not None:
        p2p.prepare(km)
    km.feed_dev(d_ptr, nbytes, False)
    if p2p is not None:
        p2p.exchange(km, consume=False)
    return finish_step(km)


# ------------------------------------------------------------------------------------------------
def mgpu_parity(nat, rank, world, dev):
    """tools/mgpu_check.py's comparison, before anything is timed: a small read set sharded over the ranks, default
    exchange, k = 25 and k = 31 (configs[4]'s k: 62-bit keys); the union of the ranks' counted shares must be the
    oracle's kmerize of all the reads, bit for bit.  The oracle is only the checker here."""
    import torch
    import torch.distributed as dist
    from zotmer_b200 import multigpu
    from tools import synth
    from oracle import c_oracle as co
    g = synth.genome(300000, seed=5)
    nreads = 20000
    shards = [synth.fastq_array(g, nreads, seed=50 + r).reshape(-1).tobytes() for r in range(world)]
    out = {}
    for k in (25, 31):
        p2p = multigpu.P2PExchange(nat, dist, rank, world, dev, (READ_LEN - k + 1) * nreads * 2)
        ok = True
        expect = co.kmerize(k, [(sh, False) for sh in shards])[:2] if rank == 0 else None
        for it in range(3):          # three steps: the receive buffers are reused
            km = p2p.prepare(nat.Kmerizer(k, dev))
            km.feed(shards[rank], False)
            p2p.exchange(km)
            s, nr = km.finish()
            km.close()
            ks, cs = s.fetch()
            s.free()
            parts = [None] * world
            dist.all_gather_object(parts, (ks, cs))
            if rank == 0 and it in (0, 2):
                gk = np.concatenate([p[0] for p in parts])
                gc = np.concatenate([p[1] for p in parts])
                order = np.argsort(gk, kind="stable")
                ek, ec = expect
                ok = ok and len(np.unique(gk)) == len(gk) and np.array_equal(gk[order], ek) and np.array_equal(gc[order], ec)
        p2p.close()
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda:%d" % dev)
        dist.broadcast(flag, src=0)
        out["k%d" % k] = "ok" if int(flag.item()) else "MISMATCH"
    out["how"] = ("%d ranks x %d reads of a 300 kbp genome, exchange %s, 3 steps; union of the per-rank counted shares == oracle "
                  "kmerize of all reads (k-mers and counts, bit for bit)" % (
                      world, nreads, "p2p_reserve" if os.environ.get("ZB_P2P_RESERVE", "0") == "1" else "p2p"))
    if any(v == "MISMATCH" for v in out.values()):
        raise SystemExit("bench.py: multi-GPU parity FAILED: %r" % out)
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
    # ... and since NCCL still prints its version banner to file descriptor 1 from C, everything written to stdout
    # while the bench runs goes to stderr; the JSON line alone is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from zotmer_b200 import _native as nat
    if nat.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (zotmer_b200 has no CPU path)")
    dev = local_rank
    torch.cuda.set_device(dev)
    dist = None
    p2p = None
    parity = None
    inflight = int(os.environ.get("ZB_E2E_INFLIGHT", 6))   # measured: r3_bench6_f*.json (N = 1), r3_bench_n2_f*.json (N = 2)
    if world > 1:
        import torch.distributed as dist
        from zotmer_b200 import multigpu
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % dev))
        parity = mgpu_parity(nat, rank, world, dev)
        # 126 keys per read; an owner receives about one rank's worth of keys (+ 30 % head room); one buffer more than
        # steps in flight (multigpu.P2PExchange)
        p2p = multigpu.P2PExchange(nat, dist, rank, world, dev, int(READS_PER_RANK * 126 * 1.3) + (1 << 20), nbuf=inflight + 1)

    def barrier():
        torch.cuda.synchronize(dev)
        nat.device_sync(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fq = make_reads(rank, READS_PER_RANK, world)
    nbytes = fq.nbytes
    bases = READS_PER_RANK * READ_LEN
    d_in = torch.from_numpy(fq).to("cuda:%d" % dev)
    with near_gpu(dev):
        # ZB_PINNED_WC=1: write-combined pinned input (the CPU only writes it; an experiment for hosts where 8 ranks copying
        # at once are bound by the host side of PCIe)
        pinned_in = nat.PinnedArray(nbytes, np.uint8, write_combined=os.environ.get("ZB_PINNED_WC", "0") == "1")
        pinned_in.a[:] = fq
    h_in = pinned_in.a

    # ---------------- device-resident: warm-up, then K timed steps
    res = None
    for _ in range(max(1, args.warmup)):
        res = step_device(nat, dev, d_in.data_ptr(), nbytes, p2p)
    n_full, n_trim, st0, wsz, twsz = res
    if p2p is not None:
        p2p.route_ms.clear(); p2p.remote_bytes.clear()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()          # NVML init happens here, outside the timed region
    barrier()
    nat.dbg_profile(True, dev)
    launches0 = nat.launch_count(dev)
    nat.timer_start(dev)
    d_times = []
    for _ in range(args.steps):
        t_it = time.perf_counter()
        step_device(nat, dev, d_in.data_ptr(), nbytes, p2p)
        d_times.append((time.perf_counter() - t_it) * 1e3)
    ms_dev = nat.timer_stop(dev)
    print("rank %d device-resident per-step wall ms: %s" % (rank, [round(x, 1) for x in d_times]), file=sys.stderr)
    barrier()
    launches = nat.launch_count(dev) - launches0
    prof = nat.dbg_profile(False, dev)
    route_ms_dev = list(p2p.route_ms) if p2p is not None else []
    remote_bytes_dev = list(p2p.remote_bytes) if p2p is not None else []

    # ---------------- end to end through host buffers (pinned in, pinned out)
    # `inflight` steps are kept in flight by as many host threads that simply call the API (the library gives every host
    # thread its own stream and allocator and hands the copy-in / copy-out engines from thread to thread), so the H2D
    # copy of one step overlaps the kernels / D2H of the others -- what a user with more than one input file does.  Every
    # step still copies its own 315 MB in and its own encoded result out inside the timed region.  N > 1: the exchange of
    # step s is a collective; P2PExchange issues them in step order whatever thread a step runs on.
    def make_out():
        with near_gpu(dev):
            bufs = [nat.PinnedArray(max(n, 1) + 64, np.uint64) for n in (twsz[0], twsz[1])]
            for b in bufs:
                b.a[:] = 0
        return bufs

    trace = [] if os.environ.get("ZB_E2E_TRACE") else None

    def feed_plain(km):
        km.feed(h_in, False)

    def feed_resident(km):
        km.feed_dev(d_in.data_ptr(), nbytes, False)

    def step_e2e(seq, out, feed=feed_plain):
        torch.cuda.set_device(dev)
        tr = [threading.get_ident() % 1000, time.perf_counter()] if trace is not None else None
        km = nat.Kmerizer(K, dev)
        if p2p is not None:
            p2p.prepare(km)
        feed(km)
        if tr: tr.append(time.perf_counter())
        if p2p is not None:
            p2p.exchange(km, seq=seq, consume=False)
        if tr: tr.append(time.perf_counter())
        r = finish_step(km, fetch=[b.a for b in out] if out is not None else None)
        if tr:
            tr.append(time.perf_counter())
            trace.append(tr)
        return r

    def run_e2e(feed, fetch=True):
        """`args.steps` timed steps (after warm-up) kept `inflight` at a time -> (ms per step, stage profile, result of the
        last step of one thread, its output buffers).  fetch=False: nothing leaves the device (the device-resident region);
        the time is then taken with CUDA events recorded while the device is idle on both sides of the region."""
        warm = min(args.warmup, 2)
        total_steps = inflight * warm + args.steps
        base_seq = p2p.step if p2p is not None else 0
        gate = threading.Barrier(inflight + 1)     # everybody is warm
        go = threading.Barrier(inflight + 1)       # the per-stage profile has been reset: the timed region starts
        results = [None] * inflight
        errors = []

        def worker(i):
            try:
                out = make_out() if fetch else None
                # step numbers: thread i runs i, i + inflight, ... -- the same assignment on every rank
                seqs = list(range(i, total_steps, inflight))
                for s_ in seqs[:warm]:
                    step_e2e(base_seq + s_, out, feed)
                gate.wait()
                go.wait()
                r = None
                for s_ in seqs[warm:]:
                    r = step_e2e(base_seq + s_, out, feed)
                results[i] = (r, out)
            except BaseException as e:   # pragma: no cover
                errors.append(e)
                for b_ in (gate, go):
                    try:
                        b_.abort()
                    except Exception:
                        pass

        nat.dbg_profile(True, dev)
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(inflight)]
        for t_ in ths:
            t_.start()
        e0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            gate.wait()
            barrier()                    # all ranks are warm
            nat.dbg_profile(True, dev)   # drop the warm-up stages (no worker touches the library between the two barriers)
            ev0.record()                 # the device is idle: the event completes at once
            go.wait()
            e0 = time.perf_counter()
        except threading.BrokenBarrierError:
            pass
        for t_ in ths:
            t_.join()
        if errors:
            raise errors[0]
        nat.device_sync(dev)             # every stream of the library has drained
        ev1.record()
        barrier()
        ms = (time.perf_counter() - e0) * 1e3 / args.steps
        if not fetch:
            ms = ev0.elapsed_time(ev1) / args.steps
        pr = nat.dbg_profile(False, dev)
        outs = [r for r in results if r is not None]
        for (_, o) in outs[1:]:
            for b_ in (o or []):
                b_.free()
        return (ms, pr) + outs[0]

    # ---------------- device-resident, `inflight` steps kept in flight (one host thread and one stream each): the host
    # round trips of one step (a dozen few-byte read-backs) and, at N > 1, its exchange over NVLink run under the kernels
    # of the others.  This is the region `value` is quoted on when it is the faster one; the single-stream region above
    # gives the per-stage times and the roofline of the kernels running alone.
    (pipe_ms, pipe_prof, pipe_res, _) = run_e2e(feed_resident, fetch=False)
    pipe_ok = (pipe_res[0], pipe_res[1], pipe_res[3], pipe_res[4]) == (n_full, n_trim, wsz, twsz) and pipe_res[2]["hist"] == st0["hist"]
    if rank == 0:
        print("device-resident, %d in flight: %.2f ms/step (%s)" % (inflight, pipe_ms, "ok" if pipe_ok else "MISMATCH"), file=sys.stderr)

    (e2e_ms, e_prof, e_res, e_out) = run_e2e(feed_plain)
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions (device-resident and e2e)
    # what the timed e2e steps fetched must be the streams a fresh device-resident step produces (sizes, histogram, words)
    e2e_ok = (e_res[0], e_res[1], e_res[3], e_res[4]) == (n_full, n_trim, wsz, twsz) and e_res[2]["hist"] == st0["hist"]
    ref_k, ref_c = np.zeros(twsz[0] + 64, np.uint64), np.zeros(twsz[1] + 64, np.uint64)
    km = nat.Kmerizer(K, dev)
    if p2p is not None:
        p2p.prepare(km)
    km.feed_dev(d_in.data_ptr(), nbytes, False)
    if p2p is not None:
        p2p.exchange(km, consume=False)
    finish_step(km, fetch=[ref_k, ref_c])
    e2e_ok = e2e_ok and np.array_equal(ref_k[:twsz[0]], e_out[0].a[:twsz[0]]) and np.array_equal(ref_c[:twsz[1]], e_out[1].a[:twsz[1]])
    if rank == 0:
        print("e2e (%d in flight) %.2f ms/step; stage ms/step: %s" % (
            inflight, e2e_ms, {k: round(v[0] / args.steps, 3) for k, v in e_prof.items()}), file=sys.stderr)
        if trace:
            t00 = min(r[1] for r in trace)
            for r in sorted(trace, key=lambda r: r[1]):
                print("trace thread %3d: start %.2f | fed %.2f | exchanged %.2f | counted, encoded, fetched %.2f" % (
                    (r[0],) + tuple((x - t00) * 1e3 for x in r[1:])), file=sys.stderr)

    # ---------------- reduce over ranks (max time), aggregate throughput
    per_step_ms = ms_dev / args.steps   # CUDA events on the library stream around the K steps
    if world > 1:
        tt = torch.tensor([per_step_ms, e2e_ms, pipe_ms, 0.0 if pipe_ok else 1.0], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per_step_ms, e2e_ms, pipe_ms, pipe_ok = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3]) == 0.0
        tl = torch.tensor([float(launches), 1.0 if e2e_ok else 0.0], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])
        e2e_ok = int(tl[1]) == world
    # ---------------- the same end-to-end step fed from BLOCK-COMPRESSED input (what `zot kmerize reads.fq.gz` does for a
    # bgzip'd file): the compressed bytes cross PCIe, every gzip member is inflated by one warp on the device
    # (csrc/inflate.cu), the text never exists on the host.  Informational: the headline `e2e` stays the plain-text one.
    bgzf = None
    if not args.no_bgzf:
        from concurrent.futures import ThreadPoolExecutor
        from tools import synth
        t0 = time.perf_counter()
        raw = fq.tobytes()
        chunk = 65280 * 64
        with ThreadPoolExecutor(min(16, os.cpu_count() or 4)) as ex:
            parts = list(ex.map(lambda o: synth.bgzf_bytes(raw[o:o + chunk], level=6, eof=False), range(0, len(raw), chunk)))
        z = b"".join(parts) + synth.bgzf_bytes(b"")
        del raw, parts
        t_comp = time.perf_counter() - t0
        pin_z = nat.PinnedArray(len(z), np.uint8)
        pin_z.a[:] = np.frombuffer(z, dtype=np.uint8)

        def feed_bgzf(km):
            st, _ = nat.stage_bgzf(pin_z.a, dev)
            km.feed_staged(st, False)

        (g_ms, g_prof, g_res, g_out) = run_e2e(feed_bgzf)
        g_ok = (g_res[0], g_res[1], g_res[3], g_res[4]) == (n_full, n_trim, wsz, twsz) and g_res[2]["hist"] == st0["hist"] and \
            np.array_equal(g_out[0].a[:twsz[0]], e_out[0].a[:twsz[0]]) and np.array_equal(g_out[1].a[:twsz[1]], e_out[1].a[:twsz[1]])
        zbytes = len(z)
        if world > 1:
            tg = torch.tensor([g_ms, 0.0 if g_ok else 1.0, float(zbytes)], dtype=torch.float64, device="cuda:%d" % dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            g_ms, g_ok, zbytes = float(tg[0]), float(tg[1]) == 0.0, int(tg[2])
        bgzf = {"value": bases * world / (g_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": g_ms, "steps_in_flight": inflight,
                "h2d_bytes_per_step": int(zbytes), "d2h_bytes_per_step": int(8 * sum(twsz)),
                "inflate_ms_per_step": round(g_prof.get("inflate", (0.0, 0))[0] / args.steps, 3),
                "result_check": "ok" if g_ok else "MISMATCH",
                "input": "the rank's FASTQ as BGZF (bgzip framing, zlib level 6: %d -> %d bytes, %.1f s to compress, not timed) in "
                         "pinned host memory; inflated on the device, one warp per member" % (nbytes, len(z), t_comp)}
        if rank == 0:
            print("e2e from BGZF (%d in flight) %.2f ms/step; stage ms/step: %s" % (
                inflight, g_ms, {k: round(v[0] / args.steps, 3) for k, v in g_prof.items()}), file=sys.stderr)
        for b_ in g_out:
            b_.free()
        pin_z.free()
        del z
    keys_per_step = stage_keys(nat, dev, d_in, nbytes)
    if p2p is not None:
        p2p.close()
        p2p = None
    for b in e_out:
        b.free()
    del d_in
    pinned_in.free()
    nat.release_cache(dev)
    torch.cuda.empty_cache()

    human = None
    if not args.no_human:
        human = bench_human(nat, dev, rank, world)
        nat.release_cache(dev)
        torch.cuda.empty_cache()
    pairs = None
    if not args.no_pairs:
        pairs = bench_pairs(nat, dev, rank, world, max(1, min(args.steps, 2)))
        nat.release_cache(dev)
        torch.cuda.empty_cache()
    cli = None
    if not args.no_cli and world == 1:
        cli = bench_cli(nat, dev)
    if rank != 0:
        return
    total_bases = bases * world
    one_ms = per_step_ms                      # one step in flight: the region the per-stage times and the roofline come from
    in_flight_dev = 1
    if pipe_ok and pipe_ms < per_step_ms:
        per_step_ms, in_flight_dev = pipe_ms, inflight
    value = total_bases / (per_step_ms * 1e-3) / 1e9
    e2e = total_bases / (e2e_ms * 1e-3) / 1e9

    peak, peak_kind = load_peaks()
    # dominant kernel: one onesweep radix pass over the canonical keys of this rank
    sp = prof.get("sort_pass_keys", (0.0, 0))
    pass_ms = sp[0] / sp[1] if sp[1] else None
    roofline = None
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if pass_ms and keys_per_step:
        achieved = BYTES_PER_KEY_PASS * keys_per_step / (pass_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "onesweep_kernel (one LSD radix pass, 16 B/key)", "achieved": achieved,
                    "peak": peak, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)", "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "launch_ms": pass_ms, "keys_per_launch": keys_per_step,
                    "pipeline_achieved": BYTES_PER_BASE * bases / (per_step_ms * 1e-3) / 1e9,
                    "pipeline_frac": BYTES_PER_BASE * bases / (per_step_ms * 1e-3) / 1e9 / peak,
                    "pipeline_bytes_per_base": BYTES_PER_BASE,
                    "pipeline_note": "SURVEY.md 8d's fixed numerator (7 full radix passes over both strands); this design sorts "
                                     "canonical keys with 2 top-bit passes, so the bytes that really cross HBM are in `kernels`",
                    "stage_ms_per_step": stage_ms,
                    "kernels": kernel_table(prof, args.steps, one_ms, peak, nbytes, bases, keys_per_step, n_full, n_trim, wsz, twsz)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_ms, "steps_in_flight": in_flight_dev,
        "one_in_flight": {"value": total_bases / (one_ms * 1e-3) / 1e9, "ms_per_step": one_ms,
                          "note": "the same K steps one after the other on one stream (CUDA events on that stream): the region "
                                  "`roofline` and its per-stage times are measured in"},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(8 * sum(twsz)),
                "ms_per_step": e2e_ms, "steps_in_flight": inflight,
                "result": "the two codec64 word streams of the trimmed set (what `zot trim` writes) in pinned host memory + count "
                          "histogram / acgt of the full set; the full set's two streams are encoded too and stay in HBM",
                "result_check": "ok" if e2e_ok else "MISMATCH",
                "pinned_near_gpu_cpus": len(gpu_cpus(dev)) if gpu_cpus(dev) else None},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "result": {"distinct_kmers": int(n_full), "after_trim": int(n_trim), "kmer_words": int(wsz[0]), "count_words": int(wsz[1]),
                   "trimmed_kmer_words": int(twsz[0]), "trimmed_count_words": int(twsz[1])},
    }
    if bgzf is not None:
        line["e2e_bgzf"] = bgzf
    if parity is not None:
        line["mgpu_parity"] = parity
    if route_ms_dev and "route_p2p" in prof:
        r_ms = prof["route_p2p"][0] / max(1, prof["route_p2p"][1])
        r_b = float(np.mean(remote_bytes_dev)) if remote_bytes_dev else 0.0
        reserve = os.environ.get("ZB_P2P_RESERVE", "0") == "1"
        line["nvlink"] = {"exchange": "fused: route_p2p_kernel stores every key into its owner's buffer over NVLink peer memory (CUDA IPC); "
                                      + ("a thread block reserves its run there with one system-scope atomic on the owner's cursor word; "
                                         "NCCL only for the closing all-reduce" if reserve else "NCCL only for the count matrix and the closing all-reduce"),
                          "route_kernel_ms": r_ms, "remote_bytes_per_gpu": r_b, "GBps_per_gpu_out": r_b / r_ms / 1e6 if r_ms else None,
                          "note": "the kernel also moves this rank's own share locally, so the NVLink rate is a lower bound",
                          "peak_GBps_per_direction": 900.0, "measured_peer_copy_GBps": 770.0}
    if human is not None:
        line["human"] = human
    if pairs is not None:
        if world == 1 and not args.no_cpu_baseline:
            pairs["cpu_baseline"] = cpu_baseline_pairs(pairs)
        line["pairs"] = pairs
    if cli is not None:
        line["cli"] = cli
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def kernel_table(prof, steps, step_ms, peak, text_bytes, bases, nkeys, n_full, n_trim, wsz, twsz):
    """every stage that takes >= 5 % of the device-resident step: ms per step, the bytes its algorithm has to move, GB/s and
    the fraction of the measured HBM peak.  T = text bytes, C = base codes (one per base + one break per read), N = canonical
    keys, D = distinct canonical k-mers, S = both-strand set, R = set after trim."""
    T, C, N = float(text_bytes), float(bases + bases // READ_LEN), float(nkeys)
    S, R = float(n_full), float(n_trim)
    D = S / 2.0
    npass = max(1, prof.get("sort_pass_keys", (0, 1))[1] // max(1, steps))
    model = {
        "parse": (2 * T + C, "fq_count + fq_scan + fastq_kernel: text read twice, codes written"),
        "extract": (C + 8 * N, "extract_kernel: codes read, canonical keys written"),
        "sort_hist": (8 * N, "sort_hist_kernel: keys read once for all digit histograms"),
        "sort_pass_keys": (16 * N * npass, "onesweep_kernel x %d: 8 B read + 8 B written per key and pass" % npass),
        "segcount": (8 * N + 2 * 12 * D + 12 * D, "bc_bounds + bucket_count + scan + compact: keys read, distinct run staged, re-read, written"),
        "mirror": (12 * D + 12 * D, "mirror_kernel: canonical run read, reverse complements written"),
        "sort_pass_pairs": (2 * 24 * D, "onesweep_kernel<pairs> x 2 over the mirrored half (top bits only)"),
        "mirror_buckets": (2 * 12 * D + 12 * S, "bc_bounds x 2 + mirror_merge_kernel: both halves read, both-strand set written"),
        "stats": (12 * S, "stats_kernel: set read once (histogram bins + first occurrence, acgt)"),
        "encode": (12 * S * 2 + 8 * (wsz[0] + wsz[1]) * 2 + 12 * R * 2 + 8 * (twsz[0] + twsz[1]) * 2,
                   "enc_tile + enc_emit x 4 streams: values read twice, words written (+ one copy into exact-size buffers)"),
        "trim": (12 * S + 12 * R, "compact_kernel<TrimOp>: set read, kept entries written"),
        "route_p2p": (16 * N, "route_p2p_kernel: keys read, stored into the owners' buffers (local HBM or NVLink)"),
    }
    nested = {"sort", "mirror_merge"}     # containers of other stages
    rows = []
    for name, (ms_tot, calls) in prof.items():
        ms = ms_tot / steps
        if name in nested or ms < 0.05 * step_ms:
            continue
        by, what = model.get(name, (None, ""))
        row = {"stage": name, "ms_per_step": round(ms, 4), "share_of_step": round(ms / step_ms, 3), "launch_groups_per_step": calls // max(1, steps),
               "what": what}
        if by:
            row["algorithmic_bytes"] = int(by)
            row["GBps"] = round(by / (ms * 1e-3) / 1e9, 1)
            row["frac_of_hbm_peak"] = round(by / (ms * 1e-3) / 1e9 / peak, 3)
        rows.append(row)
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


# ------------------------------------------------------------------------------------------------
HUMAN_K = 31
HUMAN_GENOME_PER_RANK = int(os.environ.get("ZB_HUMAN_GENOME", 375000000))
HUMAN_READS_PER_RANK = int(os.environ.get("ZB_HUMAN_READS", 75000000))
HUMAN_BATCH = int(os.environ.get("ZB_HUMAN_BATCH", 3000000))
HUMAN_ERR = 0.001


def bench_human(nat, dev, rank, world):
    """BASELINE.json configs[4]: human-scale synthetic genome at 30x 150 bp reads, kmerize+count k=31 with the hash-range
    exchange; weak scaling as the config defines it -- 375 Mbp of genome and 75 M reads per GPU, so N = 8 is the 3 Gbp /
    600 M reads / 90 Gbases instance.  A 200 GB FASTQ cannot be staged on the box: the reads are generated
    on the device batch by batch as base codes (torch: generation is plumbing and is NOT timed) and fed through
    zb_kmerize_feed_codes_dev; timed per batch: extraction, route_p2p exchange, sort + count of what the previous exchange
    delivered, fold into the rank's running counted set; then finish (last count, mirror, merge), stats and trim -c 2.
    Checked through invariants (the oracle cannot run 11 Gbases): sum of counts = 2 x windows, every share strictly
    ascending, sum(hist c x freq) = sum of counts, record count."""
    import torch
    from zotmer_b200 import multigpu
    dist = None
    if world > 1:
        import torch.distributed as dist
    dv = "cuda:%d" % dev
    L, Kh = READ_LEN, HUMAN_K
    G = HUMAN_GENOME_PER_RANK * world
    nreads = HUMAN_READS_PER_RANK
    B = min(HUMAN_BATCH, nreads)
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
        p2p = multigpu.P2PExchange(nat, dist, rank, world, dev, int(B * (L - Kh + 1) * 1.15) + (1 << 20))

    def barrier():
        torch.cuda.synchronize(dev)
        nat.device_sync(dev)
        if world > 1:
            dist.barrier()

    rgen = torch.Generator(device=dv)
    ar = torch.arange(L, device=dv, dtype=torch.int64)

    def make_batch(b, seed):
        rgen.manual_seed(seed)
        pos = torch.randint(0, G - L, (b,), device=dv, generator=rgen)
        codes = torch.empty((b, L + 1), dtype=torch.uint8, device=dv)
        rd = genome[(pos[:, None] + ar[None, :]).reshape(-1)].reshape(b, L)
        rev = torch.rand(b, device=dv, generator=rgen) < 0.5
        rd = torch.where(rev[:, None], 3 - rd.flip(1), rd)
        e = torch.rand((b, L), device=dv, generator=rgen) < HUMAN_ERR
        sub = torch.randint(1, 4, (b, L), dtype=torch.uint8, device=dv, generator=rgen)
        rd = torch.where(e, (rd + sub) & 3, rd)
        codes[:, :L] = rd
        codes[:, L] = 4
        return codes

    # warm-up: two batches through a throw-away kmerizer (device allocator, peer mappings, kernel attributes)
    kw = nat.Kmerizer(Kh, dev)
    if p2p is not None:
        p2p.prepare(kw)
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

    nat.dbg_profile(True, dev)
    km = nat.Kmerizer(Kh, dev)
    if p2p is not None:
        p2p.prepare(km)
    timed, fed, windows, nb = 0.0, 0, 0, 0
    while fed < nreads:
        b = min(B, nreads - fed)
        codes = make_batch(b, 1000003 * (rank + 1) + nb)          # NOT timed
        barrier()
        t0 = time.perf_counter()
        km.feed_codes_dev(codes.data_ptr(), codes.numel(), b)
        if p2p is not None:
            p2p.exchange(km)
        nat.device_sync(dev)
        timed += time.perf_counter() - t0
        del codes
        fed += b
        windows += b * (L - Kh + 1)
        nb += 1
    barrier()
    del genome
    torch.cuda.empty_cache()                   # the generator's scratch goes back to the driver before the big allocations
    t_feed = timed
    t0 = time.perf_counter()
    s, nr = km.finish()
    km.close()
    st = s.stats()
    t = s.trim(2)
    nat.device_sync(dev)
    timed += time.perf_counter() - t0
    t_finish = timed - t_feed
    prof = nat.dbg_profile(False, dev)
    barrier()
    # ---- invariants
    kp, cp = s.dev_ptrs()
    n = len(s)
    ok_sorted = True
    step = 1 << 27
    kt = multigpu._as_tensor(kp, n, torch.int64, dev)
    for a in range(0, n - 1, step):            # keys < 2^62: signed comparison is fine
        bnd = min(n - 1, a + step)
        ok_sorted = ok_sorted and bool((kt[a + 1:bnd + 1] > kt[a:bnd]).all())
    hist_total = sum(int(c) * int(f) for c, f in st["hist"])
    good = ok_sorted and hist_total == st["total"]
    vals = torch.tensor([float(st["total"]), float(windows), float(n), float(len(t)), float(nr)], dtype=torch.float64, device=dv)
    tmax = torch.tensor([timed, t_feed], dtype=torch.float64, device=dv)
    flags = torch.tensor([1.0 if good else 0.0], dtype=torch.float64, device=dv)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    total, win, ndist, ntrim, nrec = [float(x) for x in vals.tolist()]
    checks = {"sum_of_counts_is_2x_windows": total == 2.0 * win, "shares_strictly_ascending_and_hist_adds_up": flags.item() == 1.0,
              "records": nrec == float(nreads) * world}
    s.free()
    t.free()
    if p2p is not None:
        p2p.close()
    if not all(checks.values()):
        raise SystemExit("bench.py: human-scale invariants FAILED: %r" % checks)
    peak, _ = load_peaks()
    secs = float(tmax[0])
    bases = float(nreads) * L * world
    val = bases / secs / 1e9
    return {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "scaling": "weak",
            "config": {"workload": "config[4]: %.3f Gbp synthetic genome (5 %% repeats), %d reads x %d bp (%.2f Gbases, 30x), error %.1f %%, "
                                   "kmerize+count k=%d, stats, trim min-count 2; %d GPU(s), %d reads and %d Mbp per GPU" % (
                                       G / 1e9, nreads * world, L, bases / 1e9, 100 * HUMAN_ERR, Kh, world, nreads, HUMAN_GENOME_PER_RANK // 1000000),
                       "k": Kh, "batches_per_gpu": nb, "batch_reads": B,
                       "data": "reads generated on the device as base codes (not timed); inputs resident in HBM"},
            "seconds": secs, "seconds_batches": float(tmax[1]), "seconds_finish_stats_trim_rank0": t_finish,
            "distinct_kmers_both_strands": int(ndist), "after_trim": int(ntrim), "sum_of_counts": int(total),
            "roofline": {"bound": "hbm", "bytes_per_base": HUMAN_BYTES_PER_BASE, "achieved": HUMAN_BYTES_PER_BASE * bases / world / secs / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": HUMAN_BYTES_PER_BASE * bases / world / secs / 1e9 / peak,
                         "note": "SURVEY.md 8d's fixed numerator for k=31 (8 full passes over both strands), per GPU"},
            "stage_ms_rank0": {k: round(v[0], 1) for k, v in prof.items()},
            "checks": checks}


# ------------------------------------------------------------------------------------------------
PAIR_SETS = int(os.environ.get("ZB_BENCH_PAIR_SETS", 1000))
PAIR_CLADES = 10
PAIR_KEYS = int(os.environ.get("ZB_BENCH_PAIR_KEYS", 9950000))


def bench_pairs(nat, dev, rank, world, steps):
    """BASELINE.json's second metric, pairwise Jaccard set-pairs/s, on configs[3] at FULL size: all 499,500 pairs of 1,000
    synthetic bacterial k-mer sets (k=25: 50-bit keys, ~10 M k-mers each = 80 GB), through zb_allpairs_abc -- the call behind
    `zot dist` / `zot jaccard -a`.  The sets are made on the device (1,000 x 80 MB cannot come through PCIe in a bounded
    run; not timed): 10 clade bases of distinct random keys; a member keeps a base key with probability 1 - q (q = share
    of k-mers hit by a substitution, 2.5 % .. 22 % for 0.1 % .. 1 % divergence at k=25) and draws fresh keys for the rest.
    value: kernel time of the call (CUDA events); e2e: the call incl. the D2H of the (a, b, c) matrix, the all-reduce over
    ranks and the Jaccard values on the host.  N > 1: every rank holds all sets and computes its share of the work units
    (tile x key-range shard); one all-reduce adds the partial matrices up.
    Checked: a + b = |X_i| and a + c = |X_j| for EVERY pair; 24 sampled pairs against the pair-at-a-time merge-path
    kernel (zb_pairs_abc); inside a clade |X n Y| within 1 % of (1 - qi)(1 - qj)|base|."""
    import torch
    from zotmer_b200 import multigpu
    dv = torch.device("cuda:%d" % dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
    g = torch.Generator(device=dv)
    g.manual_seed(1000)
    t0 = time.time()
    clades = min(PAIR_CLADES, PAIR_SETS)
    basesets = [torch.unique(torch.randint(0, 1 << 50, (PAIR_KEYS,), generator=g, device=dv, dtype=torch.int64)) for _ in range(clades)]
    sets, q_of, clade_of = [], [], []
    for i in range(PAIR_SETS):
        c = i % clades
        q = 0.025 + 0.195 * ((i // clades) / max(1, PAIR_SETS // clades - 1)) if PAIR_SETS > clades else 0.05
        b = basesets[c]
        drop = torch.rand(b.numel(), generator=g, device=dv) < q
        fresh = torch.randint(0, 1 << 50, (int(drop.sum().item()),), generator=g, device=dv, dtype=torch.int64)
        keys = torch.unique(torch.cat([b[~drop], fresh]))
        sets.append(nat.KmerSet.from_device(keys.data_ptr(), None, keys.numel(), device=dev))
        q_of.append(q)
        clade_of.append(c)
        del keys, fresh, drop
    base_sizes = np.array([b.numel() for b in basesets], np.float64)
    del basesets
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    build_s = time.time() - t0
    sizes = np.array([len(s) for s in sets], np.uint64)
    npairs = PAIR_SETS * (PAIR_SETS - 1) // 2
    b, e, stp = multigpu.unit_share(PAIR_SETS, rank, world)

    def step():
        part = nat.allpairs_abc(sets, b, e, stp)
        if world > 1:
            t = torch.from_numpy(part.view(np.int64)).to(dv)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            part = t.cpu().numpy().view(np.uint64)
        jac = part[:, 0].astype(np.float64) / np.maximum(part.sum(axis=1), 1).astype(np.float64)
        return part, jac

    step()
    torch.cuda.synchronize(dev)
    nat.device_sync(dev)
    if world > 1:
        dist.barrier()
    nat.dbg_profile(True, dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        abc, jac = step()
    nat.device_sync(dev)
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    prof = nat.dbg_profile(False, dev)
    kern_ms = (prof.get("allpairs", (0.0, 1))[0] + prof.get("allpairs_offsets", (0.0, 1))[0]) / steps   # both kernels of the call
    if world > 1:
        tt = torch.tensor([kern_ms, wall_ms], dtype=torch.float64, device=dv)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kern_ms, wall_ms = float(tt[0]), float(tt[1])
    # ---- checks (rank 0 holds the full matrix after the all-reduce; so does every rank)
    I, J = np.triu_indices(PAIR_SETS, 1)
    adds_up = bool(np.array_equal(abc[:, 0] + abc[:, 1], sizes[I]) and np.array_equal(abc[:, 0] + abc[:, 2], sizes[J]))
    rng = np.random.default_rng(5)
    pick = rng.choice(npairs, size=min(24, npairs), replace=False)
    cl, qq = np.array(clade_of), np.array(q_of)
    same = cl[I] == cl[J]
    if same.any():
        pick[: min(8, len(pick))] = np.flatnonzero(same)[rng.choice(int(same.sum()), size=min(8, len(pick)), replace=False)]
    ref = nat.pairs_abc(sets, I[pick], J[pick])
    sampled = bool(np.array_equal(ref, abc[pick]))
    clade_ok, rel_max = True, 0.0
    if same.any() and PAIR_KEYS >= 1000000:
        expect = (1 - qq[I]) * (1 - qq[J]) * base_sizes[cl[I]]
        rel = np.abs(abc[same, 0].astype(np.float64) - expect[same]) / expect[same]
        rel_max = float(rel.max())
        clade_ok = rel_max < 0.01
    for s in sets:
        s.free()
    checks = {"every_pair_adds_up": adds_up, "sampled_pairs_equal_pair_at_a_time_kernel": sampled, "clade_structure": clade_ok}
    if not all(checks.values()):
        raise SystemExit("bench.py: all-pairs checks FAILED: %r" % checks)
    peak, _ = load_peaks()
    nblk = -(-PAIR_SETS // multigpu.AP_S)
    pair_bytes = 8.0 * float(sizes.sum()) * (PAIR_SETS - 1)      # sum over pairs of 8 (|X| + |Y|)
    moved_bytes = 8.0 * float(sizes.sum()) * (1 + nblk)         # offsets pass + one gather per tile a set belongs to
    return {"metric": "pairwise Jaccard set-pairs/s", "value": npairs / (kern_ms * 1e-3), "unit": "set-pairs/s",
            "e2e": {"value": npairs / (wall_ms * 1e-3), "unit": "set-pairs/s", "d2h_bytes_per_step": int(npairs * 24)},
            "config": {"workload": "config[3]: all pairs of %d synthetic bacterial k-mer sets (k=25, ~%d k-mers each, %d clades, %.1f GB of "
                                   "keys resident in HBM; built on the device in %.1f s, not timed)" % (
                                       PAIR_SETS, int(sizes.mean()), clades, sizes.sum() * 8 / 1e9, build_s),
                       "pairs": npairs, "steps": steps,
                       "parallelism": "1 GPU" if world == 1 else "%d GPUs: work units (pairs of 32-set blocks x 8 key-range shards) sharded, one all-reduce" % world},
            "ms_per_step": kern_ms, "n_gpus": world,
            "stage_ms_rank0": {k: round(v[0] / steps, 2) for k, v in prof.items()},
            "roofline": {"bound": "hbm", "achieved": moved_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": moved_bytes / (kern_ms * 1e-3) / 1e9 / (peak * world),
                         "note": "numerator = the bytes this design has to move: every k-mer (8 B) once for the bucket offsets and "
                                 "once per tile its set belongs to (N/32 block pairs); the kernel is bound by shared-memory work "
                                 "(hash dedupe, bit-matrix transpose), not by HBM",
                         "pair_at_a_time_model_GBps": pair_bytes / (kern_ms * 1e-3) / 1e9,
                         "pair_at_a_time_note": "8 (|X| + |Y|) B per pair (SURVEY.md 8d) over the same time: what a merge per pair would "
                                                "have to sustain to keep up"},
            "check": dict(checks, jaccard_first_pair=float(jac[0]), max_jaccard=float(jac.max()), clade_rel_err_max=rel_max)}


# ------------------------------------------------------------------------------------------------
def bench_cli(nat, dev):
    """the drop-in command itself, as a user runs it: `zot kmerize 25 out.k25 reads.fq` on configs[1]'s FASTQ (315 MB file in
    the page cache -> k-mer set file), wall clock of cli.main() in a warm process (library loaded, CUDA context up)."""
    import io
    import contextlib
    import shutil
    import tempfile
    from zotmer_b200 import cli
    tmp = tempfile.mkdtemp(prefix="zb_cli_", dir=os.environ.get("ZB_TMP"))
    try:
        fq = os.path.join(tmp, "reads_5M_30x.fq")
        make_reads(0, READS_PER_RANK).tofile(fq)
        times = []
        for it in range(5):
            out = os.path.join(tmp, "r%d.k25" % it)     # a new file every time, as a user's run writes one
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                cli.main(["kmerize", str(K), out, fq])
            nat.device_sync(dev)
            times.append((time.perf_counter() - t0) * 1e3)
            if it < 4:
                os.remove(out)
        nat.dbg_profile(True, dev)
        with contextlib.redirect_stdout(io.StringIO()):
            cli.main(["kmerize", str(K), os.path.join(tmp, "p.k25"), fq])
        stages = {k: round(v[0], 2) for k, v in nat.dbg_profile(False, dev).items()}
        ms = float(np.median(times[2:]))
        return {"command": "zot kmerize %d r.k25 reads_5M_30x.fq" % K, "wall_ms": ms, "wall_ms_all": [round(x, 1) for x in times],
                "value": READS_PER_RANK * READ_LEN / ms / 1e6, "unit": UNIT, "input_bytes": os.path.getsize(fq),
                "output_bytes": os.path.getsize(out), "tmpdir": os.path.dirname(tmp),
                "stage_ms": stages,
                "note": "files in -> file out through zotmer_b200/cli.py (docopt grammar, staged H2D through the pinned ring, kernels, "
                        "codec64 on the device, D2H + pwrite by the I/O threads, casket table); median of the last 3 of 5 runs"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def stage_keys(nat, dev, d_in, nbytes):
    """number of canonical keys one step sorts (valid windows of the rank's reads)"""
    km = nat.Kmerizer(K, dev)
    km.feed_dev(d_in.data_ptr(), nbytes, False)
    n = km.pending()
    km.close()
    return int(n)


def cpu_baseline():
    """oracle port of the reference (pure Python, 1 thread) on a bounded sample of the same reads"""
    from oracle import zot_oracle as zo
    sample_reads = int(os.environ.get("ZB_CPU_SAMPLE_READS", 30000))   # ~11 s of CPython on the GPU box
    fq = make_reads(0, READS_PER_RANK)[:sample_reads * 315].tobytes()
    t0 = time.perf_counter()
    reference_step(zo, fq)
    dt = time.perf_counter() - t0
    import shutil
    pypy = shutil.which("pypy") or shutil.which("pypy3")
    return {"value": sample_reads * READ_LEN / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "pypy": pypy or "not installed on this box and not installable offline (north star: PyPy only if it installs offline)",
            "sample": "first %d of the %d reads (%.1f Mbases, %.1f s): kmerize+count (hist, acgt), codec64 encode, trim, encode; CPython %s "
                      "single thread (the reference has no parallelism); host has %d cores" % (
                          sample_reads, READS_PER_RANK, sample_reads * READ_LEN / 1e6, dt, sys.version.split()[0],
                          os.cpu_count())}


def cpu_baseline_pairs(pairs):
    """the reference's two-pointer split() (library/dist.py:241-265, oracle port, pure Python, 1 thread) on a bounded
    sample: two sorted arrays of 300,000 k-mers with half of them shared; cost is linear in |X| + |Y|"""
    from oracle import zot_oracle as zo
    rng = np.random.default_rng(4)
    m = 300000
    pool = np.unique(rng.integers(0, 2 ** 50, 2 * m, dtype=np.uint64))
    xs = [int(v) for v in np.sort(rng.choice(pool, m, replace=False))]
    ys = [int(v) for v in np.sort(rng.choice(pool, m, replace=False))]
    t0 = time.perf_counter()
    zo.split(xs, ys)
    dt = time.perf_counter() - t0
    per_pair_elems = 2.0 * float(pairs["config"]["workload"].split("~")[1].split(" ")[0])
    return {"value": 1.0 / (dt * per_pair_elems / (2.0 * m)), "unit": "set-pairs/s", "cores": 1, "kind": "port",
            "sample": "split() on 2 x %d k-mers took %.2f s; scaled linearly to the %.0f k-mers of one pair of this "
                      "workload (one measure; the reference repeats split() per measure)" % (m, dt, per_pair_elems)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pairs", action="store_true", help="skip the set-pairs/s measurement (configs[3])")
    ap.add_argument("--no-human", action="store_true", help="skip the human-scale k=31 measurement (configs[4])")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the end-to-end step fed from block-compressed input")
    ap.add_argument("--no-cli", action="store_true", help="skip the command-level measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
