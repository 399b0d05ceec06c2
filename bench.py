#!/usr/bin/env python
"""
bench.py -- measures BASELINE.json's metric "kmerize+count Gbases/s" on the configuration it is
quoted on: synthetic 30x 150 bp Illumina-like FASTQ of a 5 Mbp genome (1,000,000 reads, 150 Mbases),
kmerize+count at k=25 followed by zot trim at min-count 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch (the whole read set of a rank):
  value : inputs resident in HBM (raw FASTQ text on the device) -> parse, extract, sort, count,
          mirror, trim, all through the C ABI (zb_kmerize_feed_dev ... zb_trim); device-timed.
  e2e   : the same through the host-buffer C ABI (zb_kmerize_feed from pinned host memory, results
          fetched back to pinned host memory), H2D and D2H inside the timed region.
N > 1 (weak scaling): every rank kmerizes its own 1M-read shard of the same genome, routes each
canonical k-mer to its owner rank (high bits of a 64-bit mix) with one NCCL all-to-all over NVLink,
and sorts/counts/trims its disjoint key range locally.
--impl reference: the reference's own algorithm (pure Python, single thread, as the reference is)
timed on a bounded sample of the same workload on the host.
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
            idx = self.gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except Exception:
                    idx = self.gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
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


def make_reads(rank, nreads, world=1):
    """rank's shard of the read set.  Weak scaling keeps the COVERAGE fixed as well as the reads per GPU: N ranks
    sample a genome of N x 5 Mbp (config[4]'s shape -- the genome grows with the machine), so every owner rank
    sees the same k-mer statistics (30x, ~10 M distinct genomic k-mers) as the single-GPU run."""
    from tools import synth
    g = synth.genome(GENOME * world, seed=17)
    return synth.fastq_array(g, nreads, L=READ_LEN, seed=18 + 1000 * rank).reshape(-1)


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference's own CPU path (pure Python, single thread): oracle port timed on a bounded sample."""
    if rank != 0:
        return
    from oracle import zot_oracle as zo
    sample_reads = int(os.environ.get("ZB_REF_SAMPLE_READS", 4000))
    fq = make_reads(0, READS_PER_RANK)[:sample_reads * 315].tobytes()
    bases = sample_reads * READ_LEN

    def step():
        xs, cs, h, acgt, nr = zo.kmerize_core(K, [("reads.fq", fq)])
        zo.words_to_bytes(zo.encode(zo.delta(xs)))       # the reference writes the set (54 % of its time)
        zo.words_to_bytes(zo.encode(cs))
        tx, tc = zo.trim_core(xs, cs, 2, None)
        zo.words_to_bytes(zo.encode(zo.delta(tx)))
        zo.words_to_bytes(zo.encode(tc))
        return len(xs)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
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
                        "then trim min-count 2" % (GENOME * n // 1000000, READS_PER_RANK),
            "k": K, "reads_per_gpu": READS_PER_RANK, "read_len": READ_LEN, "bases_per_gpu": READS_PER_RANK * READ_LEN,
            "genome_bp": GENOME * n,
            "parallelism": "1 GPU" if n == 1 else "%d GPUs: %d reads per GPU from a %d Mbp genome (30x), canonical k-mers "
                                                  "routed to their hash-range owner over NVLink" % (n, READS_PER_RANK, GENOME * n // 1000000),
            "l2": "inputs (315 MB of text, 1 GB of keys per GPU) are larger than the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
def step_device(nat, dev, d_ptr, nbytes, dist_ctx):
    """one step with inputs resident in HBM -> (trimmed set, full set)"""
    km = nat.Kmerizer(K, dev)
    km.feed_dev(d_ptr, nbytes, False)
    if dist_ctx is not None:
        exchange(nat, km, dist_ctx)
    s, nr = km.finish()
    km.close()
    t = s.trim(2)
    return s, t


def exchange(nat, km, ctx):
    """route every pending canonical k-mer to its owner rank (zotmer_b200/multigpu.py): fused routing + transfer
    over NVLink peer memory when the ranks could map each other's buffers, else bucket -> NCCL all-to-all"""
    from zotmer_b200 import multigpu
    if ctx.get("p2p") is not None:
        ctx["p2p"].exchange(km)
    else:
        multigpu.exchange_pending(nat, km, ctx)


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
    dist_ctx = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % dev))
        dist_ctx = {"world": world, "rank": rank, "dev": dev, "a2a_ms": [], "a2a_bytes": [],
                    "send": torch.empty(16, dtype=torch.int64, device="cuda:%d" % dev),
                    "recv": torch.empty(16, dtype=torch.int64, device="cuda:%d" % dev)}

    if dist_ctx is not None and os.environ.get("ZB_EXCHANGE", "p2p") == "p2p":
        from zotmer_b200 import multigpu
        import torch.distributed as dist
        ok = True
        try:
            # 126 keys per read; an owner receives about one rank's worth of keys (+ 30 % head room)
            dist_ctx["p2p"] = multigpu.P2PExchange(nat, dist, rank, world, dev, int(READS_PER_RANK * 126 * 1.3) + (1 << 20))
        except Exception as e:   # no peer mapping on this box: every rank must agree to fall back
            print("rank %d: P2P exchange unavailable (%r), using NCCL all-to-all" % (rank, e), file=sys.stderr)
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            dist_ctx["p2p"] = None

    def barrier():
        torch.cuda.synchronize(dev)
        nat.device_sync(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    fq = make_reads(rank, READS_PER_RANK, world)
    nbytes = fq.nbytes
    bases = READS_PER_RANK * READ_LEN
    d_in = torch.from_numpy(fq).to("cuda:%d" % dev)
    pinned_in = torch.from_numpy(fq).pin_memory()
    h_in = pinned_in.numpy()

    # ---------------- device-resident: warm-up, then K timed steps
    n_trim = n_full = 0
    for _ in range(args.warmup):
        s, t = step_device(nat, dev, d_in.data_ptr(), nbytes, dist_ctx)
        n_full, n_trim = len(s), len(t)
        s.free(); t.free()
    if dist_ctx is not None:
        dist_ctx["a2a_ms"].clear(); dist_ctx["a2a_bytes"].clear()
        if dist_ctx.get("p2p") is not None:
            dist_ctx["p2p"].route_ms.clear(); dist_ctx["p2p"].remote_bytes.clear()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()          # NVML init happens here, outside the timed region
    barrier()
    nat.dbg_profile(True, dev)
    launches0 = nat.launch_count(dev)
    nat.timer_start(dev)
    w0 = time.perf_counter()
    d_times = []
    for _ in range(args.steps):
        t_it = time.perf_counter()
        s, t = step_device(nat, dev, d_in.data_ptr(), nbytes, dist_ctx)
        s.free(); t.free()
        d_times.append((time.perf_counter() - t_it) * 1e3)
    ms_dev = nat.timer_stop(dev)
    print("rank %d device-resident per-step wall ms: %s" % (rank, [round(x, 1) for x in d_times]), file=sys.stderr)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    launches = nat.launch_count(dev) - launches0
    prof = nat.dbg_profile(False, dev)

    # ---------------- end to end through host buffers (pinned in, pinned out)
    # N = 1: four steps are kept in flight by four host threads (the library gives every host thread its own stream and
    # allocator and hands the copy-in / copy-out engines from thread to thread), so the H2D copy of one step overlaps the
    # kernels / D2H of the others -- what a user with more than one input file does.  Every step still copies its own
    # 315 MB in and its own result out inside the timed region.
    # N > 1: one step in flight (the exchange is a collective; all ranks must issue it in the same order).
    inflight = 1 if world > 1 else int(os.environ.get("ZB_E2E_INFLIGHT", 4))

    def make_out():
        return (torch.empty(max(n_trim, 1), dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                torch.empty(max(n_trim, 1), dtype=torch.int32).pin_memory().numpy().view(np.uint32))

    # With several steps in flight the LIBRARY hands the device's copy engines from step to step (one lock per engine
    # inside libzot_b200: copy-in for the H2D of the input, copy-out for the D2H of the result; kernels of different
    # steps overlap freely), so the steps of the host threads form a pipeline; the user code below is just the plain
    # sequence of API calls.
    import threading
    trace = [] if os.environ.get("ZB_E2E_TRACE") else None

    def step_e2e(out_k, out_c):
        tr = [threading.get_ident() % 1000, time.perf_counter()] if trace is not None else None
        km = nat.Kmerizer(K, dev)
        km.feed(h_in, False)
        if tr: tr.append(time.perf_counter())
        if dist_ctx is not None:
            exchange(nat, km, dist_ctx)
        s, nr = km.finish()
        km.close()
        t = s.trim(2)
        st = s.stats()
        if tr: tr.append(time.perf_counter())
        k_, c_ = t.fetch(out_k=out_k, out_c=out_c)
        s.free(); t.free()
        if tr:
            tr.append(time.perf_counter())
            trace.append(tr)
        return len(k_), st

    share = [args.steps // inflight + (1 if i < args.steps % inflight else 0) for i in range(inflight)]
    gate = threading.Barrier(inflight + 1)     # everybody is warm
    go = threading.Barrier(inflight + 1)       # the per-stage profile has been reset: the timed region starts
    results = [None] * inflight
    errors = []

    def worker(i):
        try:
            bufs = make_out()
            for _ in range(min(args.warmup, 2)):
                step_e2e(*bufs)
            gate.wait()
            go.wait()
            r = None
            for _ in range(share[i]):
                r = step_e2e(*bufs)
            results[i] = r
        except Exception as e:   # pragma: no cover
            errors.append(e)
            for b_ in (gate, go):
                try:
                    b_.abort()
                except Exception:
                    pass

    nat.dbg_profile(True, dev)
    if inflight == 1:
        bufs = make_out()
        for _ in range(min(args.warmup, 2)):
            step_e2e(*bufs)
        barrier()
        nat.dbg_profile(True, dev)
        e0 = time.perf_counter()
        e_times = []
        for _ in range(args.steps):
            t_it = time.perf_counter()
            nk, st = step_e2e(*bufs)
            e_times.append((time.perf_counter() - t_it) * 1e3)
    else:
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(inflight)]
        for t_ in ths:
            t_.start()
        e0 = time.perf_counter()
        try:
            gate.wait()
            nat.dbg_profile(True, dev)   # drop the warm-up stages (no worker touches the library between the two barriers)
            go.wait()
            e0 = time.perf_counter()
        except threading.BrokenBarrierError:
            pass
        for t_ in ths:
            t_.join()
        if errors:
            raise errors[0]
        nk, st = [r for r in results if r is not None][0]
        e_times = []
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / args.steps
    e_prof = nat.dbg_profile(False, dev)
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions (device-resident and e2e)
    if rank == 0:
        print("e2e (%d in flight) per-step wall ms: %s; stage ms/step: %s" % (
            inflight, [round(x, 1) for x in e_times] if e_times else round(e2e_ms, 2),
            {k: round(v[0] / args.steps, 3) for k, v in e_prof.items()}), file=sys.stderr)

        if trace:
            t00 = min(r[1] for r in trace)
            for r in sorted(trace, key=lambda r: r[1]):
                print("trace thread %3d: start %.2f | fed %.2f | counted + trimmed %.2f | fetched %.2f" % (
                    (r[0],) + tuple((x - t00) * 1e3 for x in r[1:])), file=sys.stderr)

    # ---------------- reduce over ranks (max time), aggregate throughput
    per_step_ms = ms_dev / args.steps   # CUDA events on the library stream around the K steps
    vals = [per_step_ms, e2e_ms, float(launches)]
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor(vals[:2], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per_step_ms, e2e_ms = float(tt[0]), float(tt[1])
        tl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])
    pairs = None
    if not args.no_pairs:
        pairs = bench_pairs(nat, dev, rank, world, max(2, min(args.steps, 3)))
    if rank != 0:
        return
    total_bases = bases * world
    value = total_bases / (per_step_ms * 1e-3) / 1e9
    e2e = total_bases / (e2e_ms * 1e-3) / 1e9

    peak, peak_kind = load_peaks()
    # dominant kernel: one onesweep radix pass over the canonical keys of this rank
    sp = prof.get("sort_pass_keys", (0.0, 0))
    pass_ms = sp[0] / sp[1] if sp[1] else None
    roofline = None
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}
    keys_per_step = stage_keys(nat, dev, d_in, nbytes) if pass_ms else None
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
                    "stage_ms_per_step": stage_ms}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nk * 12),
                "ms_per_step": e2e_ms, "steps_in_flight": inflight,
                "result": "trimmed (k-mer u64, count u32) arrays + count histogram of the full set"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "result": {"distinct_kmers": int(n_full), "after_trim": int(n_trim)},
    }
    if dist_ctx is not None and dist_ctx.get("p2p") is not None and "route_p2p" in prof:
        r_ms = prof["route_p2p"][0] / max(1, prof["route_p2p"][1])
        r_b = float(np.mean(dist_ctx["p2p"].remote_bytes)) if dist_ctx["p2p"].remote_bytes else 0.0
        line["nvlink"] = {"exchange": "fused: route_p2p_kernel stores every key into its owner's buffer over NVLink peer memory (CUDA IPC); "
                                      + ("a thread block reserves its run there with one system-scope atomic on the owner's cursor word; "
                                         "NCCL only for the barrier" if dist_ctx["p2p"].reserve else "NCCL only for the count matrix and the barrier"),
                          "route_kernel_ms": r_ms, "remote_bytes_per_gpu": r_b, "GBps_per_gpu_out": r_b / r_ms / 1e6 if r_ms else None,
                          "note": "the kernel also moves this rank's own share locally, so the NVLink rate is a lower bound",
                          "peak_GBps_per_direction": 900.0, "measured_peer_copy_GBps": 770.0}
    elif dist_ctx is not None and dist_ctx["a2a_ms"]:
        a_ms = float(np.mean(dist_ctx["a2a_ms"]))
        a_b = float(np.mean(dist_ctx["a2a_bytes"]))
        line["nvlink"] = {"all_to_all_ms": a_ms, "bytes_sent_per_gpu": a_b, "GBps_per_gpu_out": a_b / a_ms / 1e6,
                          "peak_GBps_per_direction": 900.0, "measured_peer_copy_GBps": 770.0}
    if pairs is not None:
        if world == 1 and not args.no_cpu_baseline:
            pairs["cpu_baseline"] = cpu_baseline_pairs(pairs)
        line["pairs"] = pairs
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


PAIR_SETS = int(os.environ.get("ZB_BENCH_PAIR_SETS", 32))


def bench_pairs(nat, dev, rank, world, steps):
    """BASELINE.json's second metric, pairwise Jaccard set-pairs/s, on a bounded instance of config[3]:
    PAIR_SETS synthetic bacterial k-mer sets (k=25, both strands, ~9.9 M k-mers each; 4 clades of related
    genomes), all pairs.  Device-resident: the sets live in HBM (zb_allpairs_abc, CUDA-event kernel time);
    e2e: the same call including the D2H of the (a, b, c) matrix and the Jaccard values on the host.
    N > 1: every rank holds all sets and computes its share of the work units (tile x key-range shard); one all-reduce
    adds the partial (a, b, c) up."""
    import torch
    from tools import synth
    from zotmer_b200 import multigpu
    nclades = max(1, PAIR_SETS // 8)
    base = [synth.genome(GENOME, seed=1000 + c) for c in range(nclades)]
    sets = []
    for i in range(PAIR_SETS):
        g = synth.mutate(base[i % nclades], 0.001 + 0.009 * (i // nclades) / max(1, PAIR_SETS // nclades), 2000 + i)
        km = nat.Kmerizer(K, dev)
        km.feed(synth.fasta_bytes(g), True)
        s, _ = km.finish()
        km.close()
        sets.append(s.project(0))     # Measure.prep: k-mers only (commands/dist.py:29-49)
        s.free()
    npairs = PAIR_SETS * (PAIR_SETS - 1) // 2
    b, e, st = multigpu.unit_share(PAIR_SETS, rank, world)
    dist = None
    if world > 1:
        import torch.distributed as dist

    def step():
        part = nat.allpairs_abc(sets, b, e, st)
        if world > 1:
            t = torch.from_numpy(part.view(np.int64)).to("cuda:%d" % dev)
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
    vals = [kern_ms, wall_ms]
    if world > 1:
        tt = torch.tensor(vals, dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kern_ms, wall_ms = float(tt[0]), float(tt[1])
    sizes = [len(s) for s in sets]
    pair_bytes = 8.0 * float(sum(sizes)) * (PAIR_SETS - 1)      # sum over pairs of 8 (|X| + |Y|)
    nblk = -(-PAIR_SETS // multigpu.AP_S)
    moved_bytes = 8.0 * float(sum(sizes)) * (1 + nblk)         # offsets pass + one gather per tile a set belongs to
    for s in sets:
        s.free()
    peak, _ = load_peaks()
    return {"metric": "pairwise Jaccard set-pairs/s", "value": npairs / (kern_ms * 1e-3), "unit": "set-pairs/s",
            "e2e": {"value": npairs / (wall_ms * 1e-3), "unit": "set-pairs/s", "d2h_bytes_per_step": int(npairs * 24)},
            "config": {"workload": "config[3] bounded: all pairs of %d synthetic bacterial k-mer sets (k=25, ~%d k-mers each, "
                                   "%d clades)" % (PAIR_SETS, int(np.mean(sizes)), nclades),
                       "pairs": npairs, "parallelism": "1 GPU" if world == 1 else "%d GPUs: work units (pairs of 32-set blocks x 8 key-range shards) sharded, one all-reduce" % world},
            "ms_per_step": kern_ms,
            "roofline": {"bound": "hbm", "achieved": moved_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": moved_bytes / (kern_ms * 1e-3) / 1e9 / (peak * world),
                         "note": "numerator = the bytes this design has to move: every k-mer (8 B) once for the bucket offsets and "
                                 "once per tile its set belongs to (N/32 block pairs); the kernel is bound by shared-memory work "
                                 "(hash dedupe, bit-matrix transpose), not by HBM",
                         "pair_at_a_time_model_GBps": pair_bytes / (kern_ms * 1e-3) / 1e9,
                         "pair_at_a_time_note": "8 (|X| + |Y|) B per pair (SURVEY.md 8d) over the same time: what a merge per pair would "
                                                "have to sustain to keep up"},
            "check": {"jaccard_first_pair": float(jac[0]), "max_jaccard": float(jac.max())}}


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
    xs, cs, h, acgt, nr = zo.kmerize_core(K, [("reads.fq", fq)])
    zo.words_to_bytes(zo.encode(zo.delta(xs)))
    zo.words_to_bytes(zo.encode(cs))
    tx, tc = zo.trim_core(xs, cs, 2, None)
    zo.words_to_bytes(zo.encode(zo.delta(tx)))
    zo.words_to_bytes(zo.encode(tc))
    dt = time.perf_counter() - t0
    import shutil
    pypy = shutil.which("pypy") or shutil.which("pypy3")
    return {"value": sample_reads * READ_LEN / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "pypy": pypy or "not installed on this box and not installable offline (north star: PyPy only if it installs offline)",
            "sample": "first %d of the %d reads (%.1f Mbases, %.1f s): kmerize+count, codec64 encode, trim; CPython %s "
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
    ap.add_argument("--no-pairs", action="store_true", help="skip the set-pairs/s measurement")
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
