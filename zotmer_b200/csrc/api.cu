// api.cu -- the C ABI (include/zotmer_b200.h): handles, memory, orchestration of the kernels.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"

namespace zb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// One context (stream, allocator, scratch scalars) per (device, host thread): two host threads that drive the same
// GPU get two streams, so the H2D copy of one batch overlaps the kernels of another (bench.py's e2e does that).
static std::mutex g_ctx_mu;
static std::map<std::pair<int, std::thread::id>, Ctx*> g_ctx;
static bool g_profile_on[64] = {false};

static std::vector<Ctx*> contexts_of(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    std::vector<Ctx*> v;
    for (auto& kv : g_ctx)
        if (kv.first.first == device) v.push_back(kv.second);
    return v;
}

Ctx* ctx_for(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s): libzot_b200 has no CPU path", cudaGetErrorString(e));
        throw Fail{ZB_E_NOGPU};
    }
    if (device < 0 || device >= ndev) ZB_FAIL(ZB_E_ARG, "device %d out of range (have %d)", device, ndev);
    ZB_CUDA(cudaSetDevice(device));
    const auto key = std::make_pair(device, std::this_thread::get_id());
    auto it = g_ctx.find(key);
    if (it != g_ctx.end()) return it->second;
    Ctx* c = new Ctx();
    c->profile = g_profile_on[device & 63];
    c->device = device;
    cudaDeviceProp prop;
    ZB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    ZB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ZB_CUDA(cudaHostAlloc((void**)&c->h_scalars, 64 * sizeof(uint64_t), cudaHostAllocMapped));
    ZB_CUDA(cudaHostGetDevicePointer((void**)&c->d_h_scalars, c->h_scalars, 0));
    ZB_CUDA(cudaHostAlloc((void**)&c->h_big, H_BIG, cudaHostAllocMapped));
    ZB_CUDA(cudaHostGetDevicePointer((void**)&c->d_h_big, c->h_big, 0));
    g_ctx[key] = c;
    return c;
}

__global__ void readback_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_host, int nwords) {
    if ((int)threadIdx.x < nwords) dst_host[threadIdx.x] = src[threadIdx.x];
}

cudaError_t read_back(Ctx* c, const void* d_src, size_t bytes) {
    if (bytes > 128 || (bytes & 3)) return cudaErrorInvalidValue;
    readback_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(d_src), c->d_h_scalars, (int)(bytes / 4));
    c->launches++;
    return cudaGetLastError();
}

// generic fill: bytes up to the first 16-byte boundary and after the last one singly, the body as uint4
__global__ void __launch_bounds__(256) fill_kernel(uint8_t* __restrict__ p, uint8_t v, size_t n) {
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)p & 15)) & 15));
    const size_t body = (n - head) / 16;
    const size_t t0 = (size_t)blockIdx.x * 256 + threadIdx.x, stride = (size_t)gridDim.x * 256;
    const uint32_t w = 0x01010101u * v;
    uint4* __restrict__ q = reinterpret_cast<uint4*>(p + head);
    for (size_t i = t0; i < body; i += stride) q[i] = make_uint4(w, w, w, w);
    if (t0 < head) p[t0] = v;
    const size_t tail0 = head + body * 16;
    if (t0 < n - tail0) p[tail0 + t0] = v;
}

template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(T* __restrict__ dst, const T* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

cudaError_t read_back_big(Ctx* c, const void* d_src, size_t bytes) {
    if (bytes > H_BIG || (bytes & 15) || ((uintptr_t)d_src & 15)) return cudaErrorInvalidValue;
    copy_kernel<uint4><<<(unsigned)div_up(bytes / 16, 256), 256, 0, c->stream>>>((uint4*)c->d_h_big, (const uint4*)d_src, bytes / 16);
    c->launches++;
    return cudaGetLastError();
}

void copy_small_to_host(Ctx* c, const void* d_src, size_t host_off, size_t bytes) {
    if (host_off + bytes > H_BIG || (bytes & 15) || (host_off & 15) || ((uintptr_t)d_src & 15)) ZB_FAIL(ZB_E_ARG, "copy_small_to_host: bad range");
    copy_kernel<uint4><<<(unsigned)div_up(bytes / 16, 256), 256, 0, c->stream>>>((uint4*)(c->d_h_big + host_off), (const uint4*)d_src, bytes / 16);
    ZB_LAUNCH_CHECK(c);
}

cudaError_t dev_memset(Ctx* c, void* p, int value, size_t bytes) {
    if (bytes == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(std::max<size_t>(bytes / 16, 32), 256));
    fill_kernel<<<blocks, 256, 0, c->stream>>>(reinterpret_cast<uint8_t*>(p), (uint8_t)value, bytes);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t dev_copy(Ctx* c, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return cudaSuccess;
    const uintptr_t al = (uintptr_t)dst | (uintptr_t)src | (uintptr_t)bytes;
    const size_t unit = (al & 15) == 0 ? 16 : (al & 7) == 0 ? 8 : (al & 3) == 0 ? 4 : 1;
    const size_t n = bytes / unit;
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(n, 256));
    if (unit == 16) copy_kernel<uint4><<<blocks, 256, 0, c->stream>>>((uint4*)dst, (const uint4*)src, n);
    else if (unit == 8) copy_kernel<uint64_t><<<blocks, 256, 0, c->stream>>>((uint64_t*)dst, (const uint64_t*)src, n);
    else if (unit == 4) copy_kernel<uint32_t><<<blocks, 256, 0, c->stream>>>((uint32_t*)dst, (const uint32_t*)src, n);
    else copy_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>((uint8_t*)dst, (const uint8_t*)src, n);
    c->launches++;
    return cudaGetLastError();
}

// ZB_GUARD=1: every block carries GUARD pattern bytes in front and behind (checked when the block is released and by
// zb_dbg_guard_check).  The pool's compute-sanitizer is closed, so this is how an out-of-bounds store next to a library
// buffer is caught in the tests.
static const size_t GUARD = 256;
static bool guard_on() {
    static const bool on = [] { const char* e = getenv("ZB_GUARD"); return e && atoi(e) != 0; }();
    return on;
}

__global__ void guard_check_kernel(const uint64_t* __restrict__ blocks /*[n][2] = user pointer, size*/, uint32_t n,
                                   unsigned long long* __restrict__ n_bad) {
    const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(blocks[2 * b]);
    const uint64_t sz = blocks[2 * b + 1];
    bool bad = false;
    for (unsigned i = lane_id(); i < 256; i += 32) bad |= (p[(long long)i - 256] != 0xA5) | (p[sz + i] != 0xA5);
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicAdd(n_bad, 1ull);
}

// Device memory: a caching allocator per context.  cudaMalloc'ed SEGMENTS are cut into blocks; a request takes the
// smallest free block that fits and splits off the rest; a released block merges with free neighbours.  Requests
// below 1 MiB live in their own 2 MiB segments so that scalars never nibble at the gigabyte blocks.  All work of a
// context is ordered on its one stream, so a block released by the host can be handed out again at once.
// (History: cudaMallocAsync pools showed 100-600 ms re-mapping stalls between steps; the first in-house version kept
// whole cudaMalloc blocks in size classes and, when a long kmerize had left 60 GB of them behind, a later small
// request missed, evicted, and paid 100-170 ms of cudaFree -- inside whatever stage happened to allocate next,
// bench.py's human leg: "stats 168 ms" for an 8 ms kernel.)
static const size_t SMALL_LIMIT = (size_t)1 << 20;
static const size_t SMALL_SEG = (size_t)2 << 20;
static const size_t MIN_SPLIT = (size_t)1 << 20;     // a large block is split only when the rest is at least this

static size_t round_req(size_t b) {
    if (b < 512) return 512;
    return (b + 511) & ~(size_t)511;
}

static void idx_erase(Ctx* c, Ctx::Seg* sg, size_t off, size_t size) {
    auto& idx = c->free_idx[sg->small ? 1 : 0];
    auto rg = idx.equal_range(size);
    for (auto it = rg.first; it != rg.second; ++it)
        if (it->second.first == sg && it->second.second == off) { idx.erase(it); return; }
}

static void seg_release(Ctx* c, Ctx::Seg* sg) {   // a wholly free segment goes back to the driver
    for (auto& kv : sg->blocks)
        if (kv.second.free) idx_erase(c, sg, kv.first, kv.second.size);
    c->cached_bytes -= sg->free_bytes;
    c->segs.erase(sg->base);
    cudaFree(sg->base);
    delete sg;
}

static void release_free_segments(Ctx* c, size_t down_to) {
    std::vector<std::pair<uint64_t, Ctx::Seg*>> byage;
    for (auto& kv : c->segs)
        if (kv.second->free_bytes == kv.second->size) byage.push_back({kv.second->last_use, kv.second});
    std::sort(byage.begin(), byage.end());
    for (auto& ap : byage) {
        if (c->cached_bytes <= down_to) break;
        seg_release(c, ap.second);
    }
}

void* dalloc(Ctx* c, size_t bytes) {
    std::lock_guard<std::mutex> lk(c->alloc_mu);   // a set may be freed by another thread than the one that made it
    const size_t user = round_req(bytes);
    const size_t want = user + (guard_on() ? 2 * GUARD : 0);
    const bool small = want < SMALL_LIMIT;
    auto& idx = c->free_idx[small ? 1 : 0];
    Ctx::Seg* sg = nullptr;
    size_t off = 0;
    auto it = idx.lower_bound(want);
    if (it != idx.end()) {
        sg = it->second.first;
        off = it->second.second;
        idx.erase(it);
    } else {
        // a miss: a new segment.  Above the limit, wholly free segments (least recently used first) go back first.
        if (c->cache_limit == 0) {
            size_t fr = 0, tot = 0;
            c->cache_limit = (cudaMemGetInfo(&fr, &tot) == cudaSuccess && tot) ? tot / 4 : ((size_t)32 << 30);
        }
        if (c->cached_bytes > c->cache_limit) release_free_segments(c, c->cache_limit / 2);
        const size_t seg_size = small ? SMALL_SEG : ((want + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1));
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, seg_size);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(c->stream);
            release_free_segments(c, 0);
            e = cudaMalloc(&p, seg_size);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("out of device memory: %zu bytes requested, %zu live, %zu cached in partly used segments", want, c->live_bytes,
                      c->cached_bytes);
            throw Fail{ZB_E_NOMEM};
        }
        sg = new Ctx::Seg();
        sg->base = (char*)p;
        sg->size = seg_size;
        sg->small = small;
        sg->blocks[0] = Ctx::Blk{seg_size, true};
        sg->free_bytes = seg_size;
        c->segs[sg->base] = sg;
        c->cached_bytes += seg_size;
        off = 0;
    }
    Ctx::Blk& blk = sg->blocks[off];
    const size_t rest = blk.size - want;
    if (rest >= (small ? (size_t)512 : MIN_SPLIT)) {
        blk.size = want;
        sg->blocks[off + want] = Ctx::Blk{rest, true};
        idx.insert({rest, {sg, off + want}});
    }
    Ctx::Blk& mine = sg->blocks[off];
    mine.free = false;
    sg->free_bytes -= mine.size;
    sg->last_use = ++c->tick;
    c->cached_bytes -= mine.size;
    c->live_bytes += mine.size;
    char* up = sg->base + off;
    const size_t asked = (bytes + 15) & ~(size_t)15;
    if (guard_on()) {
        dev_memset(c, up, 0xA5, GUARD);
        dev_memset(c, up + GUARD + asked, 0xA5, mine.size - GUARD - asked);   // everything behind the request is band
        up += GUARD;
    }
    c->live_blocks[up] = Ctx::Live{sg, off, asked};
    return up;
}

void dfree(Ctx* c, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(c->alloc_mu);
    auto it = c->live_blocks.find(p);
    if (it == c->live_blocks.end()) return;
    Ctx::Seg* sg = it->second.seg;
    size_t off = it->second.off;
    const uint64_t span[2] = {(uint64_t)(uintptr_t)p, (uint64_t)it->second.user};
    c->live_blocks.erase(it);
    if (guard_on()) {
        // the bands are looked at before the block can be handed out again (stream order); damage is tallied in the context
        uint64_t* d = nullptr;
        if (cudaMalloc((void**)&d, 24) == cudaSuccess) {
            cudaMemsetAsync(d, 0, 8, c->stream);
            cudaMemcpyAsync(d + 1, span, 16, cudaMemcpyHostToDevice, c->stream);
            guard_check_kernel<<<1, 32, 0, c->stream>>>(d + 1, 1u, reinterpret_cast<unsigned long long*>(d));
            uint64_t bad = 0;
            cudaMemcpyAsync(&bad, d, 8, cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
            cudaFree(d);
            c->guard_bad += bad;
        }
    }
    auto bi = sg->blocks.find(off);
    size_t size = bi->second.size;
    c->live_bytes -= size;
    c->cached_bytes += size;
    sg->free_bytes += size;
    sg->last_use = ++c->tick;
    // merge with the free neighbours
    auto nx = std::next(bi);
    if (nx != sg->blocks.end() && nx->second.free) {
        idx_erase(c, sg, nx->first, nx->second.size);
        size += nx->second.size;
        sg->blocks.erase(nx);
    }
    if (bi != sg->blocks.begin()) {
        auto pv = std::prev(bi);
        if (pv->second.free) {
            idx_erase(c, sg, pv->first, pv->second.size);
            size += pv->second.size;
            off = pv->first;
            sg->blocks.erase(bi);
            bi = pv;
        }
    }
    bi->second.size = size;
    bi->second.free = true;
    c->free_idx[sg->small ? 1 : 0].insert({size, {sg, off}});
}

// The block behind `p` keeps its first `bytes` bytes; what lies behind them goes back to the free list (merged with a free
// neighbour) without a copy.  A no-op when the rest is too small to be worth a block of its own, and under ZB_GUARD (the
// band behind the request would have to move).
void dshrink(Ctx* c, void* p, size_t bytes) {
    if (!p || guard_on()) return;
    std::lock_guard<std::mutex> lk(c->alloc_mu);
    auto it = c->live_blocks.find(p);
    if (it == c->live_blocks.end()) return;
    Ctx::Seg* sg = it->second.seg;
    const size_t off = it->second.off;
    auto bi = sg->blocks.find(off);
    const size_t want = round_req(bytes);
    if (sg->small || bi->second.size < want + MIN_SPLIT) return;
    size_t rest = bi->second.size - want;
    bi->second.size = want;
    it->second.user = (bytes + 15) & ~(size_t)15;
    c->live_bytes -= rest;
    c->cached_bytes += rest;
    sg->free_bytes += rest;
    auto nx = std::next(bi);
    if (nx != sg->blocks.end() && nx->second.free) {
        idx_erase(c, sg, nx->first, nx->second.size);
        rest += nx->second.size;
        sg->blocks.erase(nx);
    }
    sg->blocks[off + want] = Ctx::Blk{rest, true};
    c->free_idx[0].insert({rest, {sg, off + want}});
}

void dtrim(Ctx* c) {
    std::lock_guard<std::mutex> lk(c->alloc_mu);
    release_free_segments(c, 0);
}

}  // namespace zb

using namespace zb;

struct zb_kmerizer {
    Ctx* c;
    int k;
    DBuf<uint64_t> pending;              // canonical keys not yet counted
    size_t pending_cap = 0;
    size_t pending_upper = 0;            // host-side upper bound of the device counter
    DBuf<unsigned long long> d_count;    // [0] number of pending keys
    // Counted canonical runs, one per batch (each sorted, duplicate-free).  They are NOT folded into one running set
    // batch by batch: when max_runs of them have piled up (or at finish) they are united in ONE pass by the N-way
    // bucket merge (nwaymerge.cu), slab by slab of the key space.  (The first version merged every batch into the
    // running set with merge-path + reduce-by-key, i.e. re-read and re-wrote the whole accumulated set 25 times for
    // the human-scale read set: more than half of that run's time, profiles/r02_human_scale.md.)
    struct Run { DBuf<uint64_t> k; DBuf<uint32_t> c; size_t n = 0; };
    std::vector<Run> runs;
    // Runs pile up until there are max_runs of them or they hold `run_budget` bytes (a share of the device's memory).
    // Every compaction re-reads and re-writes what has been accumulated so far, so few of them is better -- human-scale
    // shape on one B200: 10.5 / 14.9 / 17.3 Gbases/s with a compaction every 5 / 8 / 13 batches (gpurun_out/r3_human_r*.json)
    // -- but one merge of all 25 runs at the end is no better than two of 13 (16.5, r3_human_c55.json: the slices a bucket
    // gets from each input become short).
    size_t max_runs = 13;
    size_t run_budget = 0;               // bytes; 0 = not yet asked (ZB_RUN_BUDGET_PCT of the device's memory, default 35)
    DBuf<uint64_t> acc_k;                // the one run left by compact_runs at finish
    DBuf<uint32_t> acc_c;
    size_t acc_n = 0;
    uint64_t n_records = 0;
    size_t max_pending = (size_t)1 << 29;
    const zb_set* baits = nullptr;       // capture mode (`zot kmerize -C`): only records that hold one of these k-mers count
    uint64_t* adopted = nullptr;         // caller-owned key array that stands in for `pending` (multi-GPU receive buffer)
    size_t adopted_n = 0;
    DBuf<unsigned long long> route_cur;  // cursors of a routing kernel in flight on the side stream (route_p2p_begin / _end)
    bool routing = false;
    int owners = 0;                      // zb_kmerize_set_owners: extraction tallies the keys per owner among this many GPUs
    DBuf<unsigned long long> owner_cnt;  // [64] tallies of the pending keys
    bool owner_cnt_valid = false;        // every pending key came through the tallying extraction
    // digit histograms of the pending keys, tallied by the extraction for the plan the coming sort is EXPECTED to use
    // (chosen from the size of the first piece of a batch; the sort checks it against its real plan)
    SortPre pre;
    DBuf<uint32_t> pre_hist;             // [4][256]
    bool pre_on = false;                 // every pending key is in pre_hist
};

static size_t read_pending_count(zb_kmerizer* h) {
    Ctx* c = h->c;
    ZB_CUDA(read_back(c, h->d_count.get(), 8));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return (size_t)c->h_scalars[0];
}

static void ensure_pending(zb_kmerizer* h, size_t need_total) {
    if (need_total <= h->pending_cap) return;
    Ctx* c = h->c;
    size_t ncap = std::max(need_total, h->pending_cap * 2);
    DBuf<uint64_t> nb(c, ncap);
    if (h->pending_upper) {
        size_t have = read_pending_count(h);
        ZB_CUDA(dev_copy(c, nb.get(), h->pending.get(), have * 8));
    }
    h->pending = std::move(nb);
    h->pending_cap = ncap;
}

// the pending list has been consumed (routed, bucketed or counted): the owner tallies start again
static void pending_consumed(zb_kmerizer* h) {
    h->pending_upper = 0;
    h->pre_on = false;
    ZB_CUDA(dev_memset(h->c, h->d_count.get(), 0, 8));
    if (h->owners > 1) {
        ZB_CUDA(dev_memset(h->c, h->owner_cnt.get(), 0, 64 * 8));
        h->owner_cnt_valid = true;
    }
}

// count the pending canonical keys: one more run
static void count_keys(zb_kmerizer* h, uint64_t* keys, size_t n, const SortPre* pre = nullptr, bool last = false);
static void compact_runs(zb_kmerizer* h);

// last: no batch follows (zb_kmerize_finish)
static void flush_pending(zb_kmerizer* h, bool last = false) {
    Ctx* c = h->c;
    if (h->adopted) {
        uint64_t* keys = h->adopted;
        const size_t n = h->adopted_n;
        h->adopted = nullptr;
        h->adopted_n = 0;
        count_keys(h, keys, n);
    }
    if (h->pending_upper == 0) return;
    const size_t n = read_pending_count(h);
    const bool pre_on = h->pre_on;
    pending_consumed(h);
    count_keys(h, h->pending.get(), n, pre_on ? &h->pre : nullptr, last);
    (void)c;
}

// sort + count `keys` (destroyed) and fold the result into the accumulated run
static void count_keys(zb_kmerizer* h, uint64_t* keys, size_t n, const SortPre* pre, bool last) {
    Ctx* c = h->c;
    if (n == 0) return;
    // sort + count in one go (segsort.cu): distinct canonical keys -> `dk`, counts -> `dc`
    DBuf<uint64_t> tmp(c, n), dk(c, n);
    DBuf<uint32_t> dc(c, n);
    const size_t nd = sort_count(c, keys, tmp.get(), nullptr, nullptr, n, 2 * h->k, dk.get(), dc.get(), false, pre);
    tmp.release();
    // the run at its real size (sort_count's outputs are sized for n distinct keys).  The last batch of a kmerizer keeps the
    // buffers and gives what lies behind its nd entries back to the allocator (no copy: 0.08 ms of a bench step).  A batch in
    // the middle of a long input copies its run into buffers of its own instead: shrunk blocks pin their segments, the
    // 3 GB buffers of the NEXT batch then find no free block of their size, and a cudaMalloc per batch costs far more than
    // the copy (human-scale shape with 17 runs held: 1.8 s instead of 0.65, gpurun_out/r3_human_b35.json).
    zb_kmerizer::Run r;
    r.n = nd;
    if (last) {
        dk.shrink(nd);
        dc.shrink(nd);
        r.k = std::move(dk);
        r.c = std::move(dc);
    } else {
        r.k.alloc(c, nd);
        r.c.alloc(c, nd);
        ZB_CUDA(dev_copy(c, r.k.get(), dk.get(), nd * 8));
        ZB_CUDA(dev_copy(c, r.c.get(), dc.get(), nd * 4));
        dk.release();
        dc.release();
    }
    h->runs.push_back(std::move(r));
    if (h->run_budget == 0) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || tot == 0) tot = (size_t)64 << 30;
        int pct = 35;
        if (const char* e = getenv("ZB_RUN_BUDGET_PCT")) pct = std::min(80, std::max(1, atoi(e)));
        h->run_budget = tot / 100 * (size_t)pct;
    }
    size_t held = 0;
    for (const auto& q : h->runs) held += q.n * 12;
    if (h->runs.size() >= h->max_runs || (h->runs.size() > 1 && held > h->run_budget)) compact_runs(h);
}

// the pairwise fold of the first version: merge path + reduce-by-key, run after run (fallback for key spaces too
// skewed for the bucket merge)
static void fold_runs_pairwise(zb_kmerizer* h) {
    Ctx* c = h->c;
    while (h->runs.size() > 1) {
        zb_kmerizer::Run b = std::move(h->runs.back());
        h->runs.pop_back();
        zb_kmerizer::Run a = std::move(h->runs.back());
        h->runs.pop_back();
        const size_t tot = a.n + b.n;
        DBuf<uint64_t> mk(c, tot);
        DBuf<uint32_t> mc(c, tot);
        merge_pairs(c, a.k.get(), a.c.get(), a.n, b.k.get(), b.c.get(), b.n, mk.get(), mc.get());
        a.k.release(); a.c.release(); b.k.release(); b.c.release();
        zb_kmerizer::Run r;
        r.k.alloc(c, tot);
        r.c.alloc(c, tot);
        r.n = reduce_by_key(c, mk.get(), mc.get(), tot, r.k.get(), r.c.get());
        h->runs.push_back(std::move(r));
    }
}

// all runs -> one run (counts of equal k-mers summed): KmerAccumulator2's merge of a flushed batch into the counted
// run (kmerize.py:41-132, :412-424), done for several batches at once.  The key space is cut into slabs at quantiles of
// the largest run so that the bucket merge's staging (12 B per input entry of a slab) stays below ~1/16 of the device.
static void compact_runs(zb_kmerizer* h) {
    Ctx* c = h->c;
    const size_t nr = h->runs.size();
    if (nr <= 1) return;
    Stage st(c, "compact_runs");
    size_t fr = 0, tot_mem = 0;
    if (cudaMemGetInfo(&fr, &tot_mem) != cudaSuccess || tot_mem == 0) tot_mem = (size_t)64 << 30;
    const size_t slab_entries = std::max<size_t>((size_t)1 << 24, std::min<size_t>(tot_mem / 16 / 12, ((size_t)1 << 32) - 1));
    std::vector<const uint64_t*> ks;
    std::vector<const uint32_t*> cs;
    std::vector<size_t> ns;
    for (size_t i = 0; i < nr; i++) { ks.push_back(h->runs[i].k.get()); cs.push_back(h->runs[i].c.get()); ns.push_back(h->runs[i].n); }
    zb_kmerizer::Run r;
    const bool ok = merge_nway_slabs(c, ks, cs, ns, 2 * h->k, slab_entries, &r.k, &r.c, &r.n);
    if (!ok) {   // key space too skewed for the buckets: fold pairwise instead
        fold_runs_pairwise(h);
        return;
    }
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    h->runs.clear();
    h->runs.push_back(std::move(r));
}

// extraction of a parsed code stream (device buffer with 32-byte front pad and tile tail pad)
static bool extract_hist_on() {
    static const bool on = [] { const char* e = getenv("ZB_EXTRACT_HIST"); return !e || atoi(e) != 0; }();
    return on;
}

// keys_hint: how many keys the caller expects these codes to give (0: unknown) -- only used to pick the digit plan
static void extract_codes(zb_kmerizer* h, const uint8_t* codes, size_t n_codes, size_t keys_hint = 0) {
    Ctx* c = h->c;
    size_t off = 0;
    while (off < n_codes) {
        if (h->pending_upper >= h->max_pending) flush_pending(h);
        size_t room = h->max_pending - h->pending_upper;
        size_t len = std::min(n_codes - off, std::max<size_t>(room, EXTRACT_TILE));
        if (len < n_codes - off) len = (len / EXTRACT_TILE) * EXTRACT_TILE;  // slices end on tile boundaries
        // a slice may overshoot into the next tile: bound by whole tiles
        const size_t upper = div_up(len, EXTRACT_TILE) * EXTRACT_TILE;
        ensure_pending(h, h->pending_upper + upper);
        if (h->pending_upper == 0 && h->owners <= 1 && extract_hist_on()) {
            // a new batch: the sort that will count it takes its digit histograms from the extraction, for the plan
            // that fits the number of keys this piece is expected to give
            const size_t expect = (keys_hint && len == n_codes) ? keys_hint : len;
            h->pre_on = sort_count_plan(expect, 2 * h->k, &h->pre);
            if (h->pre_on) {
                if (!h->pre_hist.get()) h->pre_hist.alloc(c, 4 * 256);
                h->pre.d_hist = h->pre_hist.get();
                ZB_CUDA(dev_memset(c, h->pre_hist.get(), 0, 4 * 256 * 4));
            }
        }
        if (h->owners > 1) h->pre_on = false;   // the tallying extraction for several owners keeps no digit histograms
        Stage st(c, "extract");
        extract_canonical(c, h->k, codes + off, len, h->pending.get(), h->d_count.get(), h->owners,
                          (h->owners > 1 && h->owner_cnt_valid) ? h->owner_cnt.get() : nullptr, h->pre_on ? &h->pre : nullptr);
        h->pending_upper += upper;
        off += len;
    }
}

static void feed_dev_impl(zb_kmerizer* h, const uint8_t* d_raw, size_t n, int is_fasta) {
    Ctx* c = h->c;
    if (n == 0) return;
    if (n >= ((size_t)1 << 31)) ZB_FAIL(ZB_E_ARG, "feed: piece of %zu bytes; split the input at record boundaries below 2 GiB", n);
    if (((uintptr_t)d_raw & 15) != 0) ZB_FAIL(ZB_E_ARG, "feed_dev: device pointer must be 16-byte aligned");
    const size_t cap = 32 + n + 2 * EXTRACT_TILE + 64;
    DBuf<uint8_t> codes(c, cap);
    uint8_t* cd = codes.get() + 32;
    ZB_CUDA(dev_memset(c, codes.get(), 4, 32));
    size_t n_codes = 0;
    uint64_t n_rec = 0;
    {
        Stage st(c, "parse");
        if (is_fasta) parse_fasta(c, d_raw, n, cd, &n_codes, &n_rec, h->baits != nullptr);
        else parse_fastq(c, d_raw, n, cd, &n_codes, &n_rec, h->baits != nullptr);
    }
    h->n_records += n_rec;
    if (n_codes == 0) return;
    const size_t padded = div_up(n_codes, EXTRACT_TILE) * EXTRACT_TILE + 32;
    ZB_CUDA(dev_memset(c, cd + n_codes, 4, padded - n_codes));
    if (h->baits) capture_records(c, h->k, cd, n_codes, h->baits->k.get(), h->baits->n);
    // every record of L bases gives L - k + 1 windows and one break code (fewer where it holds other letters)
    const uint64_t lost = n_rec * (uint64_t)h->k;
    extract_codes(h, cd, n_codes, n_codes > lost ? (size_t)(n_codes - lost) : 1);
}

namespace zb {
// Large host <-> device copies go out in 64 MiB pieces with at most three of them queued: a copy engine serves the
// streams of all host threads in the order the transfers were queued, so whatever another thread's step needs from
// the engine waits for three pieces, not for a whole 315 MB input (measured: with one 315 MB copy queued the other
// thread's sort + count took 13 ms instead of 6).
void copy_chunked(Ctx* c, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    static const size_t piece_mb = [] { const char* e = getenv("ZB_COPY_PIECE_MB"); return e ? (size_t)atoi(e) : (size_t)64; }();
    const size_t piece = piece_mb << 20;   // ZB_COPY_PIECE_MB=0: one unthrottled copy
    if (piece == 0 || bytes <= 2 * piece) {
        ZB_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, c->stream));
        return;
    }
    if (!c->copy_ev[0])
        for (int i = 0; i < 3; i++) ZB_CUDA(cudaEventCreateWithFlags(&c->copy_ev[i], cudaEventDisableTiming));
    size_t i = 0;
    for (size_t o = 0; o < bytes; o += piece, i++) {
        if (i >= 3) ZB_CUDA(cudaEventSynchronize(c->copy_ev[i % 3]));
        ZB_CUDA(cudaMemcpyAsync((char*)dst + o, (const char*)src + o, std::min(piece, bytes - o), kind, c->stream));
        ZB_CUDA(cudaEventRecord(c->copy_ev[i % 3], c->stream));
    }
}

// The copy engines of a device are handed from host thread to host thread: the bulk copies of a kmerize step (H2D
// of the input, D2H of the result) each hold their engine's lock, so several host threads that drive the same GPU
// copy one after the other at full PCIe rate while their kernels overlap freely (measured on bench.py's e2e with 3 - 4
// steps in flight: 7.1 - 7.3 ms/step; without the copy locks 12.0 -- the engine interleaves the pieces of all copies
// and every step waits for all of them; with a lock around the SM phases as well 8.7 - 9.0 -- the host round trips
// inside a phase, a dozen few-byte read-backs, then leave the GPU idle).  Locks are never nested; a single-threaded
// caller never waits.  ZB_ENGINE_LOCKS overrides the policy (bit 0 copy-in, bit 1 SMs, bit 2 copy-out; default 5).
std::mutex g_engine_mu[64][3];
int engine_lock_mask() {
    static const int m = [] { const char* e = getenv("ZB_ENGINE_LOCKS"); return e ? atoi(e) : 5; }();
    return m;
}

}  // namespace zb

// true counts of a wide set's entries that survive in a subset of it (defined with the wide-set code further down)
static void zb_set_carry_wide(const zb_set* in, zb_set* out);

// ZB_E_RANGE for a set with counts beyond 2^32-1
static void no_wide(const zb_set* s, const char* what) {
    if (s && s->wide.get()) ZB_FAIL(ZB_E_RANGE, "%s: the set holds counts beyond 2^32-1 (only merge output, stats, encode and fetch handle them)", what);
}

static zb_set* new_set(Ctx* c, size_t n) {
    zb_set* s = new zb_set();
    s->c = c;
    s->n = n;
    s->k.alloc(c, n);
    s->cnt.alloc(c, n);
    return s;
}

// hostio.cu: a staged piece is fed in stream order (the kmerizer's stream already waits for the copies)
int zb_kmerize_feed_dev_ordered(zb_kmerizer* h, const uint8_t* d_raw, size_t n, int is_fasta) {
    ZB_TRY
    EngineLock el(h->c, ENG_SM);
    feed_dev_impl(h, d_raw, n, is_fasta);
    ZB_CATCH
}
zb::Ctx* zb_kmerizer_ctx(zb_kmerizer* h) { return h->c; }

extern "C" {

const char* zb_last_error(void) { return zb::g_err; }
int zb_version(void) { return 100; }

int zb_device_count(int* n) {
    int nd = 0;
    cudaError_t e = cudaGetDeviceCount(&nd);
    if (e != cudaSuccess) nd = 0;
    if (n) *n = nd;
    return ZB_OK;
}

int zb_launch_count(int device, uint64_t* n) {
    ZB_TRY
    ctx_for(device);
    uint64_t tot = 0;
    for (Ctx* c : contexts_of(device)) tot += c->launches;   // every host thread's context on this device
    *n = tot;
    ZB_CATCH
}

int zb_release_cache(int device) {
    ZB_TRY
    ctx_for(device);
    for (Ctx* c : contexts_of(device)) {
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        dtrim(c);
    }
    ZB_CATCH
}

int zb_dev_read_small(int device, const void* d_src, size_t bytes, void* host_dst) {
    ZB_TRY
    if (bytes == 0) return ZB_OK;
    if (!d_src || !host_dst || bytes > H_BIG || (bytes & 3) || ((uintptr_t)d_src & 3)) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = ctx_for(device);
    copy_kernel<uint32_t><<<(unsigned)div_up(bytes / 4, 256), 256, 0, c->stream>>>((uint32_t*)c->d_h_big, (const uint32_t*)d_src, bytes / 4);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(host_dst, c->h_big, bytes);
    ZB_CATCH
}

int zb_dev_write_small(int device, void* d_dst, const void* host_src, size_t bytes) {
    ZB_TRY
    if (bytes == 0) return ZB_OK;
    if (!d_dst || !host_src || bytes > H_BIG || (bytes & 3) || ((uintptr_t)d_dst & 3)) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = ctx_for(device);
    memcpy(c->h_big, host_src, bytes);
    copy_kernel<uint32_t><<<(unsigned)div_up(bytes / 4, 256), 256, 0, c->stream>>>((uint32_t*)d_dst, (const uint32_t*)c->d_h_big, bytes / 4);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_device_sync(int device) {
    ZB_TRY
    ctx_for(device);
    for (Ctx* c : contexts_of(device)) ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

// ------------------------------------------------------------------------------- kmerize
int zb_kmerize_open(int k, int device, zb_kmerizer** out) {
    ZB_TRY
    if (!out) ZB_FAIL(ZB_E_ARG, "null out");
    if (k < 1 || k > 32) ZB_FAIL(ZB_E_ARG, "k=%d outside 1..32 (k-mers are packed into 64 bits)", k);
    Ctx* c = ctx_for(device);
    zb_kmerizer* h = new zb_kmerizer();
    h->c = c;
    h->k = k;
    h->d_count.alloc(c, 2);
    ZB_CUDA(dev_memset(c, h->d_count.get(), 0, 16));
    {   // tuning / cross-check switches, re-read at every open (unset = default)
        const char* e = getenv("ZB_SORT_COUNT");
        g_sort_count_mode = e ? atoi(e) : 0;
        e = getenv("ZB_SORT_CFG");
        g_sort_cfg = e ? atoi(e) : 0;
        e = getenv("ZB_MM_CFG");
        g_mm_cfg = e ? atoi(e) : 0;
        e = getenv("ZB_ROUTE_PER");
        g_route_per = (e && atoi(e) == 16) ? 16 : 8;
    }
    if (const char* e = getenv("ZB_MAX_RUNS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 1000) h->max_runs = (size_t)v;
    }
    if (const char* e = getenv("ZB_MAX_PENDING")) {
        size_t v = strtoull(e, nullptr, 10);
        if (v >= (size_t)EXTRACT_TILE && v < ((size_t)1 << 30)) h->max_pending = v;
    }
    *out = h;
    ZB_CATCH
}

int zb_kmerize_set_baits(zb_kmerizer* h, const zb_set* baits) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    if (baits && baits->c != h->c) ZB_FAIL(ZB_E_ARG, "the bait set must live in the kmerizer's context");
    h->baits = baits;
    ZB_CATCH
}

int zb_kmerize_set_owners(zb_kmerizer* h, int nranks) {
    ZB_TRY
    if (!h || nranks < 0 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    h->owners = nranks;
    h->owner_cnt_valid = false;
    if (nranks > 1) {
        if (!h->owner_cnt.get()) h->owner_cnt.alloc(c, 64);
        ZB_CUDA(dev_memset(c, h->owner_cnt.get(), 0, 64 * 8));
        h->owner_cnt_valid = (h->pending_upper == 0);   // keys extracted before this call were not tallied
    }
    ZB_CATCH
}

int zb_kmerize_feed_dev(zb_kmerizer* h, const uint8_t* d_raw, size_t n, int is_fasta) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    ZB_CUDA(cudaSetDevice(h->c->device));
    EngineLock el(h->c, ENG_SM);
    feed_dev_impl(h, d_raw, n, is_fasta);
    ZB_CATCH
}

int zb_kmerize_feed(zb_kmerizer* h, const uint8_t* raw, size_t n, int is_fasta) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    if (n == 0) return ZB_OK;
    if (!raw) ZB_FAIL(ZB_E_ARG, "null input");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    DBuf<uint8_t> d(c, n + 16);
    {
        EngineLock el(c, ENG_H2D);
        Stage st(c, "h2d");
        copy_chunked(c, d.get(), raw, n, cudaMemcpyHostToDevice);
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    EngineLock el(c, ENG_SM);
    feed_dev_impl(h, d.get(), n, is_fasta);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_kmerize_feed_codes_dev(zb_kmerizer* h, const uint8_t* d_codes, size_t n, uint64_t n_records) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    if (n == 0) { h->n_records += n_records; return ZB_OK; }
    const size_t padded = div_up(n, EXTRACT_TILE) * EXTRACT_TILE + 32;
    DBuf<uint8_t> codes(c, 32 + padded);
    ZB_CUDA(dev_memset(c, codes.get(), 4, 32));
    ZB_CUDA(dev_copy(c, codes.get() + 32, d_codes, n));
    ZB_CUDA(dev_memset(c, codes.get() + 32 + n, 4, padded - n));
    extract_codes(h, codes.get() + 32, n);
    h->n_records += n_records;
    ZB_CATCH
}

int zb_kmerize_finish(zb_kmerizer* h, zb_set** result, uint64_t* n_records) {
    ZB_TRY
    if (!h || !result) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    EngineLock el(c, ENG_SM);
    flush_pending(h, true);
    h->pending.release();
    h->pending_cap = 0;
    compact_runs(h);
    if (!h->runs.empty()) {
        h->acc_k = std::move(h->runs[0].k);
        h->acc_c = std::move(h->runs[0].c);
        h->acc_n = h->runs[0].n;
        h->runs.clear();
    }
    const size_t n = h->acc_n;
    // both strands: mirror the canonical run and unite the two halves
    DBuf<uint64_t> rk(c, n), rk2(c, n);
    DBuf<uint32_t> rc(c, n), rc2(c, n);
    size_t nm;
    {
        Stage st(c, "mirror");
        nm = mirror_keys(c, h->k, h->acc_k.get(), h->acc_c.get(), n, rk.get(), rc.get());
    }
    zb_set* s = new_set(c, n + nm);
    bool fused = false;
    int which = 0;
    if (g_sort_count_mode == 0 && n + nm > 0) {
        Stage st(c, "mirror_merge");
        fused = merge_mirrored(c, h->acc_k.get(), h->acc_c.get(), n, rk.get(), rk2.get(), rc.get(), rc2.get(), nm, 2 * h->k,
                               s->k.get(), s->cnt.get(), &which);
    }
    if (!fused) {
        // skewed key space (or a cross-check route): sort the mirrored half completely, then merge path
        uint64_t* m0 = which ? rk2.get() : rk.get();
        uint64_t* m1 = which ? rk.get() : rk2.get();
        uint32_t* v0 = which ? rc2.get() : rc.get();
        uint32_t* v1 = which ? rc.get() : rc2.get();
        DBuf<uint64_t> mk(c, n);
        DBuf<uint32_t> mc(c, n);
        {
            // the mirrored keys are distinct, so "sort + count" is a sort of (key, count) pairs
            Stage st(c, "mirror_sort");
            const size_t nm2 = sort_count(c, m0, m1, v0, v1, nm, 2 * h->k, mk.get(), mc.get(), true);
            if (nm2 != nm) { delete s; ZB_FAIL(ZB_E_CUDA, "mirror: %zu distinct reverse complements of %zu keys", nm2, nm); }
        }
        Stage st_merge(c, "mirror_merge");
        merge_pairs(c, h->acc_k.get(), h->acc_c.get(), n, mk.get(), mc.get(), nm, s->k.get(), s->cnt.get());
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    rk.release(); rk2.release(); rc.release(); rc2.release();
    h->acc_k.release();
    h->acc_c.release();
    h->acc_n = 0;
    if (n_records) *n_records = h->n_records;
    *result = s;
    ZB_CATCH
}

int zb_kmerize_close(zb_kmerizer* h) {
    ZB_TRY
    if (h) {
        cudaSetDevice(h->c->device);
        delete h;
    }
    ZB_CATCH
}

int zb_kmerize_pending(zb_kmerizer* h, uint64_t* n_keys) {
    ZB_TRY
    if (!h || !n_keys) ZB_FAIL(ZB_E_ARG, "null argument");
    ZB_CUDA(cudaSetDevice(h->c->device));
    *n_keys = h->pending_upper ? read_pending_count(h) : 0;
    ZB_CATCH
}

int zb_kmerize_take_bucketed_dev(zb_kmerizer* h, int nranks, uint64_t* d_keys, uint64_t* bucket_counts) {
    ZB_TRY
    if (!h || !bucket_counts || nranks < 1 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    const size_t n = h->pending_upper ? read_pending_count(h) : 0;
    DBuf<unsigned long long> cnt(c, 128);
    ZB_CUDA(dev_memset(c, cnt.get(), 0, 128 * 8));
    bucket_count(c, h->pending.get(), n, nranks, cnt.get());
    std::vector<unsigned long long> hc(64, 0), start(64, 0);
    ZB_CUDA(cudaMemcpyAsync(hc.data(), cnt.get(), 64 * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long s = 0;
    for (int r = 0; r < nranks; r++) { start[r] = s; s += hc[r]; bucket_counts[r] = hc[r]; }
    ZB_CUDA(cudaMemcpyAsync(cnt.get() + 64, start.data(), 64 * 8, cudaMemcpyHostToDevice, c->stream));
    if (n && !d_keys) ZB_FAIL(ZB_E_ARG, "null d_keys");
    bucket_scatter(c, h->pending.get(), n, nranks, cnt.get() + 64, d_keys);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    pending_consumed(h);
    ZB_CATCH
}

int zb_kmerize_bucket_counts(zb_kmerizer* h, int nranks, uint64_t* bucket_counts) {
    ZB_TRY
    if (!h || !bucket_counts || nranks < 1 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    if (h->owners == nranks && h->owner_cnt_valid) {   // tallied while the keys were extracted: no pass over them
        ZB_CUDA(read_back_big(c, h->owner_cnt.get(), 64 * 8));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < nranks; r++) bucket_counts[r] = reinterpret_cast<const uint64_t*>(c->h_big)[r];
        return ZB_OK;
    }
    const size_t n = h->pending_upper ? read_pending_count(h) : 0;
    DBuf<unsigned long long> cnt(c, 64);
    ZB_CUDA(dev_memset(c, cnt.get(), 0, 64 * 8));
    bucket_count(c, h->pending.get(), n, nranks, cnt.get());
    ZB_CUDA(read_back_big(c, cnt.get(), 64 * 8));      // by a kernel: the copy engines may be busy with other threads' bulk copies
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < nranks; r++) bucket_counts[r] = reinterpret_cast<const uint64_t*>(c->h_big)[r];
    ZB_CATCH
}

int zb_kmerize_route_p2p(zb_kmerizer* h, int nranks, uint64_t* const* d_dst) {
    ZB_TRY
    if (!h || !d_dst || nranks < 1 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    const size_t n = h->pending_upper ? read_pending_count(h) : 0;
    PeerPtrs pp;
    for (int r = 0; r < 64; r++) pp.p[r] = (r < nranks) ? d_dst[r] : nullptr;
    DBuf<unsigned long long> cur(c, 64);
    ZB_CUDA(dev_memset(c, cur.get(), 0, 64 * 8));
    {
        Stage st(c, "route_p2p");
        route_p2p(c, h->pending.get(), n, nranks, pp, cur.get());
    }
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // every store, local or over NVLink, has been issued and completed
    pending_consumed(h);
    ZB_CATCH
}

// The same in two halves: begin launches the routing kernel on the context's SECOND stream and returns at once, so the
// caller can count what the previous exchange delivered (zb_kmerize_flush: sort + count on the first stream) while this
// batch's keys travel over NVLink; end waits for the kernel.  Nothing may be fed between begin and end.
int zb_kmerize_route_p2p_begin(zb_kmerizer* h, int nranks, uint64_t* const* d_dst) {
    ZB_TRY
    if (!h || !d_dst || nranks < 1 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    if (h->routing) ZB_FAIL(ZB_E_ARG, "a routing kernel is already in flight");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    const size_t n = h->pending_upper ? read_pending_count(h) : 0;
    if (!c->side) {
        ZB_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        ZB_CUDA(cudaEventCreateWithFlags(&c->side_ev, cudaEventDisableTiming));
    }
    PeerPtrs pp;
    for (int r = 0; r < 64; r++) pp.p[r] = (r < nranks) ? d_dst[r] : nullptr;
    h->route_cur.alloc(c, 64);
    ZB_CUDA(dev_memset(c, h->route_cur.get(), 0, 64 * 8));
    ZB_CUDA(cudaEventRecord(c->side_ev, c->stream));          // the keys (and the zeroed cursors) are ready
    ZB_CUDA(cudaStreamWaitEvent(c->side, c->side_ev, 0));
    route_p2p(c, h->pending.get(), n, nranks, pp, h->route_cur.get(), nullptr, 0, nullptr, c->side);
    h->routing = true;
    pending_consumed(h);       // the pending list is spoken for; its buffer is reused only after _end
    ZB_CATCH
}

int zb_kmerize_route_p2p_end(zb_kmerizer* h) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    if (!h->routing) return ZB_OK;
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    ZB_CUDA(cudaStreamSynchronize(c->side));   // every store, local or over NVLink, has completed
    h->routing = false;
    h->route_cur.release();
    ZB_CATCH
}

int zb_kmerize_route_p2p_reserve(zb_kmerizer* h, int nranks, uint64_t* const* d_dst, uint64_t* const* d_cursor,
                                 uint64_t capacity_keys, uint64_t* sent_counts) {
    ZB_TRY
    if (!h || !d_dst || !d_cursor || nranks < 1 || nranks > 64) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    const size_t n = h->pending_upper ? read_pending_count(h) : 0;
    PeerPtrs pp, pc;
    for (int r = 0; r < 64; r++) {
        pp.p[r] = (r < nranks) ? d_dst[r] : nullptr;
        pc.p[r] = (r < nranks) ? d_cursor[r] : nullptr;
    }
    DBuf<unsigned long long> cur(c, 66);      // [0, 64): keys sent per owner; [64]: overflow flag
    ZB_CUDA(dev_memset(c, cur.get(), 0, 66 * 8));
    {
        Stage st(c, "route_p2p");
        route_p2p(c, h->pending.get(), n, nranks, pp, cur.get(), &pc, capacity_keys, reinterpret_cast<unsigned int*>(cur.get() + 64));
    }
    std::vector<unsigned long long> hc(66, 0);
    ZB_CUDA(read_back_big(c, cur.get(), 66 * 8));
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // every store and reservation, local or over NVLink, has completed
    memcpy(hc.data(), c->h_big, 66 * 8);
    pending_consumed(h);
    if (sent_counts)
        for (int r = 0; r < nranks; r++) sent_counts[r] = hc[r];
    if (hc[64] & 0xffffffffull) ZB_FAIL(ZB_E_RANGE, "route_p2p: a receive buffer of %llu keys overflowed", (unsigned long long)capacity_keys);
    ZB_CATCH
}

int zb_peer_enable(int device, int peer) {
    ZB_TRY
    ctx_for(device);
    ctx_for(peer);
    if (device == peer) return ZB_OK;
    ZB_CUDA(cudaSetDevice(device));
    int can = 0;
    ZB_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) ZB_FAIL(ZB_E_CUDA, "device %d cannot map the memory of device %d", device, peer);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else ZB_CUDA(e);
    ZB_CATCH
}

int zb_ipc_alloc(int device, size_t bytes, void** d_ptr, uint8_t handle[64]) {
    ZB_TRY
    if (!d_ptr || !handle) ZB_FAIL(ZB_E_ARG, "null argument");
    ctx_for(device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    ZB_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        ZB_CUDA(e);
    }
    memcpy(handle, &hd, 64);
    *d_ptr = p;
    ZB_CATCH
}

int zb_ipc_open(int device, const uint8_t handle[64], void** d_ptr) {
    ZB_TRY
    if (!d_ptr || !handle) ZB_FAIL(ZB_E_ARG, "null argument");
    ctx_for(device);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    ZB_CUDA(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    ZB_CATCH
}

int zb_ipc_close(int device, void* d_ptr) {
    ZB_TRY
    ctx_for(device);
    if (d_ptr) ZB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    ZB_CATCH
}

int zb_ipc_free(int device, void* d_ptr) {
    ZB_TRY
    ctx_for(device);
    if (d_ptr) ZB_CUDA(cudaFree(d_ptr));
    ZB_CATCH
}

int zb_kmerize_adopt_canonical_dev(zb_kmerizer* h, uint64_t* d_keys, size_t n) {
    ZB_TRY
    if (!h || (n && !d_keys)) ZB_FAIL(ZB_E_ARG, "null argument");
    ZB_CUDA(cudaSetDevice(h->c->device));
    if (n >= ((size_t)1 << 30)) return zb_kmerize_add_canonical_dev(h, d_keys, n);   // too large for one sort: copy in pieces
    flush_pending(h);          // an earlier adopted array must be consumed before its owner re-uses it
    h->adopted = d_keys;
    h->adopted_n = n;
    ZB_CATCH
}

int zb_kmerize_flush(zb_kmerizer* h) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    ZB_CUDA(cudaSetDevice(h->c->device));
    EngineLock el(h->c, ENG_SM);
    flush_pending(h);
    ZB_CATCH
}

int zb_kmerize_add_canonical_dev(zb_kmerizer* h, const uint64_t* d_keys, size_t n) {
    ZB_TRY
    if (!h) ZB_FAIL(ZB_E_ARG, "null handle");
    Ctx* c = h->c;
    ZB_CUDA(cudaSetDevice(c->device));
    size_t off = 0;
    while (off < n) {
        size_t have = h->pending_upper ? read_pending_count(h) : 0;
        if (have >= h->max_pending) { flush_pending(h); have = 0; }
        const size_t len = std::min(n - off, h->max_pending - have);
        ensure_pending(h, have + len);
        h->owner_cnt_valid = false;   // these keys were not tallied
        h->pre_on = false;
        ZB_CUDA(dev_copy(c, h->pending.get() + have, d_keys + off, len * 8));
        const unsigned long long nc = have + len;
        c->h_scalars[8] = nc;
        ZB_CUDA(cudaMemcpyAsync(h->d_count.get(), c->h_scalars + 8, 8, cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        h->pending_upper = (size_t)nc;
        off += len;
    }
    ZB_CATCH
}

// ------------------------------------------------------------------------------- sets
int zb_set_from_host(int device, const uint64_t* kmers, const uint32_t* counts, size_t n, zb_set** out) {
    ZB_TRY
    if (!out || (n && !kmers)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    zb_set* s = new_set(c, n);
    if (n) {
        ZB_CUDA(cudaMemcpyAsync(s->k.get(), kmers, n * 8, cudaMemcpyHostToDevice, c->stream));
        if (counts) {
            ZB_CUDA(cudaMemcpyAsync(s->cnt.get(), counts, n * 4, cudaMemcpyHostToDevice, c->stream));
        } else {
            fill_u32(c, s->cnt.get(), n, 1u);
        }
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    *out = s;
    ZB_CATCH
}

int zb_set_from_device(int device, const uint64_t* d_kmers, const uint32_t* d_counts, size_t n, zb_set** out) {
    ZB_TRY
    if (!out || (n && !d_kmers)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    zb_set* s = new_set(c, n);
    if (n) {
        ZB_CUDA(dev_copy(c, s->k.get(), d_kmers, n * 8));
        if (d_counts) ZB_CUDA(dev_copy(c, s->cnt.get(), d_counts, n * 4));
        else fill_u32(c, s->cnt.get(), n, 1u);
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    *out = s;
    ZB_CATCH
}

int zb_set_lower_bound(const zb_set* s, const uint64_t* probes, size_t m, uint64_t* idx) {
    ZB_TRY
    if (!s || (m && (!probes || !idx))) ZB_FAIL(ZB_E_ARG, "null argument");
    ZB_CUDA(cudaSetDevice(s->c->device));
    lower_bound(s->c, s->k.get(), s->n, probes, m, idx);
    ZB_CATCH
}

int zb_set_slice(const zb_set* s, size_t begin, size_t end, zb_set** out) {
    ZB_TRY
    if (!s || !out || begin > end || end > s->n) ZB_FAIL(ZB_E_ARG, "bad slice");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_set* r = new_set(c, end - begin);
    if (end > begin) {
        ZB_CUDA(dev_copy(c, r->k.get(), s->k.get() + begin, (end - begin) * 8));
        ZB_CUDA(dev_copy(c, r->cnt.get(), s->cnt.get() + begin, (end - begin) * 4));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        zb_set_carry_wide(s, r);
    }
    *out = r;
    ZB_CATCH
}

int zb_set_size(const zb_set* s, size_t* n) {
    if (!s || !n) { zb::set_error("null argument"); return ZB_E_ARG; }
    *n = s->n;
    return ZB_OK;
}

int zb_set_fetch(const zb_set* s, uint64_t* kmers, uint32_t* counts) {
    ZB_TRY
    if (!s) ZB_FAIL(ZB_E_ARG, "null set");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    EngineLock el(c, ENG_D2H);
    if (s->n && kmers) copy_chunked(c, kmers, s->k.get(), s->n * 8, cudaMemcpyDeviceToHost);
    if (s->n && counts) copy_chunked(c, counts, s->cnt.get(), s->n * 4, cudaMemcpyDeviceToHost);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_set_dev_ptrs(const zb_set* s, const uint64_t** d_kmers, const uint32_t** d_counts) {
    if (!s) { zb::set_error("null set"); return ZB_E_ARG; }
    if (d_kmers) *d_kmers = s->k.get();
    if (d_counts) *d_counts = s->cnt.get();
    return ZB_OK;
}

int zb_set_free(zb_set* s) {
    ZB_TRY
    if (s) {
        cudaSetDevice(s->c->device);
        delete s;
    }
    ZB_CATCH
}

int zb_set_stats(const zb_set* s, uint64_t acgt_weighted[4], uint64_t acgt_plain[4], uint64_t* total_count,
                 uint64_t* hist_vals, uint64_t* hist_freqs, size_t hist_cap, size_t* n_hist) {
    ZB_TRY
    if (!s) ZB_FAIL(ZB_E_ARG, "null set");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    uint64_t aw[4], ap[4], tot;
    std::vector<std::pair<uint64_t, uint64_t>> hist;
    std::vector<uint64_t> first;
    {
        EngineLock el(c, ENG_SM);
        set_stats(c, s->k.get(), s->cnt.get(), s->n, aw, ap, &tot, &hist, &first);
    }
    if (s->wide.get()) {
        // the entries the kernel saw as 2^32-1 are exactly the listed ones: their bin goes, their true values come in
        struct E { uint64_t first, val, freq; };
        std::vector<E> es;
        for (size_t i = 0; i < hist.size(); i++)
            if (hist[i].first != 0xffffffffull) es.push_back(E{first[i], hist[i].first, hist[i].second});
        std::map<uint64_t, E> add;
        for (size_t i = 0; i < s->exc_idx.size(); i++) {
            const uint64_t v = s->exc_val[i];
            auto it = add.find(v);
            if (it == add.end()) add[v] = E{s->exc_idx[i], v, 1};   // exc_idx ascends: the first one seen is the first position
            else it->second.freq++;
            aw[s->exc_key[i] & 3] += v - 0xffffffffull;
            tot += v - 0xffffffffull;
        }
        for (auto& kv : add) es.push_back(kv.second);
        std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.first < b.first; });
        hist.clear();
        for (auto& e : es) hist.push_back({e.val, e.freq});
    }
    for (int q = 0; q < 4; q++) {
        if (acgt_weighted) acgt_weighted[q] = aw[q];
        if (acgt_plain) acgt_plain[q] = ap[q];
    }
    if (total_count) *total_count = tot;
    if (n_hist) *n_hist = hist.size();
    if (hist_cap) {  // fills min(hist_cap, *n_hist) entries; the caller retries when *n_hist > hist_cap
        for (size_t i = 0; i < hist.size() && i < hist_cap; i++) {
            if (hist_vals) hist_vals[i] = hist[i].first;
            if (hist_freqs) hist_freqs[i] = hist[i].second;
        }
    }
    ZB_CATCH
}

// one input of a merge: a set's arrays, or the same keys with other counts (the 16-bit planes of the wide route)
struct MergeIn {
    struct K { const uint64_t* p; const uint64_t* get() const { return p; } } k;
    struct C { const uint32_t* p; const uint32_t* get() const { return p; } } cnt;
    size_t n;
    Ctx* c;
};

static int merge_core(int nsets, const MergeIn* const* sets, zb_set** out) {
    ZB_TRY
    Ctx* c = sets[0]->c;
    size_t total = 0;
    for (int i = 0; i < nsets; i++) total += sets[i]->n;
    const char* mm = getenv("ZB_MERGE");
    const bool want_tree = mm && !strcmp(mm, "tree");
    const bool want_sort = mm && !strcmp(mm, "sort");
    if (nsets >= 3 && total > 0 && !want_tree) {
        // key range: the sets are sorted, so the largest key is the largest last element
        std::vector<uint64_t> last(nsets, 0);
        for (int i = 0; i < nsets; i++)
            if (sets[i]->n) ZB_CUDA(cudaMemcpyAsync(&last[i], sets[i]->k.get() + sets[i]->n - 1, 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        uint64_t maxkey = 0;
        for (int i = 0; i < nsets; i++) maxkey = std::max(maxkey, last[i]);
        const int key_bits = maxkey ? 64 - __builtin_clzll(maxkey) : 1;
        if (!want_sort) {
            // Many inputs: ONE pass over them, bucket by bucket of the key space in shared memory (nwaymerge.cu)
            std::vector<const uint64_t*> ks(nsets);
            std::vector<const uint32_t*> cs(nsets);
            std::vector<size_t> ns(nsets);
            for (int i = 0; i < nsets; i++) { ks[i] = sets[i]->k.get(); cs[i] = sets[i]->cnt.get(); ns[i] = sets[i]->n; }
            zb_set* s = new zb_set();
            s->c = c;
            s->n = 0;
            bool done = false;
            try {
                Stage st(c, "merge_nway");
                // slabs of at most 2^27 input entries: the staging stays at 1.6 GB whatever the inputs (64 bacterial
                // sets = 637 M entries: one 7.6 GB allocation -- 80 - 200 ms of cudaMalloc in a fresh `zot merge` process)
                done = merge_nway_slabs(c, ks, cs, ns, key_bits, (size_t)1 << 27, &s->k, &s->cnt, &s->n);
            } catch (...) {
                delete s;
                throw;
            }
            if (done) {
                ZB_CUDA(cudaStreamSynchronize(c->stream));
                *out = s;
                return ZB_OK;
            }
            delete s;
        }
        if (nsets < 4 || total >= ((size_t)1 << 30)) goto tree;
        // Fallback (skewed key space, > 1024 inputs): ONE weighted sort + count of the concatenation (segsort.cu)
        DBuf<uint64_t> k0(c, total), k1(c, total);
        DBuf<uint32_t> v0(c, total), v1(c, total);
        size_t off = 0;
        for (int i = 0; i < nsets; i++) {
            if (!sets[i]->n) continue;
            ZB_CUDA(dev_copy(c, k0.get() + off, sets[i]->k.get(), sets[i]->n * 8));
            ZB_CUDA(dev_copy(c, v0.get() + off, sets[i]->cnt.get(), sets[i]->n * 4));
            off += sets[i]->n;
        }
        zb_set* s = new_set(c, total);
        try {
            Stage st(c, "merge_sort");
            s->n = sort_count(c, k0.get(), k1.get(), v0.get(), v1.get(), total, key_bits, s->k.get(), s->cnt.get());
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
        return ZB_OK;
    }
tree:
    // pairwise tree; every level is merge-path + reduce-by-key (counts summed)
    struct Run { const uint64_t* k; const uint32_t* c; size_t n; DBuf<uint64_t> ok; DBuf<uint32_t> oc; };
    std::vector<Run> cur(nsets);
    for (int i = 0; i < nsets; i++) { cur[i].k = sets[i]->k.get(); cur[i].c = sets[i]->cnt.get(); cur[i].n = sets[i]->n; }
    while (cur.size() > 1) {
        std::vector<Run> nxt((cur.size() + 1) / 2);
        for (size_t i = 0; i + 1 < cur.size(); i += 2) {
            Run& a = cur[i];
            Run& b = cur[i + 1];
            const size_t tot = a.n + b.n;
            DBuf<uint64_t> mk(c, tot);
            DBuf<uint32_t> mc(c, tot);
            merge_pairs(c, a.k, a.c, a.n, b.k, b.c, b.n, mk.get(), mc.get());
            Run& r = nxt[i / 2];
            r.ok.alloc(c, tot);
            r.oc.alloc(c, tot);
            r.n = reduce_by_key(c, mk.get(), mc.get(), tot, r.ok.get(), r.oc.get());
            r.k = r.ok.get();
            r.c = r.oc.get();
            a.ok.release(); a.oc.release(); b.ok.release(); b.oc.release();
        }
        if (cur.size() & 1) nxt.back() = std::move(cur.back());
        cur = std::move(nxt);
    }
    zb_set* s = new_set(c, cur[0].n);
    if (cur[0].n) {
        ZB_CUDA(dev_copy(c, s->k.get(), cur[0].k, cur[0].n * 8));
        ZB_CUDA(dev_copy(c, s->cnt.get(), cur[0].c, cur[0].n * 4));
    }
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    *out = s;
    ZB_CATCH
}

// ---- counts beyond 2^32-1 ------------------------------------------------------------------------------------------
// The reference sums Python ints (merge.py:56, :145-146) and codec64 carries up to 60 bits, so a merge of very deep
// sets may hold counts that no u32 array can.  Every merge kernel reports such a sum (ZB_E_RANGE); zb_merge then
// merges the inputs twice more with the low and the high 16 bits of every count as the counts -- each of those sums
// fits 32 bits for up to 65,536 inputs -- and adds the two results up in 64 bits.
// plane p of a count: bits [16 p, 16 p + 16), from the u32 counts or from the u64 counts of a wide input
__global__ void __launch_bounds__(256) count_plane_kernel(const uint32_t* __restrict__ c, const uint64_t* __restrict__ w, size_t n, int shift,
                                                          uint32_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = w ? (uint32_t)(w[i] >> shift) & 0xffffu : (shift < 32 ? (c[i] >> shift) & 0xffffu : 0u);
}
__global__ void __launch_bounds__(256) wide_add_plane_kernel(uint64_t* __restrict__ wide, const uint32_t* __restrict__ plane, size_t n, int shift,
                                                             int first) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        wide[i] = (first ? 0ull : wide[i]) + ((uint64_t)plane[i] << shift);
}
// cnt = min(wide, 2^32-1); the entries that reach 2^32-1 are listed
__global__ void __launch_bounds__(256) wide_sat_kernel(const uint64_t* __restrict__ wide, size_t n, uint32_t* __restrict__ cnt,
                                                       unsigned long long* __restrict__ n_exc, uint64_t* __restrict__ exc_idx, uint64_t exc_cap) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const uint64_t v = wide[i];
        cnt[i] = v >= 0xffffffffull ? 0xffffffffu : (uint32_t)v;
        if (v >= 0xffffffffull) {
            const unsigned long long s = atomicAdd(n_exc, 1ull);
            if (s < exc_cap) exc_idx[s] = i;
        }
    }
}
__global__ void __launch_bounds__(256) exc_gather_kernel(const uint64_t* __restrict__ idx, size_t m, const uint64_t* __restrict__ k,
                                                         const uint64_t* __restrict__ wide, uint64_t* __restrict__ out /*[2m]*/) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) { out[i] = k[idx[i]]; out[m + i] = wide[idx[i]]; }
}

}  // extern "C"

// r->k (n keys) and r->wide (their counts as u64) are in place: the saturated u32 counts and the list of the entries
// that reach 2^32-1 follow.  Also used by the decoder for a file that holds such counts (codec.cu).
void zb_set_finish_wide(zb_set* r) {
    Ctx* c = r->c;
    if (!r->cnt.get() || r->cnt.n < r->n) r->cnt.alloc(c, r->n);
    size_t cap = 1 << 16, m = 0;
    DBuf<unsigned long long> nexc(c, 2);
    DBuf<uint64_t> eidx;
    for (int attempt = 0; attempt < 2; attempt++) {
        eidx.alloc(c, cap);
        ZB_CUDA(dev_memset(c, nexc.get(), 0, 16));
        const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(std::max<size_t>(r->n, 1), 256));
        wide_sat_kernel<<<blocks, 256, 0, c->stream>>>(r->wide.get(), r->n, r->cnt.get(), nexc.get(), eidx.get(), cap);
        ZB_LAUNCH_CHECK(c);
        ZB_CUDA(read_back(c, nexc.get(), 8));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        m = (size_t)c->h_scalars[0];
        if (m <= cap) break;
        cap = m;
    }
    r->exc_idx.resize(m);
    r->exc_key.resize(m);
    r->exc_val.resize(m);
    if (m) {
        ZB_CUDA(cudaMemcpy(r->exc_idx.data(), eidx.get(), m * 8, cudaMemcpyDeviceToHost));
        std::sort(r->exc_idx.begin(), r->exc_idx.end());
        ZB_CUDA(cudaMemcpy(eidx.get(), r->exc_idx.data(), m * 8, cudaMemcpyHostToDevice));
        DBuf<uint64_t> kv(c, 2 * m);
        exc_gather_kernel<<<(unsigned)div_up(m, 256), 256, 0, c->stream>>>(eidx.get(), m, r->k.get(), r->wide.get(), kv.get());
        ZB_LAUNCH_CHECK(c);
        std::vector<uint64_t> h(2 * m);
        ZB_CUDA(cudaMemcpyAsync(h.data(), kv.get(), 2 * m * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < m; i++) { r->exc_key[i] = h[i]; r->exc_val[i] = h[m + i]; }
    } else {
        r->wide.release();   // every count fits 32 bits after all: an ordinary set
    }
}

__global__ void __launch_bounds__(256) widen_kernel(const uint32_t* __restrict__ c, size_t n, uint64_t* __restrict__ w) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) w[i] = c[i];
}
__global__ void __launch_bounds__(256) gather_keys_kernel(const uint64_t* __restrict__ idx, size_t m, const uint64_t* __restrict__ k, size_t n,
                                                          uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) out[i] = idx[i] < n ? k[idx[i]] : ~0ull;
}
__global__ void __launch_bounds__(256) patch_wide_kernel(uint64_t* __restrict__ w, const uint64_t* __restrict__ idx_val /*[2m]*/, size_t m) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) w[idx_val[i]] = idx_val[m + i];
}

// `out` holds a subset of the entries of `in` (trim, sample, restrict, slice: the counts travel unchanged).  If `in` is
// wide, the entries of `out` whose u32 count is saturated get their true counts back: the few of them are looked up by
// key in the list of `in`.  lo / hi (hi = 0: none): the count range a trim kept -- a listed entry that the saturated
// count let through although its true count is outside (thresholds beyond 2^32-1) cannot be taken out again here.
static void carry_wide(const zb_set* in, zb_set* out, uint64_t lo = 0, uint64_t hi = 0) {
    if (!in->wide.get()) return;
    Ctx* c = in->c;
    const size_t m = in->exc_key.size();
    for (size_t i = 0; i < m; i++) {
        const uint64_t v = in->exc_val[i];
        const bool want = v >= lo && (hi == 0 || v <= hi);
        const bool got = 0xffffffffull >= lo && (hi == 0 || 0xffffffffull <= hi);
        if (want != got) ZB_FAIL(ZB_E_RANGE, "trim: count thresholds beyond 2^32-1 on a set that holds such counts");
    }
    if (out->n == 0 || m == 0) return;
    std::vector<uint64_t> idx(m), keys(m);
    lower_bound(c, out->k.get(), out->n, in->exc_key.data(), m, idx.data());
    DBuf<uint64_t> d(c, 2 * m);
    ZB_CUDA(cudaMemcpyAsync(d.get(), idx.data(), m * 8, cudaMemcpyHostToDevice, c->stream));
    gather_keys_kernel<<<(unsigned)div_up(m, 256), 256, 0, c->stream>>>(d.get(), m, out->k.get(), out->n, d.get() + m);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(cudaMemcpyAsync(keys.data(), d.get() + m, m * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> iv;
    for (size_t i = 0; i < m; i++) {
        if (idx[i] < out->n && keys[i] == in->exc_key[i]) {
            out->exc_idx.push_back(idx[i]);
            out->exc_key.push_back(in->exc_key[i]);
            out->exc_val.push_back(in->exc_val[i]);
        }
    }
    const size_t kept = out->exc_idx.size();
    if (kept == 0) return;
    out->wide.alloc(c, out->n);
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(out->n, 256));
    widen_kernel<<<blocks, 256, 0, c->stream>>>(out->cnt.get(), out->n, out->wide.get());
    ZB_LAUNCH_CHECK(c);
    iv.resize(2 * kept);
    for (size_t i = 0; i < kept; i++) { iv[i] = out->exc_idx[i]; iv[kept + i] = out->exc_val[i]; }
    DBuf<uint64_t> div_(c, 2 * kept);
    ZB_CUDA(cudaMemcpyAsync(div_.get(), iv.data(), 2 * kept * 8, cudaMemcpyHostToDevice, c->stream));
    patch_wide_kernel<<<(unsigned)div_up(kept, 256), 256, 0, c->stream>>>(out->wide.get(), div_.get(), kept);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
}

static void zb_set_carry_wide(const zb_set* in, zb_set* out) { carry_wide(in, out); }

extern "C" {

static int merge_wide(int nsets, zb_set* const* sets, zb_set** out) {
    ZB_TRY
    Ctx* c = sets[0]->c;
    if (nsets > 65536) ZB_FAIL(ZB_E_RANGE, "k-mer count exceeds 2^32-1 in a merge of more than 65,536 sets");
    size_t total = 0;
    bool wide_in = false;
    for (int i = 0; i < nsets; i++) { total += sets[i]->n; wide_in = wide_in || sets[i]->wide.get() != nullptr; }
    const int planes = wide_in ? 4 : 2;     // 16-bit planes of the inputs' counts (codec64 carries at most 60 bits)
    DBuf<uint32_t> plane(c, total);
    std::vector<MergeIn> in(nsets);
    std::vector<const MergeIn*> ptr(nsets);
    zb_set* r = nullptr;
    zb_set* part = nullptr;
    try {
        for (int p = 0; p < planes; p++) {
            size_t off = 0;
            for (int i = 0; i < nsets; i++) {
                const size_t n = sets[i]->n;
                if (n) {
                    const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(n, 256));
                    count_plane_kernel<<<blocks, 256, 0, c->stream>>>(sets[i]->cnt.get(), sets[i]->wide.get(), n, 16 * p, plane.get() + off);
                    ZB_LAUNCH_CHECK(c);
                }
                in[i].k.p = sets[i]->k.get();
                in[i].cnt.p = plane.get() + off;
                in[i].n = n;
                in[i].c = c;
                ptr[i] = &in[i];
                off += n;
            }
            if (int rc = merge_core(nsets, ptr.data(), &part)) throw zb::Fail{rc};
            if (p == 0) {
                r = part;
                part = nullptr;
                r->wide.alloc(c, r->n);
            } else if (part->n != r->n) {
                ZB_FAIL(ZB_E_CUDA, "merge: the count planes disagree (%zu vs %zu keys)", r->n, part->n);
            }
            const zb_set* src = p == 0 ? r : part;
            if (r->n) {
                const unsigned blocks = (unsigned)std::min<size_t>((size_t)c->sm_count * 8, div_up(r->n, 256));
                wide_add_plane_kernel<<<blocks, 256, 0, c->stream>>>(r->wide.get(), src->cnt.get(), r->n, 16 * p, p == 0 ? 1 : 0);
                ZB_LAUNCH_CHECK(c);
            }
            if (part) {
                ZB_CUDA(cudaStreamSynchronize(c->stream));
                zb_set_free(part);
                part = nullptr;
            }
        }
        zb_set_finish_wide(r);
        *out = r;
    } catch (...) {
        if (r) zb_set_free(r);
        if (part) zb_set_free(part);
        throw;
    }
    ZB_CATCH
}

int zb_merge(int nsets, zb_set* const* sets, zb_set** out) {
    ZB_TRY
    if (nsets < 1 || !sets || !out) ZB_FAIL(ZB_E_ARG, "bad argument");
    for (int i = 0; i < nsets; i++)
        if (!sets[i] || sets[i]->c != sets[0]->c) ZB_FAIL(ZB_E_ARG, "sets must live on one device");
    Ctx* c = sets[0]->c;
    ZB_CUDA(cudaSetDevice(c->device));
    std::vector<MergeIn> in(nsets);
    std::vector<const MergeIn*> ptr(nsets);
    bool wide_in = false;
    for (int i = 0; i < nsets; i++) {
        wide_in = wide_in || sets[i]->wide.get() != nullptr;
        in[i].k.p = sets[i]->k.get();
        in[i].cnt.p = sets[i]->cnt.get();
        in[i].n = sets[i]->n;
        in[i].c = c;
        ptr[i] = &in[i];
    }
    if (wide_in) return merge_wide(nsets, sets, out);   // the saturated u32 counts of a wide input must not be added up
    const int rc = merge_core(nsets, ptr.data(), out);
    if (rc != ZB_E_RANGE) return rc;
    return merge_wide(nsets, sets, out);
    ZB_CATCH
}

int zb_set_is_wide(const zb_set* s, int* wide) {
    if (!s || !wide) { zb::set_error("null argument"); return ZB_E_ARG; }
    *wide = s->wide.get() ? 1 : 0;
    return ZB_OK;
}

int zb_set_fetch_counts64(const zb_set* s, uint64_t* counts) {
    ZB_TRY
    if (!s || (s->n && !counts)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    if (s->n == 0) return ZB_OK;
    EngineLock el(c, ENG_D2H);
    if (s->wide.get()) {
        copy_chunked(c, counts, s->wide.get(), s->n * 8, cudaMemcpyDeviceToHost);
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        std::vector<uint32_t> t(s->n);
        copy_chunked(c, t.data(), s->cnt.get(), s->n * 4, cudaMemcpyDeviceToHost);
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < s->n; i++) counts[i] = t[i];
    }
    ZB_CATCH
}

int zb_trim(const zb_set* s, uint64_t cmin, uint64_t cmax, zb_set** out) {
    ZB_TRY
    if (!s || !out) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_set* r = new_set(c, s->n);
    try {
        EngineLock el(c, ENG_SM);
        Stage st(c, "trim");
        r->n = trim_pairs(c, s->k.get(), s->cnt.get(), s->n, cmin, cmax, r->k.get(), r->cnt.get());
        carry_wide(s, r, cmin, cmax);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    ZB_CATCH
}

int zb_sample(const zb_set* s, int mode, uint64_t seed, double p, zb_set** out) {
    ZB_TRY
    if (!s || !out || mode < 0 || mode > 1) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_set* r = new_set(c, s->n);
    try {
        Stage st(c, "sample");
        r->n = sample_pairs(c, s->k.get(), s->cnt.get(), s->n, mode, seed, p, r->k.get(), r->cnt.get());
        carry_wide(s, r);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    ZB_CATCH
}

int zb_restrict(const zb_set* s, const zb_set* ref, zb_set** out) {
    ZB_TRY
    if (!s || !ref || !out) ZB_FAIL(ZB_E_ARG, "null argument");
    if (s->c != ref->c) ZB_FAIL(ZB_E_ARG, "sets must live on one device");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_set* r = new_set(c, s->n);
    try {
        Stage st(c, "restrict");
        r->n = restrict_pairs(c, s->k.get(), s->cnt.get(), s->n, ref->k.get(), ref->n, r->k.get(), r->cnt.get());
        carry_wide(s, r);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    ZB_CATCH
}

int zb_project(const zb_set* s, int shift_bits, zb_set** out) {
    ZB_TRY
    if (!s || !out || shift_bits < 0 || shift_bits > 63) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_set* r = new zb_set();
    r->c = c;
    r->k.alloc(c, s->n);
    r->n = project_keys(c, s->k.get(), s->n, shift_bits, r->k.get());
    *out = r;
    ZB_CATCH
}

int zb_pairs_abc(int nsets, zb_set* const* sets, const uint32_t* I, const uint32_t* J, size_t npairs, uint64_t* abc) {
    ZB_TRY
    if (nsets < 1 || !sets || (npairs && (!I || !J || !abc))) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = sets[0]->c;
    ZB_CUDA(cudaSetDevice(c->device));
    std::vector<SetRef> refs(nsets);
    for (int i = 0; i < nsets; i++) {
        if (!sets[i] || sets[i]->c != c) ZB_FAIL(ZB_E_ARG, "sets must live on one device");
        refs[i].k = sets[i]->k.get();
        refs[i].n = sets[i]->n;
    }
    pairs_abc_host(c, refs, I, J, npairs, abc);
    ZB_CATCH
}

int zb_allpairs_tiles(int nsets, uint64_t* n_tiles) {
    if (nsets < 0 || !n_tiles) { zb::set_error("bad argument"); return ZB_E_ARG; }
    *n_tiles = allpairs_tiles(nsets);
    return ZB_OK;
}

int zb_allpairs_abc(int nsets, zb_set* const* sets, uint64_t tile_begin, uint64_t tile_end, uint64_t* abc) {
    return zb_allpairs_abc_strided(nsets, sets, tile_begin, tile_end, 1, abc);
}

int zb_allpairs_abc_strided(int nsets, zb_set* const* sets, uint64_t unit_begin, uint64_t unit_end, uint64_t unit_stride,
                            uint64_t* abc) {
    ZB_TRY
    const uint64_t tile_begin = unit_begin, tile_end = unit_end;
    if (unit_stride < 1) ZB_FAIL(ZB_E_ARG, "unit_stride must be >= 1");
    if (nsets < 1 || !sets || (nsets > 1 && !abc)) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = sets[0]->c;
    ZB_CUDA(cudaSetDevice(c->device));
    std::vector<SetRef> refs(nsets);
    for (int i = 0; i < nsets; i++) {
        if (!sets[i] || sets[i]->c != c) ZB_FAIL(ZB_E_ARG, "sets must live on one device");
        refs[i].k = sets[i]->k.get();
        refs[i].n = sets[i]->n;
    }
    allpairs_abc(c, refs, tile_begin, tile_end, unit_stride, abc);
    ZB_CATCH
}

}  // extern "C"

namespace zb {
// (|X n Y|, |X \ Y|, |Y \ X|) for a list of pairs, one merge-path pass per pair (setops.cu pairs_abc_kernel)
void pairs_abc_host(Ctx* c, const std::vector<SetRef>& refs, const uint32_t* I, const uint32_t* J, size_t npairs, uint64_t* abc) {
    const int nsets = (int)refs.size();
    DBuf<SetRef> d_refs(c, nsets);
    ZB_CUDA(cudaMemcpyAsync(d_refs.get(), refs.data(), nsets * sizeof(SetRef), cudaMemcpyHostToDevice, c->stream));
    const size_t TILE = 4096;
    size_t p0 = 0;
    while (p0 < npairs) {
        // batch pairs so that one launch stays below 2^30 CTAs
        std::vector<uint64_t> tstart;
        tstart.push_back(0);
        size_t p1 = p0;
        while (p1 < npairs) {
            if (I[p1] >= (uint32_t)nsets || J[p1] >= (uint32_t)nsets) ZB_FAIL(ZB_E_ARG, "pair index out of range");
            const uint64_t t = div_up(refs[I[p1]].n + refs[J[p1]].n, TILE);
            if (tstart.back() + t > (1ull << 30) && p1 > p0) break;
            tstart.push_back(tstart.back() + t);
            p1++;
        }
        const size_t np = p1 - p0;
        DBuf<uint64_t> d_abc(c, 3 * np + np + 1);
        DBuf<uint32_t> d_ij(c, 2 * np);
        ZB_CUDA(dev_memset(c, d_abc.get(), 0, 3 * np * 8));
        ZB_CUDA(cudaMemcpyAsync(d_abc.get() + 3 * np, tstart.data(), (np + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(cudaMemcpyAsync(d_ij.get(), I + p0, np * 4, cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(cudaMemcpyAsync(d_ij.get() + np, J + p0, np * 4, cudaMemcpyHostToDevice, c->stream));
        pairs_abc(c, d_refs.get(), d_ij.get(), d_ij.get() + np, np, d_abc.get(), tstart.back());
        ZB_CUDA(cudaMemcpyAsync(abc + 3 * p0, d_abc.get(), 3 * np * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t p = p0; p < p1; p++) {
            const uint64_t a = abc[3 * p];
            abc[3 * p + 1] = refs[I[p]].n - a;
            abc[3 * p + 2] = refs[J[p]].n - a;
        }
        p0 = p1;
    }
}
}  // namespace zb

extern "C" {

// ------------------------------------------------------------------------------- diagnostics
// Stage-level entry points used by tests/ and bench.py (kernel isolation); not part of the drop-in surface.
int zb_dbg_sort_u64(int device, uint64_t* keys, uint32_t* vals, size_t n, int key_bits, int max_bits, int iters,
                    float* ms_per_sort) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    const int saved = g_sort_max_bits;
    if (max_bits >= 8 && max_bits <= 11) g_sort_max_bits = max_bits;
    DBuf<uint64_t> a(c, n), b(c, n), src(c, n);
    DBuf<uint32_t> va, vb, vsrc;
    if (vals) { va.alloc(c, n); vb.alloc(c, n); vsrc.alloc(c, n); }
    ZB_CUDA(cudaMemcpyAsync(src.get(), keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (vals) ZB_CUDA(cudaMemcpyAsync(vsrc.get(), vals, n * 4, cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    ZB_CUDA(cudaEventCreate(&e0));
    ZB_CUDA(cudaEventCreate(&e1));
    int which = 0;
    float total_ms = 0;
    if (iters < 1) iters = 1;
    for (int it = 0; it < iters; it++) {
        ZB_CUDA(dev_copy(c, a.get(), src.get(), n * 8));
        if (vals) ZB_CUDA(dev_copy(c, va.get(), vsrc.get(), n * 4));
        ZB_CUDA(cudaEventRecord(e0, c->stream));
        which = radix_sort(c, a.get(), b.get(), vals ? va.get() : nullptr, vals ? vb.get() : nullptr, n, key_bits);
        ZB_CUDA(cudaEventRecord(e1, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0;
        ZB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 || iters == 1) total_ms += ms;
    }
    g_sort_max_bits = saved;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_sort) *ms_per_sort = total_ms / (float)(iters > 1 ? iters - 1 : 1);
    ZB_CUDA(cudaMemcpyAsync(keys, which ? b.get() : a.get(), n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (vals) ZB_CUDA(cudaMemcpyAsync(vals, which ? vb.get() : va.get(), n * 4, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

// sort + count (segsort.cu) of host keys (+ optional weights); mode 0 = auto, 1 = classic full sort + RLE
int zb_dbg_sort_count(int device, const uint64_t* keys, const uint32_t* weights, size_t n, int key_bits, int mode,
                      int iters, uint64_t* out_k, uint32_t* out_c, size_t* n_out, float* ms_per_call) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    if (!n_out) ZB_FAIL(ZB_E_ARG, "null argument");
    const int saved = g_sort_count_mode;
    // mode: 0 = count (bucket route), 1 = classic, 2 = distinct + payload (bucket route), 3 / 4 = 0 / 2 on the segment route
    g_sort_count_mode = (mode == 1) ? 1 : (mode >= 3 ? 2 : 0);
    if (const char* e = getenv("ZB_SORT_CFG")) g_sort_cfg = atoi(e);
    DBuf<uint64_t> src(c, n), a(c, n), b(c, n), ok(c, n);
    DBuf<uint32_t> vsrc, va, vb, oc(c, n);
    if (weights) { vsrc.alloc(c, n); va.alloc(c, n); vb.alloc(c, n); }
    ZB_CUDA(cudaMemcpyAsync(src.get(), keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (weights) ZB_CUDA(cudaMemcpyAsync(vsrc.get(), weights, n * 4, cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    ZB_CUDA(cudaEventCreate(&e0));
    ZB_CUDA(cudaEventCreate(&e1));
    float total_ms = 0;
    size_t nd = 0;
    if (iters < 1) iters = 1;
    try {
        for (int it = 0; it < iters; it++) {
            ZB_CUDA(dev_copy(c, a.get(), src.get(), n * 8));
            if (weights) ZB_CUDA(dev_copy(c, va.get(), vsrc.get(), n * 4));
            ZB_CUDA(cudaEventRecord(e0, c->stream));
            nd = sort_count(c, a.get(), b.get(), weights ? va.get() : nullptr, weights ? vb.get() : nullptr, n, key_bits,
                            ok.get(), oc.get(), mode == 2 || mode == 4);
            ZB_CUDA(cudaEventRecord(e1, c->stream));
            ZB_CUDA(cudaStreamSynchronize(c->stream));
            float ms = 0;
            ZB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 || iters == 1) total_ms += ms;
        }
    } catch (...) {
        g_sort_count_mode = saved;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        throw;
    }
    g_sort_count_mode = saved;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_call) *ms_per_call = total_ms / (float)(iters > 1 ? iters - 1 : 1);
    *n_out = nd;
    if (nd && out_k) ZB_CUDA(cudaMemcpyAsync(out_k, ok.get(), nd * 8, cudaMemcpyDeviceToHost, c->stream));
    if (nd && out_c) ZB_CUDA(cudaMemcpyAsync(out_c, oc.get(), nd * 4, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_dbg_parse(int device, const uint8_t* raw, size_t n, int is_fasta, uint8_t* codes, size_t* n_codes,
                 uint64_t* n_records) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    DBuf<uint8_t> d(c, n + 16), cd(c, n + 64);
    ZB_CUDA(cudaMemcpyAsync(d.get(), raw, n, cudaMemcpyHostToDevice, c->stream));
    if (is_fasta) parse_fasta(c, d.get(), n, cd.get(), n_codes, n_records);
    else parse_fastq(c, d.get(), n, cd.get(), n_codes, n_records);
    if (codes && *n_codes) ZB_CUDA(cudaMemcpyAsync(codes, cd.get(), *n_codes, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_dbg_extract(int device, int k, const uint8_t* codes, size_t n, uint64_t* keys, size_t* n_keys) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    if (k < 1 || k > 32) ZB_FAIL(ZB_E_ARG, "k out of range");
    const size_t padded = div_up(n, EXTRACT_TILE) * EXTRACT_TILE + 32;
    DBuf<uint8_t> cd(c, 32 + padded);
    DBuf<uint64_t> out(c, padded);
    DBuf<unsigned long long> cnt(c, 1);
    ZB_CUDA(dev_memset(c, cd.get(), 4, 32 + padded));
    ZB_CUDA(dev_memset(c, cnt.get(), 0, 8));
    if (n) ZB_CUDA(cudaMemcpyAsync(cd.get() + 32, codes, n, cudaMemcpyHostToDevice, c->stream));
    extract_canonical(c, k, cd.get() + 32, n, out.get(), cnt.get());
    ZB_CUDA(read_back(c, cnt.get(), 8));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    *n_keys = (size_t)c->h_scalars[0];
    if (keys && *n_keys) ZB_CUDA(cudaMemcpy(keys, out.get(), *n_keys * 8, cudaMemcpyDeviceToHost));
    ZB_CATCH
}

int zb_dbg_guard_check(int device, uint64_t* n_blocks, uint64_t* n_bad) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    if (n_blocks) *n_blocks = 0;
    if (n_bad) *n_bad = 0;
    if (!guard_on()) ZB_FAIL(ZB_E_ARG, "guard bands are off: set ZB_GUARD=1 before the library is first used");
    std::vector<uint64_t> blocks;
    uint64_t released_bad = 0, ever = 0;
    {
        std::lock_guard<std::mutex> lk(c->alloc_mu);
        for (auto& kv : c->live_blocks) {
            blocks.push_back((uint64_t)(uintptr_t)kv.first);
            blocks.push_back((uint64_t)kv.second.user);
        }
        released_bad = c->guard_bad;
        ever = c->tick;
    }
    const size_t n = blocks.size() / 2;
    if (n_blocks) *n_blocks = ever;      // allocations + releases seen so far: every release was checked
    uint64_t bad = 0;
    if (n) {
        uint64_t* d = nullptr;   // outside the guarded allocator: the list must not change while it is scanned
        ZB_CUDA(cudaMalloc((void**)&d, (blocks.size() + 1) * 8));
        ZB_CUDA(cudaMemcpyAsync(d + 1, blocks.data(), blocks.size() * 8, cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(cudaMemsetAsync(d, 0, 8, c->stream));
        guard_check_kernel<<<(unsigned)div_up(n, 8), 256, 0, c->stream>>>(d + 1, (uint32_t)n, reinterpret_cast<unsigned long long*>(d));
        c->launches++;
        ZB_CUDA(cudaMemcpyAsync(&bad, d, 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(d);
    }
    if (n_bad) *n_bad = bad + released_bad;
    ZB_CATCH
}

// per-stage CUDA-event timing of everything this context runs between on=1 and the report
int zb_dbg_profile(int device, int on, char* report, size_t cap) {
    ZB_TRY
    ctx_for(device);
    const std::vector<Ctx*> all = contexts_of(device);   // stages of every host thread's context on this device
    for (Ctx* c : all) ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (report && cap) {
        std::map<std::string, std::pair<double, int>> agg;
        std::vector<std::string> order;
        for (Ctx* c : all) {
            for (auto& r : c->stages) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = -1;
                if (!agg.count(r.name)) order.push_back(r.name);
                agg[r.name].first += ms;
                agg[r.name].second += 1;
            }
        }
        std::string out;
        char line[160];
        for (auto& nm : order) {
            snprintf(line, sizeof line, "%s %.4f %d\n", nm.c_str(), agg[nm].first, agg[nm].second);
            out += line;
        }
        snprintf(report, cap, "%s", out.c_str());
    }
    for (Ctx* c : all) {
        for (auto& r : c->stages) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        c->stages.clear();
        c->profile = (on != 0);
    }
    g_profile_on[device & 63] = (on != 0);
    ZB_CATCH
}

// device timer on the library stream: op 0 = start (record), op 1 = stop (record, sync, elapsed ms)
int zb_dbg_timer(int device, int op, float* ms) {
    ZB_TRY
    Ctx* c = ctx_for(device);
    static thread_local cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!e0) { ZB_CUDA(cudaEventCreate(&e0)); ZB_CUDA(cudaEventCreate(&e1)); }
    if (op == 0) {
        ZB_CUDA(cudaEventRecord(e0, c->stream));
    } else {
        ZB_CUDA(cudaEventRecord(e1, c->stream));
        ZB_CUDA(cudaEventSynchronize(e1));
        float t = 0;
        ZB_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (ms) *ms = t;
    }
    ZB_CATCH
}

}  // extern "C"
