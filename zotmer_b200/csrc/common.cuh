// common.cuh -- shared host/device helpers for the zotmer_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zotmer_b200.h"

namespace zb {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

struct Fail {
    int code;
};

#define ZB_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            zb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            throw zb::Fail{ZB_E_CUDA};                                                            \
        }                                                                                         \
    } while (0)

#define ZB_FAIL(code, ...)            \
    do {                              \
        zb::set_error(__VA_ARGS__);   \
        throw zb::Fail{code};         \
    } while (0)

// ----------------------------------------------------------------------------- device context
// One context per (host thread, device): a stream, a caching device allocator (so repeated batches do not pay
// cudaMalloc), and a few scratch scalars in pinned memory.  A handle stays with the context that created it.
struct Ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;    // second stream of the context (created on first use): a routing kernel that runs beside the sort
    cudaEvent_t side_ev = nullptr;
    cudaMemPool_t pool = nullptr;
    uint64_t* h_scalars = nullptr;  // pinned, 64 x u64
    uint32_t* d_h_scalars = nullptr;  // the same buffer as the device sees it
    uint8_t* h_big = nullptr;       // pinned + mapped, H_BIG bytes: read-backs of a few KB (histograms) by a kernel
    uint8_t* d_h_big = nullptr;
    uint64_t launches = 0;          // kernels launched through this context (bench "gpu_launches")
    // optional per-stage CUDA-event timing (zb_dbg_profile): name, start, stop
    // caching allocator state (api.cu): segments from cudaMalloc, cut into blocks that are split on demand and merged
    // with their free neighbours on release
    struct Seg;
    struct Blk { size_t size; bool free; };
    struct Seg {
        char* base = nullptr;
        size_t size = 0;
        bool small = false;                        // segment of the small-block pool
        std::map<size_t, Blk> blocks;              // offset -> block, covering the segment without gaps
        size_t free_bytes = 0;
        uint64_t last_use = 0;
    };
    std::map<char*, Seg*> segs;                                        // by base address
    std::multimap<size_t, std::pair<Seg*, size_t>> free_idx[2];         // [large, small]: size -> (segment, offset)
    struct Live { Seg* seg; size_t off; size_t user; };                  // user = requested bytes (rounded up to 16)
    std::unordered_map<void*, Live> live_blocks;                        // user pointer -> its block
    size_t cached_bytes = 0, live_bytes = 0;      // free / handed-out bytes inside the segments
    uint64_t tick = 0;
    size_t cache_limit = 0;                        // free bytes above which a miss gives wholly free segments back (a quarter of the device)
    uint64_t guard_bad = 0;                        // blocks released with a damaged guard band (ZB_GUARD=1)
    std::mutex alloc_mu;
    cudaEvent_t copy_ev[3] = {nullptr, nullptr, nullptr};   // bulk host <-> device copies: at most three pieces queued
    bool profile = false;
    struct StageRec { const char* name; cudaEvent_t e0, e1; };
    std::vector<StageRec> stages;
};

// RAII stage marker: records events on the context stream when profiling is on (no sync).
struct Stage {
    Ctx* c;
    int idx = -1;
    Stage(Ctx* c_, const char* name) : c(c_) {
        if (!c->profile) return;
        Ctx::StageRec r;
        r.name = name;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, c->stream);
        idx = (int)c->stages.size();
        c->stages.push_back(r);
    }
    ~Stage() {
        if (idx >= 0) cudaEventRecord(c->stages[idx].e1, c->stream);
    }
};

Ctx* ctx_for(int device);

// Few-byte device -> host read-back into c->h_scalars, done by a one-warp kernel that stores into the (mapped) pinned
// buffer instead of a copy-engine transfer: a small cudaMemcpyAsync queues behind whatever bulk copies OTHER host
// threads have in flight on the engine (bench.py's pipelined e2e: the sort + count of one step waited up to 5.7 ms
// for the 315 MB input copy of the next step).  Synchronise the stream before reading h_scalars.
cudaError_t read_back(Ctx* c, const void* d_src, size_t bytes);
// the same for up to H_BIG bytes (a multiple of 16, 16-byte aligned source) into c->h_big
static const size_t H_BIG = 64 << 10;
cudaError_t read_back_big(Ctx* c, const void* d_src, size_t bytes);
// `bytes` (multiple of 16, 16-byte aligned source) to c->h_big + host_off; several of these, then ONE synchronisation
void copy_small_to_host(Ctx* c, const void* d_src, size_t host_off, size_t bytes);
// Fill / device-to-device copy done by kernels on the context's stream, for the same reason: cudaMemsetAsync and
// cudaMemcpyAsync(DeviceToDevice) may be served by a copy engine and then queue behind other threads' bulk transfers.
cudaError_t dev_memset(Ctx* c, void* p, int value, size_t bytes);
cudaError_t dev_copy(Ctx* c, void* dst, const void* src, size_t bytes);

// Device memory comes from a per-context caching allocator (api.cu): blocks are cudaMalloc'ed once,
// kept in a size-ordered free list and handed out again on the next request of a similar size.  All
// work of a context is ordered on its one stream, so a block freed by the host can be reused by the
// next kernel without synchronisation.  (cudaMallocAsync pools showed 100-600 ms stalls when a
// 1 GB block had to be re-mapped between steps.)
void* dalloc(Ctx* c, size_t bytes);
void dfree(Ctx* c, void* p);
void dtrim(Ctx* c);  // give every cached block back to the driver
void dshrink(Ctx* c, void* p, size_t bytes);  // keep the first `bytes` of a block, free the rest (no copy)

// RAII device buffer bound to a context's stream/pool
template <typename T>
struct DBuf {
    Ctx* c = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(Ctx* c_, size_t n_) : c(c_), n(n_) { p = (T*)dalloc(c, n * sizeof(T)); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : c(o.c), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); c = o.c; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    void release() { if (p) { dfree(c, p); p = nullptr; n = 0; } }
    void alloc(Ctx* c_, size_t n_) { release(); c = c_; n = n_; p = (T*)dalloc(c, n * sizeof(T)); }
    void shrink(size_t n_) { if (p && n_ < n) { dshrink(c, p, n_ * sizeof(T)); n = n_; } }
    T* get() const { return p; }
};

#define ZB_LAUNCH_CHECK(c)                       \
    do {                                         \
        (c)->launches++;                         \
        ZB_CUDA(cudaGetLastError());             \
    } while (0)

static inline size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// status words of the chained scans: GPU-scope relaxed accesses (served by L2; `ld.volatile` compiles
// to .STRONG.SYS, system scope, which is slower to poll)
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// streaming (read-once) 128-bit load / store that do not pollute L1
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
    const unsigned l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= (unsigned)o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread.  THREADS multiple of 32, <= 1024.
// `smem` needs THREADS/32 + 1 elements.  Returns exclusive prefix; *total gets the block sum.
template <int THREADS, typename T, bool TRAILING_SYNC = true>
__device__ __forceinline__ T block_excl_scan(T v, T* smem, T* total) {
    const unsigned l = lane_id(), w = threadIdx.x >> 5;
    T inc = warp_incl_scan(v);
    if (l == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (l < THREADS / 32) ? smem[l] : T(0);
        T xi = warp_incl_scan(x);
        if (l < THREADS / 32) smem[l] = xi - x;
        if (l == 31) smem[THREADS / 32] = xi;
    }
    __syncthreads();
    T r = smem[w] + inc - v;
    *total = smem[THREADS / 32];
    if (TRAILING_SYNC) __syncthreads();   // only needed when `smem` is reused by a later call
    return r;
}

// ---- bucket offsets of sorted key arrays: off[b * nsets + i] = first index of set i whose (key - key_base) >> shift is >= b
// (b = 0 .. nb; `off` zeroed by the caller; blockIdx.y = set).  One streaming pass over the keys: element j closes
// every bucket between its predecessor's and its own.  Two keys per thread (one 16-byte load), the predecessor of
// the first one comes from the neighbouring lane.  Ref = any struct with members `k` (sorted u64 keys) and `n`.
template <typename Ref>
__device__ __forceinline__ void bucket_offsets_body(const Ref* __restrict__ sets, int nsets, int shift, uint32_t nb,
                                                    uint32_t* __restrict__ off, uint64_t key_base = 0) {
    const int i = blockIdx.y;
    const uint64_t n = sets[i].n;
    const uint64_t* __restrict__ k = sets[i].k;
    const bool vec = (((uintptr_t)k) & 15) == 0;
    const uint64_t npair = (n + 1) >> 1;
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < npair; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = base + lane;
        uint64_t k0 = 0, k1 = 0;
        const bool has0 = 2 * t < n, has1 = 2 * t + 1 < n;
        if (has1 && vec) {
            const uint4 v = ld_stream_v4(k + 2 * t);
            k0 = (((uint64_t)v.y << 32) | v.x) - key_base;
            k1 = (((uint64_t)v.w << 32) | v.z) - key_base;
        } else {
            if (has0) k0 = __ldg(k + 2 * t) - key_base;
            if (has1) k1 = __ldg(k + 2 * t + 1) - key_base;
        }
        uint64_t prev = __shfl_up_sync(0xffffffffu, k1, 1);
        if (lane == 0 && has0 && t > 0) prev = __ldg(k + 2 * t - 1) - key_base;
        // fill tasks: off[b * nsets + i] = val for b in (lo, hi].  Short gaps (the usual case: neighbouring keys in the
        // same or the next bucket) are written by the thread; a long gap -- the buckets in front of a set's first key,
        // behind its last one, or an empty stretch of the key space -- is written by the whole warp (one thread took
        // milliseconds over the hundreds of thousands of empty buckets of a key-range slab)
        uint64_t lo_[3], hi_[3], val_[3];
#pragma unroll
        for (int u = 0; u < 3; u++) { lo_[u] = 0; hi_[u] = 0; val_[u] = 0; }
        if (has0) {
            const uint64_t j = 2 * t;
            const uint64_t cur = (shift < 64) ? (k0 >> shift) : 0ull;
            const uint64_t pb = (shift < 64 && j > 0) ? (prev >> shift) : 0ull;
            lo_[0] = pb; hi_[0] = cur; val_[0] = j;      // (j == 0: buckets 1 .. cur start at 0; off[0][i] = 0 already)
            if (j == n - 1) { lo_[2] = cur; hi_[2] = nb; val_[2] = n; }
        }
        if (has1) {
            const uint64_t j = 2 * t + 1;
            const uint64_t cur = (shift < 64) ? (k1 >> shift) : 0ull;
            const uint64_t pb = (shift < 64) ? (k0 >> shift) : 0ull;
            lo_[1] = pb; hi_[1] = cur; val_[1] = j;
            if (j == n - 1) { lo_[2] = cur; hi_[2] = nb; val_[2] = n; }
        }
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const bool longgap = hi_[u] > lo_[u] + 8;
            if (!longgap)
                for (uint64_t bb = lo_[u] + 1; bb <= hi_[u]; bb++) off[bb * nsets + i] = (uint32_t)val_[u];
            unsigned pend = __ballot_sync(0xffffffffu, longgap);
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                const uint64_t glo = __shfl_sync(0xffffffffu, lo_[u], src), ghi = __shfl_sync(0xffffffffu, hi_[u], src);
                const uint32_t gv = (uint32_t)__shfl_sync(0xffffffffu, val_[u], src);
                for (uint64_t bb = glo + 1 + lane; bb <= ghi; bb += 32) off[bb * nsets + i] = gv;
            }
        }
    }
}

// ---- chained scan ("decoupled look-back") over tiles, one 64-bit status word per tile.
// status: bits 63..62 = 0 invalid / 1 tile aggregate / 2 inclusive prefix; bits 61..0 value.
// The status array must be zeroed before the launch and tile ids must be handed out in launch
// order (atomic ticket) so that every predecessor of a running tile is itself running or done.
// Call with ALL 32 lanes of one warp; returns the exclusive prefix of `agg` in every lane.
#define ZB_ST_AGG (1ull << 62)
#define ZB_ST_PFX (2ull << 62)
#define ZB_ST_VAL ((1ull << 62) - 1)
__device__ __forceinline__ uint64_t lookback_u64(uint64_t* status, uint32_t tile, uint64_t agg) {
    const unsigned l = lane_id();
    if (tile == 0) {
        if (l == 0) st_volatile_u64(status, ZB_ST_PFX | agg);
        return 0;
    }
    if (l == 0) st_volatile_u64(status + tile, ZB_ST_AGG | agg);
    uint64_t excl = 0;
    int64_t base = (int64_t)tile - 1;
    while (true) {
        int64_t idx = base - (int64_t)l;
        uint64_t v;
        unsigned inval, haspfx;
        do {
            v = (idx >= 0) ? ld_volatile_u64(status + idx) : ZB_ST_PFX;
            inval = __ballot_sync(0xffffffffu, (v >> 62) == 0);
            haspfx = __ballot_sync(0xffffffffu, (v >> 62) == 2);
            // only lanes up to (and including) the nearest inclusive prefix matter
            unsigned need = haspfx ? ((2u << (__ffs(haspfx) - 1)) - 1u) : 0xffffffffu;
            inval &= need;
        } while (inval);
        int firstp = haspfx ? (__ffs(haspfx) - 1) : 32;
        uint64_t contrib = ((int)l <= firstp) ? (v & ZB_ST_VAL) : 0ull;
        excl += warp_sum(contrib);
        if (firstp < 32) break;
        base -= 32;
    }
    if (l == 0) st_volatile_u64(status + tile, ZB_ST_PFX | (excl + agg));
    return excl;
}

// ---- 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif  // __CUDACC__

}  // namespace zb
