// sort.cu -- hand-written LSD radix sort for u64 k-mer keys (optionally with a u32 payload).
//
// Replaces zotmer/library/misc.py:400-424 (radix_sort: MSD buckets + list.sort) -- the result is the
// same ascending order.  Design ("onesweep"): one upfront kernel reads the keys once and builds the
// histograms of ALL digits; then one kernel per digit does load -> in-tile stable ranking (warp
// match + per-warp counters) -> chained scan across tiles (decoupled look-back, one status word per
// (tile, bin)) -> shared-memory staged scatter.  Per pass every key is read once and written once:
// 16 B/key/pass (+4 B with payload), the HBM floor for an out-of-place LSD pass.
//
// Digits: the 2k significant key bits are split EVENLY over ceil(2k / maxbits) passes so that every
// digit is fully populated (k=25: 50 bits -> 7 x (8,7,7,7,7,7,7) at maxbits=8, or 5 x 10 at 10).
#include <vector>

#include "kernels.h"

namespace zb {

int g_sort_max_bits = 8;

static constexpr int MAX_PASSES = 8;
int g_sort_cfg = 0;   // 0: 256 threads x 16 keys per CTA, 1: 512 threads x 8 keys (ZB_SORT_CFG; 256-bin digits only)
// threads per CTA for a given digit width: the tile (threads x 16 keys) grows with the number of bins so
// that the per-digit work (warp prefix, chained scan) per key stays constant
template <int BINS> struct SortCfg { static constexpr int THREADS = (BINS <= 256) ? 256 : 512; };

struct SortPlan {
    int passes;
    int shift[MAX_PASSES];
    int bits[MAX_PASSES];
};

// ------------------------------------------------------------------------------- histogram
// All digits of all passes in one read of the keys.  Shared-memory bins, one RED per key per pass.
template <int BINS>
__global__ void __launch_bounds__(512) sort_hist_kernel(const uint64_t* __restrict__ keys, size_t n, SortPlan plan,
                                                        uint32_t* __restrict__ ghist /*[passes][BINS]*/) {
    extern __shared__ uint32_t sh[];  // [passes][BINS]
    const int total = plan.passes * BINS;
    for (int i = threadIdx.x; i < total; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        uint64_t a, b;
        bool two = (i + 1 < n);
        if (two) {
            ulonglong2 v = *reinterpret_cast<const ulonglong2*>(keys + i);
            a = v.x; b = v.y;
        } else {
            a = keys[i]; b = 0;
        }
#pragma unroll
        for (int p = 0; p < MAX_PASSES; p++) {
            if (p < plan.passes) {
                const uint32_t m = (1u << plan.bits[p]) - 1u;
                atomicAdd(&sh[p * BINS + ((uint32_t)(a >> plan.shift[p]) & m)], 1u);
                if (two) atomicAdd(&sh[p * BINS + ((uint32_t)(b >> plan.shift[p]) & m)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        uint32_t v = sh[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

// exclusive scan of each pass's histogram -> global start of each bin.  One block per pass.
template <int BINS>
__global__ void __launch_bounds__(BINS > 1024 ? 1024 : BINS) sort_scan_kernel(uint32_t* __restrict__ ghist) {
    constexpr int T = BINS > 1024 ? 1024 : BINS;
    constexpr int PER = BINS / T;
    __shared__ uint32_t sm[T / 32 + 1];
    uint32_t* h = ghist + (size_t)blockIdx.x * BINS;
    uint32_t v[PER];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { v[i] = h[threadIdx.x * PER + i]; s += v[i]; }
    uint32_t tot;
    uint32_t ex = block_excl_scan<T, uint32_t>(s, sm, &tot);
#pragma unroll
    for (int i = 0; i < PER; i++) { h[threadIdx.x * PER + i] = ex; ex += v[i]; }
}

// ------------------------------------------------------------------------------- onesweep pass
#define ST32_AGG 0x40000000u
#define ST32_PFX 0x80000000u
#define ST32_VAL 0x3fffffffu

// STABLE = false: keys of one digit value may leave the tile in any order (allowed for the FIRST pass of an
// LSD sort, where no earlier order has to be kept): the rank inside the tile is simply the return value of one
// shared-memory atomicAdd per key, no warp matching and no per-warp histograms.
template <int BINS, int SORT_THREADS, int SORT_ITEMS, bool HAS_VALS, bool STABLE>
__global__ void __launch_bounds__(SORT_THREADS, (SORT_ITEMS == 8 ? 3 : (SORT_THREADS * (BINS <= 512 ? 64 : 128) <= 16384 ? 4 : 2)))
onesweep_kernel(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout, const uint32_t* __restrict__ vin,
                uint32_t* __restrict__ vout, uint32_t n, int shift, int bits,
                const uint32_t* __restrict__ bin_start /*[BINS] global exclusive*/,
                uint32_t* __restrict__ status /*[tiles][BINS], zeroed*/, uint32_t* __restrict__ ticket, uint64_t uniform_mask) {
    constexpr int SORT_WARPS = SORT_THREADS / 32;
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);                       // [TILE]
    uint32_t* svals = reinterpret_cast<uint32_t*>(skeys + SORT_TILE);              // [TILE] (if HAS_VALS)
    uint32_t* whist = svals + (HAS_VALS ? SORT_TILE : 0);                          // [WARPS][BINS]
    uint32_t* s_goff = whist + SORT_WARPS * BINS;                                  // [BINS] global pos - tile pos
    // match masks [2][WARPS][BINS] live in the key staging area, which is idle while keys are ranked
    uint32_t* mmask = reinterpret_cast<uint32_t*>(smem_raw);
    static_assert(2 * SORT_WARPS * BINS * 4 <= SORT_TILE * 8 || BINS > 256, "mask area");
    __shared__ uint32_t s_scan[SORT_THREADS / 32 + 1];
    __shared__ uint32_t s_tile;

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dmask = (1u << bits) - 1u;
    const int nbins = 1 << bits;
    constexpr bool ATOMIC_MATCH = (2 * SORT_WARPS * BINS * 4 <= SORT_TILE * 8);

    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4* w4 = reinterpret_cast<uint4*>(whist);
        for (int i = tid; i < SORT_WARPS * BINS / 4; i += SORT_THREADS) w4[i] = z;
        if (ATOMIC_MATCH) {
            uint4* m4 = reinterpret_cast<uint4*>(mmask);
            for (int i = tid; i < 2 * SORT_WARPS * BINS / 4; i += SORT_THREADS) m4[i] = z;
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * SORT_TILE;
    const uint32_t n_valid = min((uint32_t)SORT_TILE, n - tile_base);

    // ---- load, warp-striped: item j of lane l sits at warp_base + j*32 + l
    uint64_t key[SORT_ITEMS];
    const uint32_t wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) {
        uint32_t idx = wbase + j * 32 + lane;
        key[j] = (idx < n_valid) ? __ldg(kin + tile_base + idx) : ~0ull;
    }

    uint16_t pos[SORT_ITEMS];
    constexpr int PER = (BINS + SORT_THREADS - 1) / SORT_THREADS;  // digits per thread (blocked)
    uint32_t cnt[PER], binst[PER];
    // A later pass of a sort whose caller accepts any order among keys that agree in the sorted bits (keys-only sorts,
    // and the top-bit passes in front of the bucket kernels): the input is ordered by the digits already done
    // (`uniform_mask`), so a tile whose first and last key agree in them holds ONE value of those digits -- 4096
    // consecutive keys of 125 M nearly always do -- and may rank without regard to arrival order, like a first pass.
    bool unstable = !STABLE;
    if (STABLE && uniform_mask != 0 && n_valid > 0)
        unstable = ((__ldg(kin + tile_base) ^ __ldg(kin + tile_base + n_valid - 1)) & uniform_mask) == 0;
    if (unstable) {
        // ---- unstable ranking: rank inside (tile, digit) = return value of the atomic
        uint32_t* thist = whist;            // [BINS] counts, then tile-local bin starts
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {   // padding keys of the last tile take no part
            const bool valid = wbase + j * 32 + lane < n_valid;
            pos[j] = valid ? (uint16_t)atomicAdd(&thist[(uint32_t)(key[j] >> shift) & dmask], 1u) : (uint16_t)0;
        }
        __syncthreads();
        uint32_t csum = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int d = tid * PER + q;
            uint32_t s = 0;
            if (d < nbins) {
                s = thist[d];
                st_volatile_u32(status + (size_t)tile * BINS + d, (tile == 0 ? ST32_PFX : ST32_AGG) | s);
            }
            cnt[q] = s;
            csum += s;
        }
        uint32_t tot;
        uint32_t ex = block_excl_scan<SORT_THREADS, uint32_t, false>(csum, s_scan, &tot);
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int d = tid * PER + q;
            binst[q] = ex;
            if (d < nbins) thist[d] = ex;
            ex += cnt[q];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {
            const uint32_t idx = wbase + j * 32 + lane;
            pos[j] = (idx < n_valid) ? (uint16_t)(pos[j] + thist[(uint32_t)(key[j] >> shift) & dmask]) : (uint16_t)idx;
        }
    } else {
    // ---- (1) early counts: per-warp digit histogram with one shared-memory atomic per key, so that
    // the tile's aggregate can be published BEFORE the (long) ranking phase and the chained scan of
    // the following tiles never has to wait for it.
    uint32_t* myhist = whist + warp * BINS;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) atomicAdd(&myhist[(uint32_t)(key[j] >> shift) & dmask], 1u);
    __syncthreads();

    // ---- (2) per digit: exclusive prefix over warps -> tile count, publish it, tile-local bin starts
    uint32_t csum = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int d = tid * PER + q;
        uint32_t s = 0;
        if (d < nbins) {
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) {
                uint32_t c = whist[w * BINS + d];
                whist[w * BINS + d] = s;
                s += c;
            }
            st_volatile_u32(status + (size_t)tile * BINS + d, (tile == 0 ? ST32_PFX : ST32_AGG) | s);
        }
        cnt[q] = s;
        csum += s;
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan<SORT_THREADS, uint32_t, false>(csum, s_scan, &tot);
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int d = tid * PER + q;
        binst[q] = ex;
        if (d < nbins) {
            // the per-(warp,digit) words become running cursors = position inside the sorted tile
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) whist[w * BINS + d] += ex;
        }
        ex += cnt[q];
    }
    __syncthreads();

    // ---- (3) stable rank inside the warp.  Lanes holding the same digit form a group; the group mask
    // is built by OR-ing lane bits into a per-(warp,digit) shared word (one ATOMS per key; MATCH.ANY is
    // microcoded on sm_100 and a ballot per digit bit costs ~40 instructions per key, see profiles/).
    // Every lane reads the cursor of its (warp,digit), the lowest lane of the group advances it.
    const unsigned lt = lanemask_lt();
    const uint32_t lanebit = 1u << lane;
    if (ATOMIC_MATCH) {
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
            uint32_t* mm = mmask + ((j & 1) * SORT_WARPS + warp) * BINS + d;   // double-buffered
            atomicOr(mm, lanebit);
            __syncwarp();
            const uint32_t m = *mm;
            const uint32_t cur = myhist[d];
            __syncwarp();
            if ((m & lt) == 0) {
                *mm = 0;
                myhist[d] = cur + __popc(m);
            }
            pos[j] = (uint16_t)(cur + __popc(m & lt));
        }
    } else {
        constexpr int LOG_BINS = (BINS == 256) ? 8 : (BINS == 512) ? 9 : (BINS == 1024) ? 10 : 11;
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
            unsigned m = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < LOG_BINS; b++) {
                const bool bit = (d >> b) & 1u;
                const unsigned v = __ballot_sync(0xffffffffu, bit);
                m &= bit ? v : ~v;
            }
            const uint32_t cur = myhist[d];
            __syncwarp();
            if ((m & lt) == 0) myhist[d] = cur + __popc(m);
            pos[j] = (uint16_t)(cur + __popc(m & lt));
            __syncwarp();
        }
    }

    }

    // ---- (4) chained scan across tiles, one status word per (tile, digit); predecessors published
    // their aggregates before their own ranking phase, so this rarely spins.
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int d = tid * PER + q;
        if (d < nbins) {
            uint32_t* st = status + (size_t)tile * BINS + d;
            uint32_t excl = 0;
            if (tile != 0) {
                // walk back over the predecessors, LB_W status words in flight at a time
                constexpr int LB_W = 4;
                int64_t p = (int64_t)tile - 1;
                bool done = false;
                while (!done) {
                    uint32_t v[LB_W];
#pragma unroll
                    for (int u = 0; u < LB_W; u++)
                        v[u] = (p - u >= 0) ? ld_volatile_u32(st - (size_t)(tile - (p - u)) * BINS) : ST32_PFX;
#pragma unroll
                    for (int u = 0; u < LB_W; u++) {
                        if (!done) {
                            if ((v[u] & (ST32_AGG | ST32_PFX)) == 0) break;   // not published yet: poll again from here
                            excl += v[u] & ST32_VAL;
                            p--;
                            if (v[u] & ST32_PFX) done = true;
                        }
                    }
                }
                st_volatile_u32(st, ST32_PFX | (excl + cnt[q]));
            }
            s_goff[d] = bin_start[d] + excl - binst[q];
        }
    }
    __syncthreads();   // masks are dead from here on: the staging area now receives the keys

    // ---- (5) place into tile-sorted order in shared memory
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) skeys[pos[j]] = key[j];
    if (HAS_VALS) {
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {
            uint32_t idx = wbase + j * 32 + lane;
            svals[pos[j]] = (idx < n_valid) ? __ldg(vin + tile_base + idx) : 0u;
        }
    }
    __syncthreads();

    // ---- coalesced-by-bin scatter: consecutive threads write consecutive keys of a bin
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) {
        const uint32_t i = j * SORT_THREADS + tid;
        if (i < n_valid) {
            const uint64_t kx = skeys[i];
            const uint32_t d = (uint32_t)(kx >> shift) & dmask;
            const uint32_t g = s_goff[d] + i;
            kout[g] = kx;
            if (HAS_VALS) vout[g] = svals[i];
        }
    }
}

template <int BINS, int SORT_THREADS, int SORT_ITEMS>
static size_t onesweep_smem(bool vals) {
    constexpr int SORT_WARPS = SORT_THREADS / 32;
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    return (size_t)SORT_TILE * 8 + (vals ? (size_t)SORT_TILE * 4 : 0) + (size_t)SORT_WARPS * BINS * 4 + BINS * 4;
}

int sort_digit_plan(int lo_bit, int nbits, int maxbits, int* shift, int* bits, int max_passes) {
    int passes = (nbits + maxbits - 1) / maxbits;
    if (passes < 1) passes = 1;
    if (passes > max_passes) return -passes;
    int base = nbits / passes, rem = nbits % passes, sh = lo_bit;
    for (int p = 0; p < passes; p++) {
        bits[p] = base + (p < rem ? 1 : 0);
        if (bits[p] < 1) bits[p] = 1;
        shift[p] = sh;
        sh += bits[p];
    }
    return passes;
}

template <int BINS, int SORT_THREADS, int SORT_ITEMS>
static int radix_sort_impl(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int lo_bit,
                           int key_bits, int maxbits, bool stable_first, const SortPre* pre) {
    SortPlan plan;
    plan.passes = sort_digit_plan(lo_bit, key_bits, maxbits, plan.shift, plan.bits, MAX_PASSES);
    if (plan.passes < 0) ZB_FAIL(ZB_E_ARG, "radix_sort: %d passes needed (max %d)", -plan.passes, MAX_PASSES);
    // histograms that came with the keys (extract.cu) are used when they were taken for exactly this plan
    bool have_hist = pre && pre->d_hist && BINS == 256 && pre->passes == plan.passes;
    for (int p = 0; have_hist && p < plan.passes; p++) have_hist = pre->shift[p] == plan.shift[p] && pre->bits[p] == plan.bits[p];
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    const bool vals = (v0 != nullptr);
    const uint32_t tiles = (uint32_t)div_up(n, SORT_TILE);
    DBuf<uint32_t> ghist(c, (size_t)plan.passes * BINS);
    DBuf<uint32_t> status(c, (size_t)tiles * BINS);
    DBuf<uint32_t> ticket(c, MAX_PASSES);
    ZB_CUDA(dev_memset(c, ticket.get(), 0, MAX_PASSES * 4));
    if (have_hist) {
        ZB_CUDA(dev_copy(c, ghist.get(), pre->d_hist, (size_t)plan.passes * BINS * 4));
        sort_scan_kernel<BINS><<<plan.passes, (BINS > 1024 ? 1024 : BINS), 0, c->stream>>>(ghist.get());
        ZB_LAUNCH_CHECK(c);
    } else {
        ZB_CUDA(dev_memset(c, ghist.get(), 0, (size_t)plan.passes * BINS * 4));
        Stage st_h(c, "sort_hist");
        int blocks = (int)std::min<size_t>((size_t)c->sm_count * 4, div_up(n, 512 * 2 * 4));
        if (blocks < 1) blocks = 1;
        size_t sm = (size_t)plan.passes * BINS * 4;
        if ((size_t)MAX_PASSES * BINS * 4 > 48 * 1024)   // a fixed limit per instantiation (see below)
            ZB_CUDA(cudaFuncSetAttribute(sort_hist_kernel<BINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_PASSES * BINS * 4));
        sort_hist_kernel<BINS><<<blocks, 512, sm, c->stream>>>(k0, n, plan, ghist.get());
        ZB_LAUNCH_CHECK(c);
        sort_scan_kernel<BINS><<<plan.passes, (BINS > 1024 ? 1024 : BINS), 0, c->stream>>>(ghist.get());
        ZB_LAUNCH_CHECK(c);
    }
    size_t sm = onesweep_smem<BINS, SORT_THREADS, SORT_ITEMS>(vals);
    {   // every instantiation gets ITS OWN fixed limit: the attribute is global to the function, and host threads
        // sorting keys and pairs at the same time must not lower each other's limit between attribute and launch
        const int sm_keys = (int)onesweep_smem<BINS, SORT_THREADS, SORT_ITEMS>(false);
        const int sm_pairs = (int)onesweep_smem<BINS, SORT_THREADS, SORT_ITEMS>(true);
        const int sm_limit = 227 * 1024;   // per CTA on sm_100; the 11-bit shape with a payload does not fit (keys only does)
        if (vals && sm_pairs > sm_limit) ZB_FAIL(ZB_E_ARG, "radix_sort: %d-bin passes cannot carry a payload", BINS);
        if (sm_pairs <= sm_limit) {
            ZB_CUDA(cudaFuncSetAttribute(onesweep_kernel<BINS, SORT_THREADS, SORT_ITEMS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_pairs));
            ZB_CUDA(cudaFuncSetAttribute(onesweep_kernel<BINS, SORT_THREADS, SORT_ITEMS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_pairs));
        }
        ZB_CUDA(cudaFuncSetAttribute(onesweep_kernel<BINS, SORT_THREADS, SORT_ITEMS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_keys));
        ZB_CUDA(cudaFuncSetAttribute(onesweep_kernel<BINS, SORT_THREADS, SORT_ITEMS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_keys));
    }
    uint64_t* kb[2] = {k0, k1};
    uint32_t* vb[2] = {v0, v1};
    int cur = 0;
    for (int p = 0; p < plan.passes; p++) {
        ZB_CUDA(dev_memset(c, status.get(), 0, (size_t)tiles * BINS * 4));
        Stage st_p(c, vals ? "sort_pass_pairs" : "sort_pass_keys");
        // the first pass has no earlier order to keep: unstable ranking (keys only, or when the caller allows it)
        const bool any_order = (!vals || !stable_first) && g_sort_cfg != 2;   // ties may come out in any order
        const bool stable = !(p == 0 && any_order);
        // bits [lo_bit, shift[p]): the digits the earlier passes have put in order
        const uint64_t below = (plan.shift[p] >= 64) ? ~0ull : ((1ull << plan.shift[p]) - 1ull);
        const uint64_t umask = (p > 0 && any_order && g_sort_cfg != 3) ? (below & ~((1ull << lo_bit) - 1ull)) : 0ull;
#define ZB_ONESWEEP(V, S)                                                                                          \
        onesweep_kernel<BINS, SORT_THREADS, SORT_ITEMS, V, S><<<tiles, SORT_THREADS, sm, c->stream>>>(              \
            kb[cur], kb[cur ^ 1], vals ? vb[cur] : nullptr, vals ? vb[cur ^ 1] : nullptr, (uint32_t)n, plan.shift[p], \
            plan.bits[p], ghist.get() + (size_t)p * BINS, status.get(), ticket.get() + p, umask)
        if (vals) { if (stable) ZB_ONESWEEP(true, true); else ZB_ONESWEEP(true, false); }
        else { if (stable) ZB_ONESWEEP(false, true); else ZB_ONESWEEP(false, false); }
#undef ZB_ONESWEEP
        ZB_LAUNCH_CHECK(c);
        cur ^= 1;
    }
    return cur;
}

int radix_sort_range(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int lo_bit, int nbits,
                     bool stable_first, const SortPre* pre) {
    if (n == 0) return 0;
    if (n >= (1ull << 30)) ZB_FAIL(ZB_E_ARG, "radix_sort: n=%zu exceeds 2^30 keys per batch", n);
    if (lo_bit < 0) lo_bit = 0;
    if (lo_bit > 63) lo_bit = 63;
    if (nbits < 1) nbits = 1;
    if (lo_bit + nbits > 64) nbits = 64 - lo_bit;
    int mb = g_sort_max_bits;
    if (mb <= 8 && g_sort_cfg == 1) return radix_sort_impl<256, 512, 8>(c, k0, k1, v0, v1, n, lo_bit, nbits, 8, stable_first, pre);
    if (mb <= 8) return radix_sort_impl<256, 256, 16>(c, k0, k1, v0, v1, n, lo_bit, nbits, 8, stable_first, pre);
    if (mb == 9) return radix_sort_impl<512, 512, 16>(c, k0, k1, v0, v1, n, lo_bit, nbits, 9, stable_first, nullptr);
    if (mb == 10) return radix_sort_impl<1024, 512, 16>(c, k0, k1, v0, v1, n, lo_bit, nbits, 10, stable_first, nullptr);
    return radix_sort_impl<2048, 512, 16>(c, k0, k1, v0, v1, n, lo_bit, nbits, 11, stable_first, nullptr);
}

int radix_sort(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits) {
    return radix_sort_range(c, k0, k1, v0, v1, n, 0, key_bits, true);
}

}  // namespace zb
