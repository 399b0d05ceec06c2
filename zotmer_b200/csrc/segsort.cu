// segsort.cu -- sort + run-length count of u64 k-mer keys: the kernel behind KmerAccumulator2.flush
// (zotmer/commands/kmerize.py:412-424: radix_sort of the pending list, misc.py:400-424, then merge()/RLE,
// kmerize.py:41-132).  Same result as a full LSD sort followed by reduce_by_key, with far fewer passes over
// the whole array:
//
//   1. stable LSD passes (sort.cu onesweep) over the TOP T = 8 * P bits of the key only.  P is chosen from n
//      so that the keys sharing those T bits ("a segment", contiguous after step 1) number ~1..16 on average
//      for well-spread keys, more where a key is repeated many times.
//   2. segsort_count_kernel: a CTA stages 4096 consecutive keys in shared memory and owns the segments that
//      START in its first 3072 positions (the last 1024 are the halo a segment may run into).  Inside a
//      tile every key is inserted into a shared-memory hash table (one CAS; the first key of a value becomes
//      its "head"), heads are compacted in position order, every key adds its weight to its head, every head
//      ranks itself among the (few) heads of its own segment.  A chained scan over CTAs turns the tile-local
//      head index into the global output position.  Work per key is O(distinct values per segment), so highly
//      repeated keys are cheap.
//   3. segments longer than 1024 keys ("big": one value repeated > 1024 times in the batch, or a badly skewed
//      key space) are skipped by step 2 and only listed; they are gathered into a side array, sorted and
//      counted by the classic path (radix_sort + reduce_by_key) and merged back (merge-path).  Usually a few
//      percent of the keys or nothing at all.
//
// Algorithmic bytes: P * 16 B/key (top passes) + 8 B/key (histogram) + 8 B/key (this kernel) + 12 B/distinct.
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace zb {

int g_sort_count_mode = 0;

static constexpr int SS_THREADS = 512;
static constexpr int SS_PER = 8;
static constexpr int SS_LOADED = SS_THREADS * SS_PER;      // positions staged by a CTA (4096)
static constexpr int SS_HALO = 1024;                        // longest segment ordered in shared memory
static constexpr int SS_TILE = SS_LOADED - SS_HALO;        // positions whose segments a CTA owns
static constexpr int SS_PADDED = SS_LOADED + SS_LOADED / SS_PER;
static constexpr int SS_HASH = 2 * SS_LOADED;              // slots of the tile-local hash table (load <= 0.5)
#define SS_EMPTY 0xffffffffu

// one pad slot per 8 keys: thread t reads positions 8t .. 8t+7, so the lane stride is 9 keys (no bank conflicts)
__device__ __forceinline__ int ss_idx(int q) { return q + (q >> 3); }

// highest set bit <= b in a 4096-bit mask (-1 if none)
__device__ __forceinline__ int mask_prev(const uint32_t* m, int b) {
    int w = b >> 5;
    uint32_t v = m[w] & (0xffffffffu >> (31 - (b & 31)));
    while (!v && w > 0) v = m[--w];
    return v ? (w << 5) + 31 - __clz(v) : -1;
}
// lowest set bit >= b (SS_LOADED if none); b may equal SS_LOADED
__device__ __forceinline__ int mask_next(const uint32_t* m, int b) {
    if (b >= SS_LOADED) return SS_LOADED;
    int w = b >> 5;
    uint32_t v = m[w] & (0xffffffffu << (b & 31));
    while (!v && w < SS_LOADED / 32 - 1) v = m[++w];
    return v ? (w << 5) + __ffs(v) - 1 : SS_LOADED;
}

// CTA `tile` writes its distinct keys, ordered, to tmp_k/tmp_c[tile * SS_LOADED ...] and their number to
// tile_heads[tile]; segcompact_kernel moves them to their final place once the per-tile counts are scanned
// (no CTA ever waits for another one).
// Shared memory: keys 36 KB + hash table 32 KB + 3 KB of bit masks / prefixes = 71 KB -> 3 CTAs (48 warps) per
// SM; the weighted form adds the staged weights (18 KB) and a 32-bit sum per position (16 KB) -> 2 CTAs per SM.
// MODE 0: count keys (weight 1 each); MODE 1: sum the u32 weights of equal keys; MODE 2: the keys are known to
// be distinct and carry a u32 payload (the mirror sort of kmerize) -- no hash table, every key is a head.
template <int MODE>
__global__ void __launch_bounds__(SS_THREADS, MODE == 1 ? 2 : 3)
segsort_count_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ w, uint64_t n, int lowbits,
                     uint64_t* __restrict__ tmp_k, uint32_t* __restrict__ tmp_c, uint32_t* __restrict__ tile_heads,
                     unsigned long long* __restrict__ big_n, uint64_t* __restrict__ big_start, uint64_t big_cap,
                     unsigned int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char ss_raw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(ss_raw);                  // [SS_PADDED] staged keys
    // slot = position of the value's head (low 16 bits) | number of keys with that value (high 16 bits,
    // unweighted form only: a segment has at most SS_HALO keys)
    uint32_t* table = reinterpret_cast<uint32_t*>(sk + SS_PADDED);       // [SS_HASH]
    uint32_t* sflag = table + (MODE == 2 ? 0 : SS_HASH);                 // [128] segment-start bits
    uint32_t* shead = sflag + SS_LOADED / 32;                            // [128] head bits
    uint32_t* shbase = shead + SS_LOADED / 32;                           // [SS_THREADS] heads before thread t's chunk
    uint32_t* sw = shbase + SS_THREADS;                                  // [SS_PADDED] weights      (WEIGHTED)
    uint32_t* wsum = sw + SS_PADDED;                                     // [SS_LOADED] sum per head position (WEIGHTED)
    __shared__ uint32_t s_scan[SS_THREADS / 32 + 1];
    __shared__ uint64_t s_halo;
    constexpr bool WEIGHTED = (MODE == 1);
    constexpr bool DISTINCT = (MODE == 2);

    const unsigned tid = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const uint64_t s = (uint64_t)tile * SS_TILE;
    const int nvalid = (int)min((uint64_t)SS_LOADED + 1, n - s);   // positions q < nvalid hold data
    if (!DISTINCT) {
        const uint4 e4 = make_uint4(SS_EMPTY, SS_EMPTY, SS_EMPTY, SS_EMPTY);
#pragma unroll
        for (int j = 0; j < SS_HASH / 4 / SS_THREADS; j++) reinterpret_cast<uint4*>(table)[j * SS_THREADS + tid] = e4;
        if (WEIGHTED) {
#pragma unroll
            for (int j = 0; j < SS_PER; j++) wsum[j * SS_THREADS + tid] = 0;
        }
    }
    // ---- stage the keys (coalesced)
#pragma unroll
    for (int j = 0; j < SS_PER; j++) {
        const int q = j * SS_THREADS + tid;
        sk[ss_idx(q)] = (q < nvalid) ? __ldg(keys + s + q) : 0ull;
        if (WEIGHTED) sw[ss_idx(q)] = (q < nvalid) ? __ldg(w + s + q) : 0u;
    }
    if (tid == 0) s_halo = (s > 0) ? __ldg(keys + s - 1) : ~__ldg(keys);   // the very first key always starts a segment
    __syncthreads();

    // ---- my 8 consecutive positions; segment-start flags (a virtual start closes the data at q == nvalid)
    const int q0 = tid * SS_PER;
    uint32_t headbits = 0;
    {
        uint64_t kx[SS_PER];
#pragma unroll
        for (int j = 0; j < SS_PER; j++) kx[j] = sk[ss_idx(q0 + j)];
        uint32_t fb = 0;
        {
            uint64_t prev = tid ? sk[ss_idx(q0 - 1)] : s_halo;
#pragma unroll
            for (int j = 0; j < SS_PER; j++) {
                const int q = q0 + j;
                const bool f = (q < nvalid) ? (((kx[j] ^ prev) >> lowbits) != 0) : (q == nvalid);
                fb |= (f ? 1u : 0u) << j;
                prev = kx[j];
            }
            reinterpret_cast<uint8_t*>(sflag)[tid] = (uint8_t)fb;
        }
        __syncthreads();

        // ---- which positions are mine (their segment starts in my first SS_TILE positions and is short
        // enough); phase A: one hash insert per key -- the first key to claim a value's slot becomes its head,
        // every other key of that value adds itself to the slot's count.
        int st = mask_prev(sflag, q0);
        const int after = mask_next(sflag, q0 + SS_PER);
        int en = after;
        {
            const uint32_t up = fb >> 1;
            if (up) en = q0 + __ffs(up);
        }
#pragma unroll
        for (int j = 0; j < SS_PER; j++) {
            const int q = q0 + j;
            if ((fb >> j) & 1u) {
                st = q;
                const uint32_t up = (j < SS_PER - 1) ? (fb >> (j + 1)) : 0u;
                en = up ? q + __ffs(up) : after;
                if (st < SS_TILE && q < nvalid && en - st > SS_HALO) {   // a big segment that I own: list it
                    const unsigned long long slot = atomicAdd(big_n, 1ull);
                    if (slot < big_cap) big_start[slot] = s + q;
                }
            }
            const bool ok = st >= 0 && st < SS_TILE && (en - st) <= SS_HALO && q < nvalid;
            if (ok && DISTINCT) headbits |= 1u << j;
            if (ok && !DISTINCT) {
                const uint64_t x = kx[j];
                uint32_t h = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> 51);   // 13 bits
                const uint32_t mine = WEIGHTED ? (uint32_t)q : ((uint32_t)q | 0x10000u);
                uint32_t hq;
                while (true) {
                    const uint32_t old = atomicCAS(&table[h], SS_EMPTY, mine);
                    if (old == SS_EMPTY) { headbits |= 1u << j; hq = (uint32_t)q; break; }
                    if (sk[ss_idx((int)(old & 0xffffu))] == x) {
                        hq = old & 0xffffu;
                        if (!WEIGHTED) atomicAdd(&table[h], 0x10000u);
                        break;
                    }
                    h = (h + 1) & (SS_HASH - 1);
                }
                if (WEIGHTED) {
                    const uint32_t wt = sw[ss_idx(q)];
                    const uint32_t old = atomicAdd(&wsum[hq], wt);
                    if (old + wt < old) atomicExch(err, 1u);
                }
            }
        }
    }
    reinterpret_cast<uint8_t*>(shead)[tid] = (uint8_t)headbits;
    uint32_t H;
    const uint32_t hbase = block_excl_scan<SS_THREADS, uint32_t, false>(__popc(headbits), s_scan, &H);
    shbase[tid] = hbase;
    if (tid == 0) tile_heads[tile] = H;
    __syncthreads();

    // ---- phase B: every head ranks itself among the heads of its segment and writes (key, count).
    // Heads are sparse (one key in six for 30x reads), so the heads of a warp's 256 positions are dealt out
    // evenly to its 32 lanes: head number g of the warp goes to lane g % 32.
    const uint64_t base = (uint64_t)tile * SS_LOADED;
    if (DISTINCT) {
        // every key is a head: nothing to deal out, each thread finishes its own 8 positions; the segment's heads are
        // simply its positions
        uint32_t hb = headbits;
        while (hb) {
            const int j = __ffs(hb) - 1;
            hb &= hb - 1;
            const int q = q0 + j;
            const uint64_t x = sk[ss_idx(q)];
            const int st = mask_prev(sflag, q);
            const int en = mask_next(sflag, q + 1);
            const uint32_t hb0 = shbase[st >> 3] + __popc((shead[st >> 5] >> (st & 24)) & ((1u << (st & 7)) - 1u));
            uint32_t r = 0;
            for (int p2 = st; p2 < en; p2++) {
                const uint64_t y = sk[ss_idx(p2)];
                r += (y < x) ? 1u : 0u;
                if (y == x && p2 != q) atomicExch(err, 2u);   // the caller's promise is broken
            }
            tmp_k[base + hb0 + r] = x;
            tmp_c[base + hb0 + r] = __ldg(w + s + q);
        }
        return;
    }
    const unsigned lane = tid & 31;
    const uint32_t hinc = warp_incl_scan<uint32_t>(__popc(headbits));
    const uint32_t Hw = __shfl_sync(0xffffffffu, hinc, 31);
    const int wq0 = (int)(tid & ~31u) * SS_PER;
    for (uint32_t g0 = 0; g0 < Hw; g0 += 32) {
        const uint32_t g = g0 + lane;
        int o = 0;   // owner lane: the first one whose inclusive head count exceeds g
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t v = __shfl_sync(0xffffffffu, hinc, o + step - 1);
            if (v <= g) o += step;
        }
        uint32_t obits = __shfl_sync(0xffffffffu, headbits, o);
        const uint32_t oinc = __shfl_sync(0xffffffffu, hinc, o);
        if (g >= Hw) continue;
        for (uint32_t i = g - (oinc - __popc(obits)); i > 0; i--) obits &= obits - 1;   // my head's bit in the owner's mask
        const int q = wq0 + o * SS_PER + __ffs(obits) - 1;
        const uint64_t x = sk[ss_idx(q)];
        const int st = mask_prev(sflag, q);
        const int en = mask_next(sflag, q + 1);
        const uint32_t hb0 = shbase[st >> 3] + __popc((shead[st >> 5] >> (st & 24)) & ((1u << (st & 7)) - 1u));
        uint32_t r = 0;
        for (int wd = st >> 5; wd <= (en - 1) >> 5; wd++) {
            uint32_t bits = shead[wd];
            if (wd == (st >> 5)) bits &= 0xffffffffu << (st & 31);
            if (wd == (en >> 5)) bits &= (1u << (en & 31)) - 1u;   // en a multiple of 32: wd never reaches it
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t y = sk[ss_idx((wd << 5) + b)];
                r += (y < x) ? 1u : 0u;
                if (DISTINCT && y == x && (wd << 5) + b != q) atomicExch(err, 2u);   // the caller's promise is broken
            }
        }
        uint32_t cnt;
        if (WEIGHTED) {
            cnt = wsum[q];
        } else if (DISTINCT) {
            cnt = __ldg(w + s + q);
        } else {
            uint32_t h = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> 51);
            uint32_t e = table[h];
            while ((e & 0xffffu) != (uint32_t)q) { h = (h + 1) & (SS_HASH - 1); e = table[h]; }
            cnt = e >> 16;
        }
        tmp_k[base + hb0 + r] = x;
        tmp_c[base + hb0 + r] = cnt;
    }
}

// exclusive scan of the per-tile head counts (one CTA; a few ten thousand tiles)
__global__ void __launch_bounds__(1024) segscan_kernel(const uint32_t* __restrict__ tile_heads, uint32_t tiles,
                                                        uint64_t* __restrict__ tile_off, uint64_t* __restrict__ totals) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < tiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = (i < tiles) ? tile_heads[i] : 0;
        uint64_t tot;
        const uint64_t ex = block_excl_scan<1024, uint64_t>(v, sm, &tot);
        if (i < tiles) tile_off[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) totals[0] = carry;
}

__global__ void __launch_bounds__(256)
segcompact_kernel(const uint64_t* __restrict__ tmp_k, const uint32_t* __restrict__ tmp_c,
                  const uint32_t* __restrict__ tile_heads, const uint64_t* __restrict__ tile_off,
                  uint64_t* __restrict__ out_k, uint32_t* __restrict__ out_c) {
    const uint32_t tile = blockIdx.x;
    const uint32_t H = tile_heads[tile];
    const uint64_t src = (uint64_t)tile * SS_LOADED, dst = tile_off[tile];
    for (uint32_t i = threadIdx.x; i < H; i += 256) {
        out_k[dst + i] = tmp_k[src + i];
        out_c[dst + i] = tmp_c[src + i];
    }
}

// first index in [lo, n) whose top bits differ from those of keys[start]
__global__ void big_len_kernel(const uint64_t* __restrict__ keys, uint64_t n, int lowbits,
                               const uint64_t* __restrict__ big_start, uint64_t nbig, uint64_t* __restrict__ big_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbig) return;
    const uint64_t st = big_start[i];
    const uint64_t v = keys[st] >> lowbits;
    uint64_t lo = st + 1, hi = n;   // keys are ordered by their top bits
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((keys[mid] >> lowbits) > v) hi = mid; else lo = mid + 1;
    }
    big_len[i] = lo - st;
}

__global__ void __launch_bounds__(256)
big_gather_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ w, const uint64_t* __restrict__ big_start,
                  const uint64_t* __restrict__ big_off /*[nbig+1] exclusive*/, uint64_t nbig, uint64_t total,
                  uint64_t* __restrict__ bk, uint32_t* __restrict__ bw) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = nbig;   // last segment with off <= i
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (big_off[mid] <= i) lo = mid; else hi = mid;
        }
        const uint64_t src = big_start[lo] + (i - big_off[lo]);
        bk[i] = keys[src];
        if (w) bw[i] = w[src];
    }
}

static size_t sort_count_classic(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits,
                                 uint64_t* out_k, uint32_t* out_c) {
    int which;
    {
        Stage st(c, "sort");
        which = radix_sort(c, k0, k1, v0, v1, n, key_bits);
    }
    Stage st(c, "count");
    return reduce_by_key(c, which ? k1 : k0, v0 ? (which ? v1 : v0) : nullptr, n, out_k, out_c);
}

static size_t sort_count_segsort(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits,
                                 uint64_t* out_k, uint32_t* out_c, bool distinct) {
    if (n == 0) return 0;
    if (distinct && !v0) ZB_FAIL(ZB_E_ARG, "sort_count: distinct mode needs a payload");
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    // top passes: T = 8 P bits with P = ceil((log2 n - 4) / 8)  ->  n / 2^T keys per segment in [1/16, 16)
    int lg = 0;
    while (lg < 63 && ((size_t)1 << lg) < n) lg++;
    int P = (lg - 4 + 7) / 8;
    if (P < 1) P = 1;
    const int T = 8 * P;
    if (key_bits < T + 8) return sort_count_classic(c, k0, k1, v0, v1, n, key_bits, out_k, out_c);
    const int lowbits = key_bits - T;
    const bool weighted = (v0 != nullptr);

    int which;
    {
        Stage st(c, "sort");
        which = radix_sort_range(c, k0, k1, v0, v1, n, lowbits, T, false);
    }
    const uint64_t* sk = which ? k1 : k0;
    const uint32_t* sv = weighted ? (which ? v1 : v0) : nullptr;

    const uint32_t tiles = (uint32_t)div_up(n, SS_TILE);
    const size_t big_cap = n / (SS_HALO + 1) + 2;
    DBuf<uint64_t> big_start(c, big_cap);
    DBuf<uint64_t> tmp_k(c, (size_t)tiles * SS_LOADED);
    DBuf<uint32_t> tmp_c(c, (size_t)tiles * SS_LOADED);
    DBuf<uint32_t> tile_heads(c, tiles);
    DBuf<uint64_t> tile_off(c, (size_t)tiles + 4);
    uint64_t* totals = tile_off.get() + tiles;                                       // [0] distinct, [1] big count
    unsigned int* err = reinterpret_cast<unsigned int*>(totals + 2);
    ZB_CUDA(dev_memset(c, totals, 0, 32));
    const int mode = !weighted ? 0 : (distinct ? 2 : 1);
    const size_t smem = (size_t)SS_PADDED * 8 + (mode == 2 ? 0 : (size_t)SS_HASH * 4) + 2 * (SS_LOADED / 32) * 4 +
                        SS_THREADS * 4 + (mode == 1 ? (size_t)SS_PADDED * 4 + (size_t)SS_LOADED * 4 : 0);
    {
        Stage st(c, "segcount");
        unsigned long long* bign = reinterpret_cast<unsigned long long*>(totals + 1);
        if (mode == 1) {
            ZB_CUDA(cudaFuncSetAttribute(segsort_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            segsort_count_kernel<1><<<tiles, SS_THREADS, smem, c->stream>>>(
                sk, sv, n, lowbits, tmp_k.get(), tmp_c.get(), tile_heads.get(), bign, big_start.get(), big_cap, err);
        } else if (mode == 2) {
            ZB_CUDA(cudaFuncSetAttribute(segsort_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            segsort_count_kernel<2><<<tiles, SS_THREADS, smem, c->stream>>>(
                sk, sv, n, lowbits, tmp_k.get(), tmp_c.get(), tile_heads.get(), bign, big_start.get(), big_cap, err);
        } else {
            ZB_CUDA(cudaFuncSetAttribute(segsort_count_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            segsort_count_kernel<0><<<tiles, SS_THREADS, smem, c->stream>>>(
                sk, nullptr, n, lowbits, tmp_k.get(), tmp_c.get(), tile_heads.get(), bign, big_start.get(), big_cap, err);
        }
        ZB_LAUNCH_CHECK(c);
        segscan_kernel<<<1, 1024, 0, c->stream>>>(tile_heads.get(), tiles, tile_off.get(), totals);
        ZB_LAUNCH_CHECK(c);
        segcompact_kernel<<<tiles, 256, 0, c->stream>>>(tmp_k.get(), tmp_c.get(), tile_heads.get(), tile_off.get(), out_k, out_c);
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(read_back(c, totals, 32));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    size_t n_out = (size_t)c->h_scalars[0];
    const size_t nbig = (size_t)c->h_scalars[1];
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 2)[0] == 2)
        ZB_FAIL(ZB_E_ARG, "sort_count: keys promised to be distinct are not");
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 2)[0] != 0)
        ZB_FAIL(ZB_E_RANGE, "k-mer count exceeds 2^32-1 (reference: array('I') OverflowError, kmerize.py:374)");
    if (nbig == 0) return n_out;
    if (nbig > big_cap) ZB_FAIL(ZB_E_CUDA, "sort_count: big-segment list overflow (%zu > %zu)", nbig, big_cap);

    // ---- big segments: gather, classic sort + count, merge back
    Stage st_big(c, "segcount_big");
    struct ProfileOff {   // the nested classic sort must not add its (tiny) passes to the per-stage report
        Ctx* c; bool saved;
        explicit ProfileOff(Ctx* c_) : c(c_), saved(c_->profile) { c->profile = false; }
        ~ProfileOff() { c->profile = saved; }
    } profile_off(c);
    DBuf<uint64_t> big_len(c, nbig);
    big_len_kernel<<<(unsigned)div_up(nbig, 128), 128, 0, c->stream>>>(sk, n, lowbits, big_start.get(), nbig, big_len.get());
    ZB_LAUNCH_CHECK(c);
    std::vector<uint64_t> off(nbig + 1);
    ZB_CUDA(cudaMemcpyAsync(off.data() + 1, big_len.get(), nbig * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    off[0] = 0;
    for (size_t i = 1; i <= nbig; i++) off[i] += off[i - 1];
    const size_t nb = (size_t)off[nbig];
    DBuf<uint64_t> d_off(c, nbig + 1);
    ZB_CUDA(cudaMemcpyAsync(d_off.get(), off.data(), (nbig + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    DBuf<uint64_t> b0(c, nb), b1(c, nb), bk(c, nb);
    DBuf<uint32_t> w0, w1, bc(c, nb);
    if (weighted) { w0.alloc(c, nb); w1.alloc(c, nb); }
    {
        const int blocks = (int)std::min<size_t>((size_t)c->sm_count * 8, div_up(nb, 256));
        big_gather_kernel<<<blocks, 256, 0, c->stream>>>(sk, sv, big_start.get(), d_off.get(), nbig, nb, b0.get(),
                                                         weighted ? w0.get() : nullptr);
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // `off` (pageable host memory) must stay alive until the copy is done
    const int bw = radix_sort(c, b0.get(), b1.get(), weighted ? w0.get() : nullptr, weighted ? w1.get() : nullptr, nb, key_bits);
    const size_t nbd = reduce_by_key(c, bw ? b1.get() : b0.get(), weighted ? (bw ? w1.get() : w0.get()) : nullptr, nb,
                                     bk.get(), bc.get());
    DBuf<uint64_t> mk(c, n_out + nbd);
    DBuf<uint32_t> mc(c, n_out + nbd);
    merge_pairs(c, out_k, out_c, n_out, bk.get(), bc.get(), nbd, mk.get(), mc.get());
    n_out += nbd;
    ZB_CUDA(dev_copy(c, out_k, mk.get(), n_out * 8));
    ZB_CUDA(dev_copy(c, out_c, mc.get(), n_out * 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return n_out;
}

// =============================================================================== bucket route
// The same result with one LSD pass fewer.  LSD passes run over the top cb bits only, cb chosen so that a "bucket"
// (the keys sharing those bits, contiguous afterwards) holds ~1000-2000 keys: 16 bits = TWO 8-bit passes for the
// 125 M keys of one bench step, where the segment route needs 24 bits = three.  bc_bounds_kernel finds the bucket
// boundaries by bisection; bucket_count_kernel gives every bucket to one CTA, which dedupes the bucket through a
// shared-memory hash table (the first key of a value is its head and collects the count), orders the heads by the
// next 11 key bits with a shared-memory counting sort and by the whole key inside those (mostly one-key) groups, and
// writes the bucket's distinct run; scan + compaction place the runs.  A bucket that does not fit (skewed key
// space, one value repeated thousands of times) is only listed: those keys are gathered and go through the segment
// route above, and the two results are merged.
static constexpr int BC_THREADS = 512;
static constexpr int BC_PER = 8;
static constexpr int BC_CAP = BC_THREADS * BC_PER;      // keys of one bucket
static constexpr int BC_HASH = 2 * BC_CAP;
static constexpr int BC_HASH_BITS = 13;
static constexpr int BC_FINE = 2048;
static constexpr int BC_FINE_BITS = 11;

// start[b] = first index whose top bits are >= b (b = 0 .. nb); the keys are ordered by those bits
__global__ void __launch_bounds__(256)
bc_bounds_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t nb, uint64_t* __restrict__ start) {
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    if (b > nb) return;
    uint64_t lo = 0, hi = n;
    if (b == nb) {
        lo = n;
    } else if (b > 0) {
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((keys[mid] >> shift) < b) lo = mid + 1; else hi = mid;
        }
    }
    start[b] = lo;
}

// MODE 0: count keys; MODE 2: keys are distinct and carry a u32 payload (the mirror sort of kmerize).
// Shared memory: keys 2 x 32 KB + hash table 32 KB + group sizes 8 KB = 104 KB -> 2 CTAs per SM.
// A bucket of up to 16 x BC_CAP keys is done in 2, 4, 8 or 16 rounds, one per value of the key bits right below the
// bucket prefix (canonical k-mers are not spread evenly: prefixes starting with A hold 7/16 of them, with T 1/16, so
// the fullest buckets are 1.75 x the average; real genomes are far more skewed).  A round compacts its share of the
// bucket into shared memory and proceeds like a small bucket; only a bucket whose share of one round still does not
// fit is left to the caller (listed in big_list).
// Persistent: a CTA takes buckets blockIdx.x, + gridDim.x, ... and, while it works on one, the TMA engine copies
// the keys of its next one into the other half of a double buffer (one cp.async.bulk per bucket, completion on an
// mbarrier) -- the first version loaded a bucket with plain loads and then spent 34 % of its time waiting for them.
template <int MODE>
__global__ void __launch_bounds__(BC_THREADS, 2)
bucket_count_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ w, const uint64_t* __restrict__ start,
                    uint32_t nb, int shift, int fine_shift, uint32_t fine_mask, uint64_t* __restrict__ tmp_k,
                    uint32_t* __restrict__ tmp_c, uint32_t* __restrict__ tile_heads, unsigned long long* __restrict__ big_n,
                    uint32_t* __restrict__ big_list, unsigned int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char bc_raw[];
    uint64_t* buf0 = reinterpret_cast<uint64_t*>(bc_raw);                // 2 x [BC_CAP + 2] keys of this / the next bucket
    // slot = position of the value's head (low 16 bits) | number of keys with that value (high 16 bits)
    uint32_t* table = reinterpret_cast<uint32_t*>(buf0 + 2 * (BC_CAP + 2));   // [BC_HASH]
    uint32_t* hist = table + BC_HASH;                                    // [BC_FINE + 1]
    uint64_t* hs = reinterpret_cast<uint64_t*>(table);                   // [BC_CAP] heads grouped by fine digit (after the dedupe)
    __shared__ uint32_t s_scan[BC_THREADS / 32 + 1];
    __shared__ uint32_t s_sub[16];
    __shared__ uint32_t s_cnt;
    __shared__ __align__(8) uint64_t s_bar[2];
    constexpr bool DISTINCT = (MODE == 2);
    // groups of the heads' counting sort: ~300 heads per bucket when keys repeat 6.5 x (MODE 0), every key a head in MODE 2
    constexpr int FINE = DISTINCT ? BC_FINE : BC_FINE / 4;

    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    uint32_t phases = 0;   // bit i = parity the next wait on s_bar[i] expects
    uint32_t b = blockIdx.x;
    uint64_t s0 = 0, mm = 0;
    if (b < nb) { s0 = __ldg(start + b); mm = __ldg(start + b + 1) - s0; }
    __syncthreads();
    // keys of a bucket as the bulk copy takes them: from the 16-byte boundary at or below the first key, an even number
    // of keys (an odd last key is fetched by a plain load)
#define BC_SKEW(s0_) ((uint32_t)(((uintptr_t)(keys + (s0_))) >> 3) & 1u)
#define BC_NCOPY(s0_, mm_) (((mm_) == 0 || (mm_) > (uint64_t)BC_CAP) ? 0u : ((BC_SKEW(s0_) + (uint32_t)(mm_)) & ~1u))
    if (tid == 0 && BC_NCOPY(s0, mm)) {
        mbar_expect_tx(&s_bar[0], BC_NCOPY(s0, mm) * 8);
        bulk_g2s(buf0, keys + s0 - BC_SKEW(s0), BC_NCOPY(s0, mm) * 8, &s_bar[0]);
    }
    for (uint32_t it = 0; b < nb; b += gridDim.x, it++) {
        const int cur = (int)(it & 1);
        uint64_t* bufc = buf0 + cur * (BC_CAP + 2);
        const uint32_t bn = b + gridDim.x;
        uint64_t s0n = 0, mmn = 0;
        if (bn < nb) { s0n = __ldg(start + bn); mmn = __ldg(start + bn + 1) - s0n; }
        __syncthreads();   // everybody is done with the previous bucket (its keys / counts lived in the other buffer)
        if (tid == 0 && BC_NCOPY(s0n, mmn)) {
            uint64_t* bufn = buf0 + (cur ^ 1) * (BC_CAP + 2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last written by ordinary stores
            mbar_expect_tx(&s_bar[cur ^ 1], BC_NCOPY(s0n, mmn) * 8);
            bulk_g2s(bufn, keys + s0n - BC_SKEW(s0n), BC_NCOPY(s0n, mmn) * 8, &s_bar[cur ^ 1]);
        }
        const uint32_t ncopy = BC_NCOPY(s0, mm);
        const uint64_t* sk = bufc + BC_SKEW(s0);                             // the bucket's keys (when it fits)
        uint32_t* hc = reinterpret_cast<uint32_t*>(bufc);                    // [BC_CAP] counts of the grouped heads (after the dedupe)

        int rbits = 0;
        bool give_up = mm > (uint64_t)BC_CAP * 16 || (mm > (uint64_t)BC_CAP && shift < 4);
        if (mm > (uint64_t)BC_CAP && !give_up) {
            // share of every value of the next 4 key bits; the fewest rounds whose shares all fit
            if (tid < 16) s_sub[tid] = 0;
            __syncthreads();
            for (uint64_t i = tid; i < mm; i += BC_THREADS)
                atomicAdd(&s_sub[(uint32_t)(__ldg(keys + s0 + i) >> (shift - 4)) & 15u], 1u);
            __syncthreads();
            for (rbits = 1; rbits <= 4; rbits++) {
                const int per = 16 >> rbits;
                bool fits = true;
                for (int r = 0; r < (1 << rbits); r++) {
                    uint32_t t = 0;
                    for (int u = 0; u < per; u++) t += s_sub[r * per + u];
                    fits &= t <= (uint32_t)BC_CAP;
                }
                if (fits) break;
            }
            give_up = rbits > 4;
        }
        if (mm == 0 || give_up) {
            if (tid == 0) {
                tile_heads[b] = 0;
                if (mm) big_list[atomicAdd(big_n, 1ull)] = b;
            }
            s0 = s0n;
            mm = mmn;
            continue;
        }
        uint32_t out_base = 0;
        for (int round = 0; round < (1 << rbits); round++) {
            if (round) __syncthreads();   // the previous round is done with the shared arrays
            if (!DISTINCT) {
                const uint4 e4 = make_uint4(SS_EMPTY, SS_EMPTY, SS_EMPTY, SS_EMPTY);
#pragma unroll
                for (int j = 0; j < BC_HASH / 4 / BC_THREADS; j++) reinterpret_cast<uint4*>(table)[j * BC_THREADS + tid] = e4;
            }
#pragma unroll
            for (int j = 0; j < FINE / BC_THREADS; j++) hist[j * BC_THREADS + tid] = 0;
            uint64_t kx[BC_PER];
            uint32_t wx[BC_PER];
            int m;
            if (rbits == 0) {
                m = (int)mm;
                if (DISTINCT) {   // the payload comes by plain loads, issued before the wait for the keys
#pragma unroll
                    for (int j = 0; j < BC_PER; j++) {
                        const int q = j * BC_THREADS + (int)tid;
                        wx[j] = (q < m) ? __ldg(w + s0 + q) : 0u;
                    }
                }
                if (tid == 0 && ((BC_SKEW(s0) + (uint32_t)m) & 1u)) bufc[BC_SKEW(s0) + m - 1] = __ldg(keys + s0 + m - 1);
                if (ncopy) {
                    mbar_wait(&s_bar[cur], (phases >> cur) & 1u);
                    phases ^= 1u << cur;
                }
                __syncthreads();   // table / hist initialised, the odd last key stored
#pragma unroll
                for (int j = 0; j < BC_PER; j++) {
                    const int q = j * BC_THREADS + (int)tid;
                    kx[j] = (q < m) ? sk[q] : 0ull;
                    if ((j + 1) * BC_THREADS >= m) break;
                }
            } else {
                // compact this round's keys (any order) into the buffer; DISTINCT: their positions into `table` for the payload
                sk = bufc;
                if (tid == 0) s_cnt = 0;
                __syncthreads();
                const uint32_t rmask = (1u << rbits) - 1u;
                for (uint64_t i0 = 0; i0 < mm; i0 += BC_THREADS) {
                    const uint64_t i = i0 + tid;
                    const uint64_t x = (i < mm) ? __ldg(keys + s0 + i) : 0ull;
                    const bool sel = (i < mm) && (((uint32_t)(x >> (shift - rbits)) & rmask) == (uint32_t)round);
                    const unsigned bal = __ballot_sync(0xffffffffu, sel);
                    uint32_t base = 0;
                    if (lane_id() == 0 && bal) base = atomicAdd(&s_cnt, (uint32_t)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (sel) {
                        const uint32_t slot = base + __popc(bal & lanemask_lt());
                        bufc[slot] = x;
                        if (DISTINCT) table[slot] = (uint32_t)i;
                    }
                }
                __syncthreads();
                m = (int)s_cnt;
#pragma unroll
                for (int j = 0; j < BC_PER; j++) {
                    const int q = j * BC_THREADS + (int)tid;
                    kx[j] = (q < m) ? sk[q] : 0ull;
                    if (DISTINCT) wx[j] = (q < m) ? __ldg(w + s0 + table[q]) : 0u;
                }
                if (DISTINCT) __syncthreads();   // `table` is about to be reused for the grouped heads
            }

            // item loops below stop at the last slice of 512 positions that holds keys (a bucket is half full on average)
            const int jn = (m + BC_THREADS - 1) / BC_THREADS;
            uint32_t headbits = 0;
            if (DISTINCT) {
#pragma unroll
                for (int j = 0; j < BC_PER; j++)
                    if (j * BC_THREADS + (int)tid < m) headbits |= 1u << j;
            } else {
                // ---- dedupe: one hash insert per key; the first key to claim a value's slot is its head
#pragma unroll
                for (int j = 0; j < BC_PER; j++) {
                    if (j >= jn) break;
                    const int q = j * BC_THREADS + (int)tid;
                    if (q < m) {
                        const uint64_t x = kx[j];
                        uint32_t h = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> (64 - BC_HASH_BITS));
                        while (true) {
                            const uint32_t old = atomicCAS(&table[h], SS_EMPTY, (uint32_t)q | 0x10000u);
                            if (old == SS_EMPTY) { headbits |= 1u << j; wx[j] = h; break; }
                            if (sk[old & 0xffffu] == x) { atomicAdd(&table[h], 0x10000u); break; }
                            h = (h + 1) & (BC_HASH - 1);
                        }
                    }
                }
                __syncthreads();
            }

            // ---- heads: counting sort by the next key bits (rank inside a group = arrival order, fixed up below)
            uint32_t rd[BC_PER];
#pragma unroll
            for (int j = 0; j < BC_PER; j++) {
                if (j >= jn) break;
                if ((headbits >> j) & 1u) {
                    const uint32_t d = (uint32_t)(kx[j] >> fine_shift) & fine_mask;
                    rd[j] = atomicAdd(&hist[d], 1u) | (d << 16);
                    if (!DISTINCT) wx[j] = table[wx[j]] >> 16;
                }
            }
            __syncthreads();
            {   // exclusive scan of the group sizes
                constexpr int GP = FINE / BC_THREADS;
                uint32_t v[GP];
                uint32_t t = 0;
#pragma unroll
                for (int u = 0; u < GP; u++) { v[u] = hist[tid * GP + u]; t += v[u]; }
                uint32_t all;
                uint32_t ex = block_excl_scan<BC_THREADS, uint32_t, false>(t, s_scan, &all);
#pragma unroll
                for (int u = 0; u < GP; u++) { hist[tid * GP + u] = ex; ex += v[u]; }
                if (tid == 0) hist[FINE] = all;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < BC_PER; j++) {
                if (j >= jn) break;
                if ((headbits >> j) & 1u) {
                    const uint32_t p = hist[rd[j] >> 16] + (rd[j] & 0xffffu);
                    hs[p] = kx[j];
                    hc[p] = wx[j];
                }
            }
            __syncthreads();

            // ---- every head ranks itself inside its group and writes (key, count)
            const int H = (int)hist[FINE];
            for (int p = (int)tid; p < H; p += BC_THREADS) {
                const uint64_t x = hs[p];
                const uint32_t d = (uint32_t)(x >> fine_shift) & fine_mask;
                const int g0 = (int)hist[d], g1 = (int)hist[d + 1];
                int r = 0;
                if (g1 - g0 > 1) {
                    for (int p2 = g0; p2 < g1; p2++) {
                        const uint64_t y = hs[p2];
                        r += (y < x) ? 1 : 0;
                        if (DISTINCT && y == x && p2 != p) atomicExch(err, 2u);   // the caller's promise is broken
                    }
                }
                tmp_k[s0 + out_base + g0 + r] = x;
                tmp_c[s0 + out_base + g0 + r] = hc[p];
            }
            out_base += (uint32_t)H;
        }
        if (tid == 0) tile_heads[b] = out_base;
        s0 = s0n;
        mm = mmn;
    }
#undef BC_SKEW
#undef BC_NCOPY
}

// one warp per bucket: its run of H distinct keys moves from the bucket's own offset to its final place
__global__ void __launch_bounds__(256)
bucket_compact_kernel(const uint64_t* __restrict__ tmp_k, const uint32_t* __restrict__ tmp_c, const uint32_t* __restrict__ tile_heads,
                      const uint64_t* __restrict__ start, const uint64_t* __restrict__ tile_off, uint32_t nb,
                      uint64_t* __restrict__ out_k, uint32_t* __restrict__ out_c) {
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= nb) return;
    const uint32_t H = tile_heads[b];
    const uint64_t src = start[b], dst = tile_off[b];
    for (uint32_t i = lane_id(); i < H; i += 32) {
        out_k[dst + i] = tmp_k[src + i];
        out_c[dst + i] = tmp_c[src + i];
    }
}

__global__ void big_bucket_ranges_kernel(const uint32_t* __restrict__ big_list, uint64_t nbig, const uint64_t* __restrict__ start,
                                         uint64_t* __restrict__ big_start, uint64_t* __restrict__ big_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbig) return;
    const uint32_t b = big_list[i];
    big_start[i] = start[b];
    big_len[i] = start[b + 1] - start[b];
}

// top key bits the LSD passes of the bucket route put in order: buckets of at most ~BC_CAP / 2 keys on average
static int bucket_top_bits(size_t n, int key_bits) {
    int cb = 0;
    while (cb < key_bits && (n >> cb) > (size_t)BC_CAP / 2) cb++;
    return cb;
}

bool sort_count_plan(size_t n, int key_bits, SortPre* plan) {
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    const int cb = bucket_top_bits(n, key_bits);
    if (cb == 0 || g_sort_max_bits > 8 || g_sort_count_mode != 0) return false;
    plan->passes = sort_digit_plan(key_bits - cb, cb, 8, plan->shift, plan->bits, 4);
    return plan->passes > 0;
}

static size_t sort_count_buckets(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits,
                                 uint64_t* out_k, uint32_t* out_c, bool distinct, const SortPre* pre) {
    const bool weighted = (v0 != nullptr);
    const int cb = bucket_top_bits(n, key_bits);
    const uint32_t nb = 1u << cb;
    const int shift = key_bits - cb;
    int which = 0;
    if (cb > 0) {
        Stage st(c, "sort");
        which = radix_sort_range(c, k0, k1, v0, v1, n, shift, cb, false, pre);
    }
    const uint64_t* sk = which ? k1 : k0;
    const uint32_t* sv = weighted ? (which ? v1 : v0) : nullptr;
    const int fb = std::min(distinct ? BC_FINE_BITS : BC_FINE_BITS - 2, shift);   // 2048 / 512 groups, see the kernel
    const int fine_shift = shift - fb;
    const uint32_t fine_mask = (1u << fb) - 1u;

    DBuf<uint64_t> start(c, (size_t)nb + 2);
    DBuf<uint64_t> tmp_k(c, n);
    DBuf<uint32_t> tmp_c(c, n);
    DBuf<uint32_t> tile_heads(c, nb);
    DBuf<uint32_t> big_list(c, nb);
    DBuf<uint64_t> tile_off(c, (size_t)nb + 4);
    uint64_t* totals = tile_off.get() + nb;                                          // [0] distinct, [1] big buckets
    unsigned int* err = reinterpret_cast<unsigned int*>(totals + 2);
    ZB_CUDA(dev_memset(c, totals, 0, 32));
    const size_t smem = (size_t)2 * (BC_CAP + 2) * 8 + (size_t)BC_HASH * 4 + (size_t)(BC_FINE + 1) * 4;
    const unsigned grid = (unsigned)std::min<size_t>(nb, (size_t)c->sm_count * 2);   // persistent: 2 CTAs per SM
    {
        Stage st(c, "segcount");
        bc_bounds_kernel<<<(unsigned)div_up((size_t)nb + 1, 256), 256, 0, c->stream>>>(sk, n, shift, nb, start.get());
        ZB_LAUNCH_CHECK(c);
        unsigned long long* bign = reinterpret_cast<unsigned long long*>(totals + 1);
        if (distinct) {
            ZB_CUDA(cudaFuncSetAttribute(bucket_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bucket_count_kernel<2><<<grid, BC_THREADS, smem, c->stream>>>(sk, sv, start.get(), nb, shift, fine_shift, fine_mask, tmp_k.get(),
                                                                        tmp_c.get(), tile_heads.get(), bign, big_list.get(), err);
        } else {
            ZB_CUDA(cudaFuncSetAttribute(bucket_count_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bucket_count_kernel<0><<<grid, BC_THREADS, smem, c->stream>>>(sk, nullptr, start.get(), nb, shift, fine_shift, fine_mask, tmp_k.get(),
                                                                        tmp_c.get(), tile_heads.get(), bign, big_list.get(), err);
        }
        ZB_LAUNCH_CHECK(c);
        segscan_kernel<<<1, 1024, 0, c->stream>>>(tile_heads.get(), nb, tile_off.get(), totals);
        ZB_LAUNCH_CHECK(c);
        bucket_compact_kernel<<<(unsigned)div_up(nb, 8), 256, 0, c->stream>>>(tmp_k.get(), tmp_c.get(), tile_heads.get(), start.get(),
                                                                             tile_off.get(), nb, out_k, out_c);
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(read_back(c, totals, 32));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    size_t n_out = (size_t)c->h_scalars[0];
    const size_t nbig = (size_t)c->h_scalars[1];
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 2)[0] == 2)
        ZB_FAIL(ZB_E_ARG, "sort_count: keys promised to be distinct are not");
    if (nbig == 0) return n_out;

    // ---- buckets that did not fit: gather their keys, segment route, merge back
    Stage st_big(c, "segcount_big");
    DBuf<uint64_t> big_start(c, nbig), big_len(c, nbig);
    big_bucket_ranges_kernel<<<(unsigned)div_up(nbig, 128), 128, 0, c->stream>>>(big_list.get(), nbig, start.get(), big_start.get(),
                                                                               big_len.get());
    ZB_LAUNCH_CHECK(c);
    std::vector<uint64_t> off(nbig + 1);
    ZB_CUDA(cudaMemcpyAsync(off.data() + 1, big_len.get(), nbig * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    off[0] = 0;
    for (size_t i = 1; i <= nbig; i++) off[i] += off[i - 1];
    const size_t nbk = (size_t)off[nbig];
    DBuf<uint64_t> d_off(c, nbig + 1);
    ZB_CUDA(cudaMemcpyAsync(d_off.get(), off.data(), (nbig + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    DBuf<uint64_t> b0(c, nbk), b1(c, nbk), bk(c, nbk);
    DBuf<uint32_t> w0, w1, bc(c, nbk);
    if (weighted) { w0.alloc(c, nbk); w1.alloc(c, nbk); }
    {
        const int blocks = (int)std::min<size_t>((size_t)c->sm_count * 8, div_up(nbk, 256));
        big_gather_kernel<<<blocks, 256, 0, c->stream>>>(sk, sv, big_start.get(), d_off.get(), nbig, nbk, b0.get(),
                                                         weighted ? w0.get() : nullptr);
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // `off` (pageable host memory) must stay alive until the copy is done
    const size_t nbd = sort_count_segsort(c, b0.get(), b1.get(), weighted ? w0.get() : nullptr, weighted ? w1.get() : nullptr, nbk,
                                          key_bits, bk.get(), bc.get(), distinct);
    DBuf<uint64_t> mk(c, n_out + nbd);
    DBuf<uint32_t> mc(c, n_out + nbd);
    merge_pairs(c, out_k, out_c, n_out, bk.get(), bc.get(), nbd, mk.get(), mc.get());
    n_out += nbd;
    ZB_CUDA(dev_copy(c, out_k, mk.get(), n_out * 8));
    ZB_CUDA(dev_copy(c, out_c, mc.get(), n_out * 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return n_out;
}

// =============================================================================== mirrored half, fused
// The last step of kmerize: the both-strand set is the canonical counted set C (sorted) united with its reverse
// complements M (mirror_keys: distinct, disjoint from C, in no order).  Instead of sorting M completely, compacting it
// and merge-pathing it with C (three trips through HBM), M is ordered by its top cb bits only (LSD passes), and one CTA
// per key-range bucket then has everything in shared memory: the bucket's slice of C (already sorted) and the bucket of
// M, which it counting-sorts by the next key bits.  An element of M lands at (its rank in M's bucket) + (keys of the C
// slice below it, by bisection); an element of C at (its index in the slice) + (keys of M's bucket below it: the
// groups before its own, plus a look at its own group).  The output position of the bucket is known beforehand --
// nothing is deduplicated -- so the result is written in place, without staging or compaction.
// Two shapes (g_mm_cfg / ZB_MM_CFG): 0 = 256 threads, buckets of <= 2048 entries, 1024 groups (48 KB: 4 CTAs per SM);
// 1 = 512 threads, <= 4096 entries, 2048 groups (98 KB: 2 CTAs per SM); 2 = shape 0 with 2048 groups (half as many keys
// per group to rank against, 57 KB: 3 CTAs per SM; not measured yet).
int g_mm_cfg = 0;

template <int THREADS, int PER, int FINE_BITS>
__global__ void __launch_bounds__(THREADS, (THREADS == 256) ? 4 : 2)
mirror_merge_kernel(const uint64_t* __restrict__ ck, const uint32_t* __restrict__ cc, const uint64_t* __restrict__ startC,
                    const uint64_t* __restrict__ mk, const uint32_t* __restrict__ mc, const uint64_t* __restrict__ startM,
                    int fine_shift, uint32_t fine_mask, uint64_t* __restrict__ out_k, uint32_t* __restrict__ out_c,
                    unsigned int* __restrict__ err) {
    constexpr int CAP = THREADS * PER;
    constexpr int FINE = 1 << FINE_BITS;
    constexpr int GP = FINE / THREADS;
    extern __shared__ __align__(16) unsigned char mm_raw[];
    uint64_t* sCk = reinterpret_cast<uint64_t*>(mm_raw);                 // [CAP] the C slice (sorted)
    uint64_t* hs = sCk + CAP;                                            // [CAP] M's bucket grouped by fine digit
    uint32_t* hc = reinterpret_cast<uint32_t*>(hs + CAP);                // [CAP] its counts
    uint32_t* hist = hc + CAP;                                           // [FINE + 1] group starts of M's bucket
    uint32_t* cstart = hist + FINE + 1;                                  // [FINE + 1] group starts of the C slice
    __shared__ uint32_t s_scan[THREADS / 32 + 1];

    const unsigned tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint64_t c0 = startC[b], m0 = startM[b];
    const uint64_t nC64 = startC[b + 1] - c0, nM64 = startM[b + 1] - m0;
    if (nC64 + nM64 > (uint64_t)CAP) {
        if (tid == 0) atomicExch(err, 1u);      // the caller falls back to sort + merge
        return;
    }
    const int nC = (int)nC64, nM = (int)nM64;
    if (nC + nM == 0) return;
    const uint64_t o0 = c0 + m0;
    uint64_t kx[PER];
    uint32_t wx[PER], rd[PER];
    const int jn = (nM + THREADS - 1) / THREADS;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (j >= jn) break;
        const int q = j * THREADS + (int)tid;
        kx[j] = (q < nM) ? __ldg(mk + m0 + q) : 0ull;
        wx[j] = (q < nM) ? __ldg(mc + m0 + q) : 0u;
    }
    for (int i = (int)tid; i < nC; i += THREADS) sCk[i] = __ldg(ck + c0 + i);
#pragma unroll
    for (int j = 0; j < GP; j++) hist[j * THREADS + tid] = 0;
    __syncthreads();
    // ---- C: group starts by the next key bits (the slice is sorted: element i opens every group between its
    //      predecessor's and its own; the groups after the last element start at nC)
    for (int i = (int)tid; i < nC; i += THREADS) {
        const int d = (int)((uint32_t)(sCk[i] >> fine_shift) & fine_mask);
        const int dp = (i > 0) ? (int)((uint32_t)(sCk[i - 1] >> fine_shift) & fine_mask) : -1;
        for (int g = dp + 1; g <= d; g++) cstart[g] = (uint32_t)i;
    }
    {
        const int dl = (nC > 0) ? (int)((uint32_t)(sCk[nC - 1] >> fine_shift) & fine_mask) : -1;
        for (int g = dl + 1 + (int)tid; g <= FINE; g += THREADS) cstart[g] = (uint32_t)nC;
    }
    // ---- M: counting sort by the same bits
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (j >= jn) break;
        if (j * THREADS + (int)tid < nM) {
            const uint32_t d = (uint32_t)(kx[j] >> fine_shift) & fine_mask;
            rd[j] = atomicAdd(&hist[d], 1u) | (d << 16);
        }
    }
    __syncthreads();
    {
        uint32_t v[GP];
        uint32_t t = 0;
#pragma unroll
        for (int u = 0; u < GP; u++) { v[u] = hist[tid * GP + u]; t += v[u]; }
        uint32_t all;
        uint32_t ex = block_excl_scan<THREADS, uint32_t, false>(t, s_scan, &all);
#pragma unroll
        for (int u = 0; u < GP; u++) { hist[tid * GP + u] = ex; ex += v[u]; }
        if (tid == 0) hist[FINE] = all;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (j >= jn) break;
        if (j * THREADS + (int)tid < nM) {
            const uint32_t p = hist[rd[j] >> 16] + (rd[j] & 0xffffu);
            hs[p] = kx[j];
            hc[p] = wx[j];
        }
    }
    __syncthreads();
    // ---- C elements: index in the slice + keys of M's bucket below (whole groups before mine, then my own group)
    for (int i = (int)tid; i < nC; i += THREADS) {
        const uint32_t w = __ldg(cc + c0 + i);
        const uint64_t x = sCk[i];
        const uint32_t d = (uint32_t)(x >> fine_shift) & fine_mask;
        const int g0 = (int)hist[d], g1 = (int)hist[d + 1];
        int r = g0;
        for (int p2 = g0; p2 < g1; p2++) r += (hs[p2] < x) ? 1 : 0;
        out_k[o0 + i + r] = x;
        out_c[o0 + i + r] = w;
    }
    // ---- M elements (in group order): rank inside the group + keys of the C slice below, the same way
    for (int p = (int)tid; p < nM; p += THREADS) {
        const uint64_t x = hs[p];
        const uint32_t d = (uint32_t)(x >> fine_shift) & fine_mask;
        const int g0 = (int)hist[d], g1 = (int)hist[d + 1];
        int r = g0;
        for (int p2 = g0; p2 < g1; p2++) r += (hs[p2] < x) ? 1 : 0;
        const int q0 = (int)cstart[d], q1 = (int)cstart[d + 1];
        int lo = q0;
        for (int q = q0; q < q1; q++) lo += (sCk[q] < x) ? 1 : 0;
        out_k[o0 + r + lo] = x;
        out_c[o0 + r + lo] = hc[p];
    }
}

template <int THREADS, int PER, int FINE_BITS>
static void launch_mirror_merge(Ctx* c, uint32_t nb, const uint64_t* ck, const uint32_t* cc, const uint64_t* startC,
                                const uint64_t* sk, const uint32_t* sv, const uint64_t* startM, int shift, uint64_t* out_k,
                                uint32_t* out_c, unsigned int* err) {
    constexpr int CAP = THREADS * PER;
    const int fb = std::min(FINE_BITS, shift);
    const size_t smem = (size_t)CAP * (8 + 8 + 4) + 2 * (size_t)((1 << FINE_BITS) + 1) * 4;
    auto kern = mirror_merge_kernel<THREADS, PER, FINE_BITS>;
    ZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // same value on every call
    kern<<<nb, THREADS, smem, c->stream>>>(ck, cc, startC, sk, sv, startM, shift - fb, (1u << fb) - 1u, out_k, out_c, err);
}

// out = C united with M (nm mirrored pairs in mk/mc, destroyed; mk2/mc2 are ping-pong scratch).  Returns false when a
// bucket does not fit (skewed key space): the caller sorts M and merges instead; M is then still complete in (*mk_out).
bool merge_mirrored(Ctx* c, const uint64_t* ck, const uint32_t* cc, size_t n, uint64_t* mk, uint64_t* mk2, uint32_t* mc,
                    uint32_t* mc2, size_t nm, int key_bits, uint64_t* out_k, uint32_t* out_c, int* which_out) {
    *which_out = 0;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    const size_t cap = g_mm_cfg == 1 ? 4096 : 2048;
    int cb = 0;
    while (cb < key_bits && ((n + nm) >> cb) > cap * 2 / 3) cb++;   // C is A-heavy where M is T-heavy: their sum is even
    if (cb > 24) return false;
    const uint32_t nb = 1u << cb;
    const int shift = key_bits - cb;
    int which = 0;
    if (cb > 0 && nm > 0) {
        Stage st(c, "sort");
        which = radix_sort_range(c, mk, mk2, mc, mc2, nm, shift, cb, false);
    }
    *which_out = which;
    const uint64_t* sk = which ? mk2 : mk;
    const uint32_t* sv = which ? mc2 : mc;
    DBuf<uint64_t> starts(c, 2 * ((size_t)nb + 1) + 2);
    uint64_t* startC = starts.get();
    uint64_t* startM = starts.get() + nb + 1;
    unsigned int* err = reinterpret_cast<unsigned int*>(starts.get() + 2 * ((size_t)nb + 1));
    ZB_CUDA(dev_memset(c, err, 0, 8));
    Stage st(c, "mirror_buckets");
    bc_bounds_kernel<<<(unsigned)div_up((size_t)nb + 1, 256), 256, 0, c->stream>>>(ck, n, shift, nb, startC);
    ZB_LAUNCH_CHECK(c);
    bc_bounds_kernel<<<(unsigned)div_up((size_t)nb + 1, 256), 256, 0, c->stream>>>(sk, nm, shift, nb, startM);
    ZB_LAUNCH_CHECK(c);
    if (g_mm_cfg == 1) launch_mirror_merge<512, 8, 11>(c, nb, ck, cc, startC, sk, sv, startM, shift, out_k, out_c, err);
    else if (g_mm_cfg == 2) launch_mirror_merge<256, 8, 11>(c, nb, ck, cc, startC, sk, sv, startM, shift, out_k, out_c, err);
    else launch_mirror_merge<256, 8, 10>(c, nb, ck, cc, startC, sk, sv, startM, shift, out_k, out_c, err);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, err, 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return reinterpret_cast<uint32_t*>(c->h_scalars)[0] == 0;
}

// g_sort_count_mode (ZB_SORT_COUNT): 0 = bucket route (weighted sums: segment route), 1 = classic full sort +
// reduce-by-key, 2 = segment route
size_t sort_count(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits,
                  uint64_t* out_k, uint32_t* out_c, bool distinct, const SortPre* pre) {
    if (n == 0) return 0;
    if (distinct && !v0) ZB_FAIL(ZB_E_ARG, "sort_count: distinct mode needs a payload");
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    if (g_sort_count_mode == 1) return sort_count_classic(c, k0, k1, v0, v1, n, key_bits, out_k, out_c);
    const bool weighted_sum = (v0 != nullptr) && !distinct;
    // the bucket kernel's bulk copies start at 16-byte boundaries inside the key arrays
    const bool aligned = (((uintptr_t)k0 | (uintptr_t)k1) & 15) == 0;
    if (g_sort_count_mode == 2 || weighted_sum || !aligned)
        return sort_count_segsort(c, k0, k1, v0, v1, n, key_bits, out_k, out_c, distinct);
    return sort_count_buckets(c, k0, k1, v0, v1, n, key_bits, out_k, out_c, distinct, pre);
}

}  // namespace zb
