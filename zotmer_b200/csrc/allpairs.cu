// allpairs.cu -- |X n Y| for ALL pairs of a collection of sorted k-mer sets: the work behind
// `zot dist` (zotmer/commands/dist.py:145-168 -> library/dist.py:241-265 split()) and `zot jaccard -a`
// (commands/jaccard.py:148-166 -> :31-54), where the reference walks two sorted arrays once per pair (and once
// more per measure).
//
// A pair-at-a-time merge reads 8 (|X| + |Y|) bytes per pair: O(N^2) trips over the sets.  Here every set is read
// once per BLOCK of 32 partner sets:
//
//   * the sets are cut into blocks of 32; a tile is a pair of blocks (bi <= bj): 32 sets on the diagonal (the 496
//     pairs inside the block), 64 sets off it (the 1024 pairs across the two blocks);
//   * the key space is cut into 2^cb buckets by the top bits of the key (bucket b of every set holds the same key
//     range, so intersections never cross buckets), cb such that a tile's sets together hold <= 2048 keys per
//     bucket; ap_offsets_kernel finds every bucket in every set with one streaming pass;
//   * ap_bucket_kernel: a CTA gathers the bucket's slice of each of the tile's sets into shared memory and
//     inserts every key into a shared-memory hash table -- the first entry of a key is its head and collects a
//     32-bit mask per block of "which sets hold this key".  A warp then takes 32 heads, transposes the 32 x 32 bit
//     matrix (5 shuffle steps), so that lane l holds "which of these 32 keys are in set l", and a pair's shared keys
//     are popc(column_i & column_j): 16 (diagonal) or 32 (cross) popcounts per lane for 32 keys, accumulated per
//     CTA in shared memory and flushed once per tile.
//   * work units: a tile's buckets are cut into AB_KS = 8 key-range shards; units [begin, end) are what one GPU
//     computes, and (a, b, c) of the shards of a pair simply add up -- 32 sets (one tile) still spread over 8 GPUs.
//
// HBM sees every set once per tile it belongs to (N / 32 times), the rest is shared-memory work.
// Skewed key spaces (a bucket that does not fit after three attempts): pair-at-a-time merge path (setops.cu
// pairs_abc), whole pairs attributed to the first shard of their tile.
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace zb {

static constexpr int AB_S = 32;               // sets per block
static constexpr int AB_KS = 8;               // key-range shards per tile
static constexpr int AB_THREADS = 256;
static constexpr int AB_WARPS = AB_THREADS / 32;
static constexpr int AB_PER = 8;
static constexpr int AB_CAP = AB_THREADS * AB_PER;   // keys of one bucket, all sets of the tile together
static constexpr int AB_HASH = 2 * AB_CAP;
static constexpr int AB_HASH_BITS = 12;
#define AB_EMPTY 0xffffffffu

// off[b * nsets + i] = first index of set i whose key >> shift is >= b   (b = 0 .. nb); `off` zeroed by the caller.
// One streaming pass over the keys (blockIdx.y = set): element j closes every bucket between its predecessor's and
// its own (a binary search per (bucket, set) misses the TLB on nearly every probe: 3.8 x slower, nwaymerge.cu).
__global__ void __launch_bounds__(256)
ap_offsets_kernel(const SetRef* __restrict__ sets, int nsets, int shift, uint32_t nb, uint32_t* __restrict__ off) {
    bucket_offsets_body(sets, nsets, shift, nb, off);
}

// largest number of keys one block of 32 sets holds in one bucket; one warp per (bucket, block)
__global__ void __launch_bounds__(256)
ap_blockmax_kernel(const uint32_t* __restrict__ off, int nsets, uint32_t nb, uint32_t nblk, unsigned int* __restrict__ mx) {
    const uint64_t wid = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= (uint64_t)nb * nblk) return;
    const uint32_t b = (uint32_t)(wid / nblk), blk = (uint32_t)(wid % nblk);
    const int i = (int)(blk * AB_S + lane_id());
    uint32_t v = 0;
    if (i < nsets) v = off[(size_t)(b + 1) * nsets + i] - off[(size_t)b * nsets + i];
    v = warp_sum(v);
    if (lane_id() == 0 && v) atomicMax(mx, v);
}

// Tiles.  More than two blocks: upper-triangular (diagonal included) block pair number t -> (bi, bj), bi <= bj < nblk;
// a diagonal tile counts the pairs inside its block (AP_WITHIN_A), the others the pairs across (AP_CROSS).  Up to two
// blocks (<= 64 sets): ONE tile does everything in one pass over the sets -- inside the first block, inside the second
// and across (as three tiles the sets would be read twice).
#define AP_WITHIN_A 1u
#define AP_WITHIN_B 2u
#define AP_CROSS 4u
__host__ __device__ inline uint64_t tile_count(uint64_t nblk) { return nblk <= 2 ? 1 : nblk * (nblk + 1) / 2; }
__host__ __device__ inline void tile_to_blocks(uint64_t t, uint64_t nblk, uint32_t& bi, uint32_t& bj, uint32_t& flags) {
    if (nblk <= 2) {
        bi = 0;
        bj = (uint32_t)(nblk - 1);
        flags = (nblk == 2) ? (AP_WITHIN_A | AP_WITHIN_B | AP_CROSS) : AP_WITHIN_A;
        return;
    }
    uint64_t r = 0, start = 0;
    while (start + (nblk - r) <= t) { start += nblk - r; r++; }
    bi = (uint32_t)r;
    bj = (uint32_t)(r + (t - start));
    flags = (bi == bj) ? AP_WITHIN_A : AP_CROSS;
}

// a run of buckets of one tile, and where it starts in the flattened list of all (tile, bucket) work items
struct ApSeg {
    uint32_t bi, bj;
    uint32_t b0, b1;
    uint64_t wstart;
    uint32_t flags, pad;
};

// 32 x 32 bit-matrix transpose across a warp: lane l gives row l, gets column l (bit r = bit l of lane r's row)
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x) {
    const unsigned l = lane_id();
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int j = 16 >> s;
        const uint32_t m = (s == 0) ? 0x0000ffffu : (s == 1) ? 0x00ff00ffu : (s == 2) ? 0x0f0f0f0fu : (s == 3) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (l & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y & m) << j));
    }
    return x;
}

// Persistent: CTA c takes the work items [c W / G, (c + 1) W / G) -- consecutive buckets of (mostly) one tile, so its
// reads of every set are sequential and its pair counters are flushed once.
// Shared memory: keys 16 KB + hash table 16 KB + 2 x masks 8 KB + set index 2 KB + counters 4 KB + slice tables = 55 KB
// -> 4 CTAs of 256 threads per SM (2 CTAs of 512 threads with 4096-key buckets: 4.1 instead of 3.2 ms for 32 sets --
// barrier and load-latency stalls, 29 % + 28 %, overlap better across four independent CTAs).
template <bool FOLD>   // FOLD: the one tile of <= 64 sets (inside both blocks + across); else diagonal and cross tiles
__global__ void __launch_bounds__(AB_THREADS, FOLD ? 3 : 4)
ap_bucket_kernel(const SetRef* __restrict__ sets, int nsets, const uint32_t* __restrict__ off, const ApSeg* __restrict__ segs,
                 uint32_t nsegs, uint64_t W, unsigned long long* __restrict__ isect) {
    constexpr int two_acc = FOLD ? 1 : 0;
    extern __shared__ __align__(16) unsigned char ab_raw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(ab_raw);                  // [AB_CAP] gathered keys
    uint32_t* table = reinterpret_cast<uint32_t*>(sk + AB_CAP);          // [AB_HASH] position of a key's head
    uint32_t* mA = table + AB_HASH;                                      // [AB_CAP] at a head: sets of block bi holding the key
    uint32_t* mB = mA + AB_CAP;                                          // [AB_CAP]            sets of block bj (off the diagonal)
    // shared keys per pair of the tile.  acc_w[i * 32 + j]: i < j = pair (i, j) inside block bi, i > j = pair (j, i)
    // inside block bj; acc_x[i * 32 + j] = set i of block bi with set j of block bj (the same array when a tile only
    // counts across)
    uint32_t* acc_w = mB + AB_CAP;                                       // [32 * 32]
    uint32_t* acc_x = two_acc ? acc_w + AB_S * AB_S : acc_w;             // [32 * 32]
    uint8_t* sb = reinterpret_cast<uint8_t*>(acc_w + (two_acc ? 2 : 1) * AB_S * AB_S);   // [AB_CAP] which of the tile's sets an entry came from
    __shared__ uint32_t spre[2 * AB_S + 1];                              // slice starts inside the bucket
    __shared__ uint32_t soff[2 * AB_S];                                  // slice starts inside the sets
    __shared__ uint32_t s_useful;                                        // useful heads of the bucket (count phase)

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;
    if (w0 >= w1) return;
    uint32_t s = 0;   // segment of w0
    {
        uint32_t lo = 0, hi = nsegs - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (segs[mid].wstart <= w0) lo = mid; else hi = mid - 1;
        }
        s = lo;
    }
    for (int idx = tid; idx < (two_acc ? 2 : 1) * AB_S * AB_S; idx += AB_THREADS) acc_w[idx] = 0;
    ApSeg seg = segs[s];
    uint64_t seg_end = seg.wstart + (seg.b1 - seg.b0);

    for (uint64_t wi = w0; wi <= w1; wi++) {
        // ---- tile change (or the end of my range): flush the pair counters
        if (wi == w1 || wi >= seg_end) {
            __syncthreads();
            for (int idx = tid; idx < AB_S * AB_S; idx += AB_THREADS) {
                const uint32_t i = idx >> 5, j = idx & 31;
                const uint32_t fl = FOLD ? (AP_WITHIN_A | AP_WITHIN_B | AP_CROSS) : (seg.bi == seg.bj ? AP_WITHIN_A : AP_CROSS);
                if (fl & (AP_WITHIN_A | AP_WITHIN_B)) {
                    const uint32_t v = acc_w[idx];
                    acc_w[idx] = 0;
                    // upper triangle: inside block bi; lower triangle: inside block bj
                    const uint64_t gi = (i < j) ? (uint64_t)seg.bi * AB_S + i : (uint64_t)seg.bj * AB_S + j;
                    const uint64_t gj = (i < j) ? (uint64_t)seg.bi * AB_S + j : (uint64_t)seg.bj * AB_S + i;
                    if (v && i != j && gj < (uint64_t)nsets)
                        atomicAdd(&isect[gi * (2ull * nsets - gi - 1) / 2 + (gj - gi - 1)], (unsigned long long)v);
                }
                if (fl & AP_CROSS) {
                    const uint32_t v = acc_x[idx];
                    acc_x[idx] = 0;
                    const uint64_t gi = (uint64_t)seg.bi * AB_S + i, gj = (uint64_t)seg.bj * AB_S + j;
                    if (v && gj < (uint64_t)nsets)
                        atomicAdd(&isect[gi * (2ull * nsets - gi - 1) / 2 + (gj - gi - 1)], (unsigned long long)v);
                }
            }
            if (wi == w1) break;
            while (wi >= seg_end) {
                s++;
                seg = segs[s];
                seg_end = seg.wstart + (seg.b1 - seg.b0);
            }
        }
        const uint32_t b = seg.b0 + (uint32_t)(wi - seg.wstart);
        const uint32_t flags = FOLD ? (AP_WITHIN_A | AP_WITHIN_B | AP_CROSS) : (seg.bi == seg.bj ? AP_WITHIN_A : AP_CROSS);
        const bool two = (flags & (AP_WITHIN_B | AP_CROSS)) != 0;                 // the tile has a second block
        const int nA = min(AB_S, nsets - (int)seg.bi * AB_S);
        const int nT = nA + (two ? min(AB_S, nsets - (int)seg.bj * AB_S) : 0);    // sets of the tile
        __syncthreads();   // the previous bucket (and the flush) are done with the shared arrays

        // ---- where every set's slice of this bucket starts
        if (warp == 0) {
            uint32_t len[2], tot = 0;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int t = (int)lane * 2 + u;
                len[u] = 0;
                if (t < nT) {
                    const int g = (t < nA) ? (int)seg.bi * AB_S + t : (int)seg.bj * AB_S + (t - nA);
                    const uint32_t o0 = __ldg(off + (size_t)b * nsets + g);
                    len[u] = __ldg(off + (size_t)(b + 1) * nsets + g) - o0;
                    soff[t] = o0;
                }
                tot += len[u];
            }
            uint32_t ex = warp_incl_scan(tot) - tot;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int t = (int)lane * 2 + u;
                if (t < nT) spre[t] = ex;
                ex += len[u];
            }
            if (lane == 31) spre[nT] = ex;
        }
        {
            const uint4 e4 = make_uint4(AB_EMPTY, AB_EMPTY, AB_EMPTY, AB_EMPTY);
#pragma unroll
            for (int j = 0; j < AB_HASH / 4 / AB_THREADS; j++) reinterpret_cast<uint4*>(table)[j * AB_THREADS + tid] = e4;
        }
        if (tid == 0) s_useful = 0;
        __syncthreads();
        const int m = (int)spre[nT];
        if (m == 0) continue;
        for (int q = (int)tid; q < m; q += AB_THREADS) { mA[q] = 0; mB[q] = 0; }

        // ---- gather: a warp copies whole slices, four at a time (eight loads in flight per lane)
        {
            constexpr int U = 4;
            for (int t0 = (int)warp; t0 < nT; t0 += U * AB_WARPS) {
                uint32_t q0[U], len[U];
                const uint64_t* pk[U];
                uint32_t mx = 0;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int t = t0 + u * AB_WARPS;
                    q0[u] = 0; len[u] = 0; pk[u] = nullptr;
                    if (t < nT) {
                        const int g = (t < nA) ? (int)seg.bi * AB_S + t : (int)seg.bj * AB_S + (t - nA);
                        q0[u] = spre[t];
                        len[u] = spre[t + 1] - q0[u];
                        pk[u] = sets[g].k + soff[t];
                    }
                    mx = max(mx, len[u]);
                }
                for (uint32_t e = lane; e < mx; e += 32) {
                    uint64_t kv[U];
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (e < len[u]) kv[u] = __ldg(pk[u] + e);
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (e < len[u]) { sk[q0[u] + e] = kv[u]; sb[q0[u] + e] = (uint8_t)(t0 + u * AB_WARPS); }
                }
            }
        }
        __syncthreads();

        // ---- dedupe: the first entry to claim a key's slot is its head; every entry sets its set's bit there
#pragma unroll
        for (int j = 0; j < AB_PER; j++) {
            const int q = j * AB_THREADS + (int)tid;
            if (q < m) {
                const uint64_t x = sk[q];
                uint32_t h = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> (64 - AB_HASH_BITS));
                uint32_t hq;
                while (true) {
                    const uint32_t old = atomicCAS(&table[h], AB_EMPTY, (uint32_t)q);
                    if (old == AB_EMPTY) { hq = (uint32_t)q; break; }
                    if (sk[old] == x) { hq = old; break; }
                    h = (h + 1) & (AB_HASH - 1);
                }
                const int t = (int)sb[q];
                if (t < nA) atomicOr(&mA[hq], 1u << t); else atomicOr(&mB[hq], 1u << (t - nA));
            }
        }
        __syncthreads();

        // ---- the heads that can contribute to a pair (a key in two sets of a block, or in both blocks) move to the front
        // of a dense list, in any order -- a pair's count is a sum over keys.  Three entries in four are not heads (a
        // key of related genomes sits in several of the tile's sets) and a head's position is that of its first entry,
        // so without this nearly every group of 32 positions held a few heads and was transposed for them: the count
        // phase was three quarters of the kernel's instructions (profiles/r02_allpairs.md).
        uint2* cm = reinterpret_cast<uint2*>(sk);                        // [AB_CAP] (mA, mB) of the useful heads; the keys are done with
#pragma unroll
        for (int j = 0; j < AB_PER; j++) {
            const int q = j * AB_THREADS + (int)tid;
            const uint32_t a = (q < m) ? mA[q] : 0u;
            const uint32_t bb = (q < m && two) ? mB[q] : 0u;
            const bool useful = ((flags & AP_WITHIN_A) && (a & (a - 1)) != 0) || ((flags & AP_WITHIN_B) && (bb & (bb - 1)) != 0) ||
                                ((flags & AP_CROSS) && a != 0 && bb != 0);
            const unsigned bal = __ballot_sync(0xffffffffu, useful);
            if (bal) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_useful, (uint32_t)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (useful) cm[base + __popc(bal & lanemask_lt())] = make_uint2(a, bb);
            }
        }
        __syncthreads();
        const int H = (int)s_useful;
        // ---- count: 32 heads per warp step; lane l ends up with "which of these 32 keys are in set l"
        for (int c0 = (int)warp * 32; c0 < H; c0 += AB_THREADS) {
            const int q = c0 + (int)lane;
            const uint2 ab = (q < H) ? cm[q] : make_uint2(0u, 0u);
            const uint32_t a = ab.x;
            const uint32_t bb = ab.y;
            const uint32_t colA = warp_transpose32(a);
            const uint32_t colB = two ? warp_transpose32(bb) : 0u;
            if (flags & (AP_WITHIN_A | AP_WITHIN_B)) {
                // pairs (l, l + d mod 32), d = 1 .. 16: every unordered pair once
#pragma unroll
                for (int d = 1; d <= 16; d++) {
                    const uint32_t o = (lane + d) & 31;
                    const uint32_t oa = __shfl_sync(0xffffffffu, colA, o);
                    const uint32_t ob = __shfl_sync(0xffffffffu, colB, o);
                    if (d < 16 || lane < 16) {
                        const uint32_t lo = min(lane, o), hi = max(lane, o);
                        const uint32_t ca = __popc(colA & oa), cbb = __popc(colB & ob);
                        if (ca && (flags & AP_WITHIN_A)) atomicAdd(&acc_w[lo * 32 + hi], ca);
                        if (cbb && (flags & AP_WITHIN_B)) atomicAdd(&acc_w[hi * 32 + lo], cbb);
                    }
                }
            }
            if (flags & AP_CROSS) {
#pragma unroll
                for (int d = 0; d < 32; d++) {
                    const uint32_t o = (lane + d) & 31;
                    const uint32_t other = __shfl_sync(0xffffffffu, colB, o);
                    const uint32_t cnt = __popc(colA & other);
                    if (cnt) atomicAdd(&acc_x[lane * 32 + o], cnt);
                }
            }
        }
    }
}

uint64_t allpairs_tiles(int nsets) {
    return tile_count(div_up((size_t)nsets, AB_S)) * AB_KS;
}

// the set pairs (i < j) a tile covers
static void tile_pairs(uint64_t t, uint32_t nblk, int nsets, std::vector<uint32_t>& I, std::vector<uint32_t>& J) {
    uint32_t bi, bj, flags;
    tile_to_blocks(t, nblk, bi, bj, flags);
    const uint32_t a0 = bi * AB_S, a1 = std::min<uint32_t>((bi + 1) * AB_S, nsets);
    const uint32_t b0 = bj * AB_S, b1 = std::min<uint32_t>((bj + 1) * AB_S, nsets);
    if (flags & AP_WITHIN_A)
        for (uint32_t i = a0; i < a1; i++)
            for (uint32_t j = i + 1; j < a1; j++) { I.push_back(i); J.push_back(j); }
    if (flags & AP_CROSS)
        for (uint32_t i = a0; i < a1; i++)
            for (uint32_t j = b0; j < b1; j++) { I.push_back(i); J.push_back(j); }
    if (flags & AP_WITHIN_B)
        for (uint32_t i = b0; i < b1; i++)
            for (uint32_t j = i + 1; j < b1; j++) { I.push_back(i); J.push_back(j); }
}

// first bucket of key-range shard s (s = 0 .. AB_KS) when the key space is cut into nb buckets
static inline uint64_t shard_bucket(uint64_t s, uint64_t nb) { return (s * nb + AB_KS - 1) / AB_KS; }

// a per pair on the device, b and c on the host from the shard sizes.  Returns false when the key space is too skewed.
// One piece of work: shards [s0, s1) of tile t.
struct ApPiece { uint64_t t; uint32_t s0, s1; };

// units ub, ub + stride, ... < ue as pieces; consecutive shards of one tile are one piece
static std::vector<ApPiece> unit_pieces(uint64_t ub, uint64_t ue, uint64_t stride) {
    std::vector<ApPiece> out;
    for (uint64_t u = ub; u < ue; u += stride) {
        const uint64_t t = u / AB_KS;
        const uint32_t s = (uint32_t)(u % AB_KS);
        if (!out.empty() && out.back().t == t && out.back().s1 == s) out.back().s1 = s + 1;
        else out.push_back(ApPiece{t, s, s + 1});
    }
    return out;
}

static bool allpairs_buckets(Ctx* c, const SetRef* d_sets, const std::vector<SetRef>& refs, int key_bits,
                             const std::vector<ApPiece>& pieces, uint64_t* abc_host) {
    const int nsets = (int)refs.size();
    const uint64_t npairs = (uint64_t)nsets * (nsets - 1) / 2;
    const uint32_t nblk = (uint32_t)div_up((size_t)nsets, AB_S);
    // the fullest tile: the two largest blocks (one block when there is only one)
    uint64_t top1 = 0, top2 = 0;
    size_t nmax = 0;
    for (uint32_t blk = 0; blk < nblk; blk++) {
        uint64_t t = 0;
        for (int i = blk * AB_S; i < std::min<int>((blk + 1) * AB_S, nsets); i++) {
            if (refs[i].n >= ((uint64_t)1 << 32)) return false;
            t += refs[i].n;
            nmax = std::max<size_t>(nmax, refs[i].n);
        }
        if (t > top1) { top2 = top1; top1 = t; } else if (t > top2) top2 = t;
    }
    const uint64_t tile_total = top1 + top2;
    if (tile_total == 0) return true;
    int cb = 0;
    while (cb < key_bits && (tile_total >> cb) > (uint64_t)AB_CAP * 2 / 3) cb++;
    cb = std::max(cb, std::min(3, key_bits));
    DBuf<uint32_t> off;
    uint32_t nb = 0;
    for (int attempt = 0;; attempt++) {
        if (cb > 30 || (((size_t)1 << cb) + 1) * (size_t)nsets > ((size_t)1 << 29)) return false;
        nb = 1u << cb;
        const size_t noff = (size_t)(nb + 1) * nsets;
        off.alloc(c, noff + 2);
        unsigned int* d_mx = off.get() + noff;
        ZB_CUDA(dev_memset(c, off.get(), 0, (noff + 2) * 4));
        {
            Stage st(c, "allpairs_offsets");
            const dim3 grid((unsigned)std::min<size_t>(std::max<size_t>(div_up(nmax, 256 * 2 * 8), 1), 65535), (unsigned)nsets);
            ap_offsets_kernel<<<grid, 256, 0, c->stream>>>(d_sets, nsets, key_bits - cb, nb, off.get());
            ZB_LAUNCH_CHECK(c);
            ap_blockmax_kernel<<<(unsigned)div_up((size_t)nb * nblk, 8), 256, 0, c->stream>>>(off.get(), nsets, nb, nblk, d_mx);
            ZB_LAUNCH_CHECK(c);
        }
        ZB_CUDA(read_back(c, d_mx, 4));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        const uint64_t mx = reinterpret_cast<uint32_t*>(c->h_scalars)[0];
        const uint64_t worst = (nblk > 1) ? 2 * mx : mx;   // any two blocks together
        if (worst <= (uint64_t)AB_CAP) break;
        if (attempt >= 3 || cb >= key_bits) return false;
        int more = 1;
        while (((worst >> more) > (uint64_t)AB_CAP * 2 / 3) && more < 8) more++;
        cb = std::min(key_bits, cb + more);
    }

    // work list: for every tile touched by the units [ub, ue), the buckets of its shards in that range
    std::vector<ApSeg> segs;
    uint64_t W = 0;
    for (const ApPiece& pc : pieces) {
        const uint64_t t = pc.t, s0 = pc.s0, s1 = pc.s1;
        ApSeg sg;
        sg.pad = 0;
        tile_to_blocks(t, nblk, sg.bi, sg.bj, sg.flags);
        sg.b0 = (uint32_t)std::min<uint64_t>(shard_bucket(s0, nb), nb);
        sg.b1 = (uint32_t)std::min<uint64_t>(shard_bucket(s1, nb), nb);
        sg.wstart = W;
        if (sg.b1 > sg.b0) {
            W += sg.b1 - sg.b0;
            segs.push_back(sg);
        }
    }
    std::vector<uint64_t> isect(npairs, 0);
    if (W) {
        DBuf<ApSeg> d_segs(c, segs.size());
        DBuf<uint64_t> d_isect(c, npairs);
        ZB_CUDA(cudaMemcpyAsync(d_segs.get(), segs.data(), segs.size() * sizeof(ApSeg), cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(dev_memset(c, d_isect.get(), 0, npairs * 8));
        const int two_acc = (nblk == 2) ? 1 : 0;   // the one tile of <= 64 sets counts inside both blocks and across
        const size_t smem = (size_t)AB_CAP * 8 + (size_t)AB_HASH * 4 + (size_t)2 * AB_CAP * 4 + (size_t)(1 + two_acc) * AB_S * AB_S * 4 + AB_CAP;
        const unsigned grid = (unsigned)std::min<uint64_t>(W, (uint64_t)c->sm_count * (two_acc ? 3 : 4));
        {
            Stage st(c, "allpairs");
            unsigned long long* d_is = reinterpret_cast<unsigned long long*>(d_isect.get());
            if (two_acc) {
                ZB_CUDA(cudaFuncSetAttribute(ap_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ap_bucket_kernel<true><<<grid, AB_THREADS, smem, c->stream>>>(d_sets, nsets, off.get(), d_segs.get(), (uint32_t)segs.size(), W, d_is);
            } else {
                ZB_CUDA(cudaFuncSetAttribute(ap_bucket_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ap_bucket_kernel<false><<<grid, AB_THREADS, smem, c->stream>>>(d_sets, nsets, off.get(), d_segs.get(), (uint32_t)segs.size(), W, d_is);
            }
            ZB_LAUNCH_CHECK(c);
        }
        ZB_CUDA(cudaMemcpyAsync(isect.data(), d_isect.get(), npairs * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    // sizes of the shards: rows of `off` at the shard boundaries
    std::vector<uint32_t> rows((size_t)(AB_KS + 1) * nsets);
    for (int s = 0; s <= AB_KS; s++)
        ZB_CUDA(cudaMemcpyAsync(rows.data() + (size_t)s * nsets, off.get() + (size_t)std::min<uint64_t>(shard_bucket(s, nb), nb) * nsets,
                                (size_t)nsets * 4, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    // |X_i| and |X_j| restricted to the shards of a tile's pieces add up piece by piece; a (from the device) already
    // holds the sum over all pieces of the tile
    for (const ApPiece& pc : pieces) {
        const uint64_t t = pc.t, s0 = pc.s0, s1 = pc.s1;
        std::vector<uint32_t> I, J;
        tile_pairs(t, nblk, nsets, I, J);
        for (size_t q = 0; q < I.size(); q++) {
            const uint64_t i = I[q], j = J[q];
            const uint64_t p = i * (2ull * nsets - i - 1) / 2 + (j - i - 1);
            const uint64_t ni = rows[s1 * nsets + i] - rows[s0 * nsets + i], nj = rows[s1 * nsets + j] - rows[s0 * nsets + j];
            abc_host[3 * p + 1] += ni;
            abc_host[3 * p + 2] += nj;
        }
    }
    uint64_t last_t = ~0ull;
    for (const ApPiece& pc : pieces) {
        if (pc.t == last_t) continue;     // pieces of one tile are adjacent in the list
        last_t = pc.t;
        std::vector<uint32_t> I, J;
        tile_pairs(pc.t, nblk, nsets, I, J);
        for (size_t q = 0; q < I.size(); q++) {
            const uint64_t i = I[q], j = J[q];
            const uint64_t p = i * (2ull * nsets - i - 1) / 2 + (j - i - 1);
            const uint64_t a = isect[p];
            abc_host[3 * p] = a;
            abc_host[3 * p + 1] -= a;
            abc_host[3 * p + 2] -= a;
        }
    }
    return true;
}

void allpairs_abc(Ctx* c, const std::vector<SetRef>& refs, uint64_t unit_begin, uint64_t unit_end, uint64_t unit_stride,
                  uint64_t* abc_host) {
    const int nsets = (int)refs.size();
    const uint64_t npairs = (uint64_t)nsets * (nsets - 1) / 2;
    if (npairs == 0) return;
    const uint32_t nblk = (uint32_t)div_up((size_t)nsets, AB_S);
    const uint64_t all_units = allpairs_tiles(nsets);
    if (unit_end == 0 || unit_end > all_units) unit_end = all_units;
    memset(abc_host, 0, npairs * 3 * 8);
    if (unit_begin >= unit_end) return;
    if (unit_stride < 1) unit_stride = 1;
    const std::vector<ApPiece> pieces = unit_pieces(unit_begin, unit_end, unit_stride);

    DBuf<SetRef> d_refs(c, nsets);
    ZB_CUDA(cudaMemcpyAsync(d_refs.get(), refs.data(), nsets * sizeof(SetRef), cudaMemcpyHostToDevice, c->stream));
    // key range: the sets are sorted, so the largest key is the largest last element
    uint64_t maxkey = 0;
    {
        std::vector<uint64_t> last(nsets, 0);
        for (int i = 0; i < nsets; i++)
            if (refs[i].n) ZB_CUDA(cudaMemcpyAsync(&last[i], refs[i].k + refs[i].n - 1, 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < nsets; i++) maxkey = std::max(maxkey, last[i]);
    }
    const int key_bits = maxkey ? 64 - __builtin_clzll(maxkey) : 1;
    if (!getenv("ZB_ALLPAIRS_PAIRWISE") && allpairs_buckets(c, d_refs.get(), refs, key_bits, pieces, abc_host)) return;

    // skewed key space: pair-at-a-time merge path (setops.cu); a pair goes, whole, to ONE unit of its tile: shard
    // t mod AB_KS, so that strided shares (every rank holds some shards of every tile) still split the tiles evenly
    memset(abc_host, 0, npairs * 3 * 8);
    std::vector<uint32_t> I, J;
    for (const ApPiece& pc : pieces) {
        const uint32_t home = (uint32_t)(pc.t % AB_KS);
        if (pc.s0 <= home && home < pc.s1) tile_pairs(pc.t, nblk, nsets, I, J);
    }
    if (I.empty()) return;
    std::vector<uint64_t> tmp(I.size() * 3);
    pairs_abc_host(c, refs, I.data(), J.data(), I.size(), tmp.data());
    for (size_t q = 0; q < I.size(); q++) {
        const uint64_t i = I[q], j = J[q];
        const uint64_t p = i * (2ull * nsets - i - 1) / 2 + (j - i - 1);
        const uint64_t a = tmp[3 * q];
        abc_host[3 * p] = a;
        abc_host[3 * p + 1] = refs[i].n - a;
        abc_host[3 * p + 2] = refs[j].n - a;
    }
}

}  // namespace zb
