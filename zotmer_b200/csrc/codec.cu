// codec.cu -- codec64 (+ delta) stream codec on the device.  Replaces zotmer/library/codec64.py:82-150
// (encode / decode) and the delta / undelta wrappers of zotmer/library/files.py:85-110, i.e. what
// writeKmersAndCounts2 (files.py:209-217) and readKmersAndCounts (files.py:219-227) spend their time in.
//
// Word format (codec64.py:13-46): low 4 bits = number of values g (1..6), then g fields of 60/g bits, first
// value in the lowest field.  The encoder is greedy (codec64.py:82-120): a word takes the longest prefix of the
// pending values (at most 6) whose widest member fits 60/g bits.
//
// Encode.  Whether position i starts a word depends on every earlier decision, but only through "how many
// positions until the next word starts" (0..5): a 6-state machine.  J[i] = the group length the greedy rule
// gives a word that starts at i depends on v[i..i+5] only (the fit test is monotone in g).
//   enc_tile_kernel   per 2048-value tile: J[] in shared memory; every thread walks its eight positions for each of
//                     the six entry states (registers only), a block scan composes the threads' maps -> (exit
//                     state, number of words) of the tile for each entry state
//   enc_scan_kernel   composes the per-tile maps in order -> entry state and first word index of every tile
//   enc_emit_kernel   the same thread maps and scan, now applied to the tile's real entry state: every thread knows
//                     the state in which the walk reaches it, walks its eight positions and packs the words that
//                     start there
// (The first version walked a tile with one thread per entry state: 2048 dependent shared-memory loads each, 3.2 ms
// for the four streams of one bench step -- the largest stage of the step once both bench arms did the same work.)
// Decode.  dec_tile_kernel: values and value sum per 512-word tile; dec_scan_kernel: exclusive scan of both;
// dec_emit_kernel: unpack, add the running sum (undelta), stage in shared memory, write coalesced.
//
// Algorithmic bytes: encode 2 x 8 B/value read + 8 B/word written; decode 2 x 8 B/word read + 8 (4) B/value
// written.  Nothing here is on the host: the set-level entry points move only the packed words over PCIe.
#include <vector>

#include "kernels.h"

namespace zb {

// 60 / g for g = 1..6, as a byte table in a register (branch-free: a ternary chain next to max(g, 1) was
// miscompiled by ptxas 12.9 into a packed VIMNMX.U16x2 whose predicate doubled as the "g == 1" test)
__device__ __forceinline__ int cw_width(int g) {
    return (int)((0x0A0C0F141E3Cull >> (8 * (g - 1))) & 0xffull);
}
__device__ __forceinline__ int bitlen64(uint64_t x) { return 64 - __clzll((long long)x); }

static constexpr int EN_THREADS = 256;
static constexpr int EN_PER = 8;
static constexpr int EN_TILE = EN_THREADS * EN_PER;   // 2048 values

// bit lengths of the tile's values (+5 beyond its end) -> sbl[]; group length for a word starting at p -> sJ[p]
// A stream may be ONE RANGE of a longer one (several GPUs encode consecutive ranges of a sorted set into one file,
// api: zb_set_encode_plan): `halo` then holds the first n_halo (<= 5) values of the following range -- a word that
// starts in this range may run into them -- and `first_base` is the value in front of vals[0] (delta only).
struct EncRange {
    uint64_t first_base;        // subtracted from vals[0] when delta (0 for a whole stream)
    uint64_t halo[5];           // values behind vals[n - 1]
    uint32_t n_halo;
};

// class of a bit length: the largest group a value of that width may belong to (60 / g bits per value), 0 = none.
// The greedy rule then reads: J[p] = the largest g with min(class[p .. p + g - 1]) >= g (feasibility is monotone in g).
__device__ __forceinline__ uint32_t enc_class_of(int bl) {
    return (bl <= 10) ? 6u : (bl <= 12) ? 5u : (bl <= 15) ? 4u : (bl <= 20) ? 3u : (bl <= 30) ? 2u : (bl <= 60) ? 1u : 0u;
}
// the values are staged with one slot of padding per eight, so that a thread's eight consecutive values start in a
// different bank pair from its neighbours' (without it the packing loop's loads are a 16-way bank conflict)
#define EN_SV(p) ((p) + ((p) >> 3))
static constexpr int EN_SV_SIZE = EN_TILE + 8 + (EN_TILE + 8) / 8 + 1;

template <typename T>
__device__ __forceinline__ int enc_bitlen(T v) {
    return sizeof(T) == 4 ? 32 - __clz((int)(uint32_t)v) : 64 - __clzll((long long)(uint64_t)v);
}

// classes of the tile's values (+ 8 beyond its end: every thread reads a 16-byte window) -> scls[], values -> sv[].
// s_lut: the class of every bit length, 65 bytes in 17 different banks -- any mix of lengths in a warp is conflict-free
// (constant memory would serialise on the lanes' different lengths; compares cost a dozen instructions per value).
// A tile that lies inside the data takes the short path: no end-of-data, halo or first-value cases.
template <typename T>
__device__ __forceinline__ void enc_prepare(const T* __restrict__ vals, uint64_t n, bool delta, uint64_t base, const EncRange& rg,
                                            uint8_t* scls, uint8_t* s_lut, uint64_t* sv, unsigned int* err) {
    const unsigned tid = threadIdx.x;
    if (tid < 65) s_lut[tid] = (uint8_t)enc_class_of((int)tid);
    __syncthreads();
    if (base > 0 && base + EN_TILE + 8 <= n) {
        const T* __restrict__ src = vals + base;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < (EN_TILE + 8 + EN_THREADS - 1) / EN_THREADS; j++) {
            const int p = j * EN_THREADS + (int)tid;
            if (p < EN_TILE + 8) {
                const T x = src[p];
                const T v = delta ? (T)(x - src[p - 1]) : x;
                const uint32_t cls = s_lut[enc_bitlen<T>(v)];
                bad |= cls == 0;
                scls[p] = (uint8_t)cls;
                if (sv) sv[EN_SV(p)] = (uint64_t)v;
            }
        }
        if (bad) atomicExch(err, 1u);   // a value (or gap) wider than 60 bits (every position here is real data)
    } else {
        for (int p = tid; p < EN_TILE + 8; p += EN_THREADS) {
            const uint64_t i = base + p;
            uint64_t v = 0;
            uint32_t cls = 0;   // past the end of the data: fits nowhere, so no group runs over the end
            if (i < n + rg.n_halo) {
                const uint64_t x = (i < n) ? (uint64_t)vals[i] : rg.halo[i - n];
                const uint64_t prev = (i == 0) ? rg.first_base : (i <= n) ? (uint64_t)vals[i - 1] : rg.halo[i - n - 1];
                v = delta ? x - prev : x;
                cls = s_lut[64 - __clzll((long long)v)];
                if (cls == 0 && i < n) atomicExch(err, 1u);   // only real data can be in error
            }
            scls[p] = (uint8_t)cls;
            if (sv) sv[EN_SV(p)] = v;
        }
    }
    __syncthreads();
}

// group lengths J of the thread's eight positions (one byte each) from the classes of positions 8t .. 8t + 12: four
// positions at a time with the byte-wise SIMD instructions (running minimum of the classes, compared with q + 1)
__device__ __forceinline__ uint32_t enc_j4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t m = a;
    uint32_t j = __vcmpgeu4(m, 0x01010101u) & 0x01010101u;
    m = __vminu4(m, __byte_perm(a, b, 0x4321));
    j += __vcmpgeu4(m, 0x02020202u) & 0x01010101u;
    m = __vminu4(m, __byte_perm(a, b, 0x5432));
    j += __vcmpgeu4(m, 0x03030303u) & 0x01010101u;
    m = __vminu4(m, __byte_perm(a, b, 0x6543));
    j += __vcmpgeu4(m, 0x04040404u) & 0x01010101u;
    m = __vminu4(m, b);
    j += __vcmpgeu4(m, 0x05050505u) & 0x01010101u;
    m = __vminu4(m, __byte_perm(b, c, 0x4321));
    j += __vcmpgeu4(m, 0x06060606u) & 0x01010101u;
    return __vmaxu4(j, 0x01010101u);   // a value that fits nowhere (flagged) still moves every walk on
}
__device__ __forceinline__ uint64_t enc_j8(const uint8_t* scls) {
    const uint64_t lo = *reinterpret_cast<const uint64_t*>(scls + threadIdx.x * EN_PER);
    const uint64_t hi = *reinterpret_cast<const uint64_t*>(scls + threadIdx.x * EN_PER + 8);
    const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi, w3 = (uint32_t)(hi >> 32);
    return (uint64_t)enc_j4(w0, w1, w2) | ((uint64_t)enc_j4(w1, w2, w3) << 32);
}

// ---- the walk inside a tile, in parallel.  Thread t owns positions [8t, 8t + 8) of the tile (cut at the tile's last
// value `lim`).  Its map: "the next word starts r positions into my region" (r = 0..5) -> (how many positions past my
// region's end the next word after it starts, how many words start inside it).  Maps compose associatively --
// (a then b)[r] = b[a[r]] -- so an inclusive scan over the threads gives every thread the state in which the walk
// reaches it for each of the six states in which it may enter the tile.
// A map is six bytes (lo: entries 0..3, hi: entries 4, 5): composing two maps is then two byte permutes (PRMT), the
// selectors being the first map's bytes squeezed into nibbles.  (Six 3-bit fields and shift/mask lookups made the
// scans a quarter of all instructions of the encoder, profiles/r02_codec.md.)
struct Map6 {
    uint32_t lo, hi;
};
__device__ __forceinline__ uint32_t bytes_to_nibbles(uint32_t b) {   // four bytes < 16 -> four nibbles
    const uint32_t y = (b | (b >> 4)) & 0x00ff00ffu;
    return (y | (y >> 8)) & 0xffffu;
}
__device__ __forceinline__ uint32_t nibbles_to_bytes(uint32_t x) {   // four nibbles -> four bytes
    const uint32_t y = (x | (x << 8)) & 0x00ff00ffu;
    return (y | (y << 4)) & 0x0f0f0f0fu;
}
// t[sel[r]] for r = 0..5: `t` a six-byte table, `sel` a map whose entries select
__device__ __forceinline__ Map6 map_lookup(Map6 sel, Map6 t) {
    Map6 r;
    r.lo = __byte_perm(t.lo, t.hi, bytes_to_nibbles(sel.lo));
    r.hi = __byte_perm(t.lo, t.hi, (sel.hi | (sel.hi >> 4)) & 0xffu) & 0xffffu;
    return r;
}
__device__ __forceinline__ Map6 map_then(Map6 a, Map6 b) { return map_lookup(a, b); }
__device__ __forceinline__ Map6 map_id() { Map6 m; m.lo = 0x03020100u; m.hi = 0x0504u; return m; }
__device__ __forceinline__ uint32_t map_at(Map6 m, uint32_t r) { return (r < 4 ? (m.lo >> (8 * r)) : (m.hi >> (8 * (r - 4)))) & 0xffu; }
__device__ __forceinline__ Map6 map_shfl_up(Map6 m, int o) {
    Map6 r;
    r.lo = __shfl_up_sync(0xffffffffu, m.lo, o);
    r.hi = __shfl_up_sync(0xffffffffu, m.hi, o);
    return r;
}

struct ThreadWalk {
    Map6 map;      // exit state per entry state
    Map6 cnt;      // words started per entry state
};

// j8: the group lengths J of the thread's eight positions (one byte each), m = positions of the region that hold data.
// Back to front: where the walk that visits position p leaves the region and how many words it starts on the way, four
// bits per position.
__device__ __forceinline__ ThreadWalk thread_walk(uint64_t j8, int m) {
    uint32_t ex = 0, cn = 0;
#pragma unroll
    for (int p = 7; p >= 0; p--) {
        const int q = p + (int)((j8 >> (8 * p)) & 0xffu);
        uint32_t e, c;
        if (q >= m) { e = (uint32_t)(q - m); c = 1; }     // (positions p >= m hold values nobody looks up)
        else { e = (ex >> (4 * q)) & 7u; c = 1 + ((cn >> (4 * q)) & 15u); }
        ex |= e << (4 * p);
        cn |= c << (4 * p);
    }
    ThreadWalk w;
    if (m == EN_PER) {   // a full region: entry r is position r
        w.map.lo = nibbles_to_bytes(ex & 0xffffu);
        w.map.hi = nibbles_to_bytes((ex >> 16) & 0xffu);
        w.cnt.lo = nibbles_to_bytes(cn & 0xffffu);
        w.cnt.hi = nibbles_to_bytes((cn >> 16) & 0xffu);
    } else {
        w.map.lo = w.map.hi = w.cnt.lo = w.cnt.hi = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const uint32_t e = (r < m) ? ((ex >> (4 * r)) & 7u) : (uint32_t)(r - m);
            const uint32_t c = (r < m) ? ((cn >> (4 * r)) & 15u) : 0u;
            if (r < 4) { w.map.lo |= e << (8 * r); w.cnt.lo |= c << (8 * r); }
            else { w.map.hi |= e << (8 * (r - 4)); w.cnt.hi |= c << (8 * (r - 4)); }
        }
    }
    return w;
}

// exclusive prefix of the threads' maps over the block (identity for thread 0); *total = the whole tile's map
__device__ __forceinline__ Map6 block_map_excl_scan(Map6 mine, Map6* s_warp /*[EN_THREADS / 32 + 1]*/, Map6* total) {
    const unsigned l = lane_id(), wid = threadIdx.x >> 5;
    Map6 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Map6 prev = map_shfl_up(inc, o);
        if (l >= (unsigned)o) inc = map_then(prev, inc);
    }
    if (l == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {   // the maps of the eight warps, scanned by the first lanes of one warp
        constexpr int NW = EN_THREADS / 32;
        Map6 winc = (l < NW) ? s_warp[l] : map_id();
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const Map6 prev = map_shfl_up(winc, o);
            if (l >= (unsigned)o) winc = map_then(prev, winc);
        }
        Map6 wex = map_shfl_up(winc, 1);
        if (l == 0) wex = map_id();
        if (l < NW) s_warp[l] = wex;
        if (l == NW - 1) s_warp[NW] = winc;
    }
    __syncthreads();
    Map6 excl = map_shfl_up(inc, 1);
    if (l == 0) excl = map_id();
    *total = s_warp[EN_THREADS / 32];
    return map_then(s_warp[wid], excl);
}

template <typename T>
__global__ void __launch_bounds__(EN_THREADS)
enc_tile_kernel(const T* __restrict__ vals, uint64_t n, int delta, const EncRange rg, uint32_t* __restrict__ tile_info /*[tiles][6]*/,
                unsigned int* __restrict__ err) {
    __shared__ __align__(16) uint8_t scls[EN_TILE + 16];
    __shared__ uint8_t s_lut[68];
    __shared__ Map6 s_warp[EN_THREADS / 32 + 1];
    __shared__ uint32_t s_cnt[6];
    const unsigned tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * EN_TILE;
    if (tid < 6) s_cnt[tid] = 0;
    enc_prepare<T>(vals, n, delta != 0, base, rg, scls, s_lut, nullptr, err);
    const int lim = (int)min((uint64_t)EN_TILE, n - base);
    const int m = max(0, min(EN_PER, lim - (int)tid * EN_PER));
    const ThreadWalk w = thread_walk(enc_j8(scls), m);
    Map6 tile_map;
    const Map6 excl = block_map_excl_scan(w.map, s_warp, &tile_map);
    // words of the tile for each of the six entry states: my count in the state in which the walk reaches me (a warp
    // starts at most 256 words: 16-bit fields)
    const Map6 c6 = map_lookup(excl, w.cnt);
    const uint32_t s02 = warp_sum(c6.lo & 0x00ff00ffu), s13 = warp_sum((c6.lo >> 8) & 0x00ff00ffu), s45 = warp_sum((c6.hi | (c6.hi << 8)) & 0x00ff00ffu);
    if (lane_id() == 0) {
        atomicAdd(&s_cnt[0], s02 & 0xffffu);
        atomicAdd(&s_cnt[2], s02 >> 16);
        atomicAdd(&s_cnt[1], s13 & 0xffffu);
        atomicAdd(&s_cnt[3], s13 >> 16);
        atomicAdd(&s_cnt[4], s45 & 0xffffu);
        atomicAdd(&s_cnt[5], s45 >> 16);
    }
    __syncthreads();
    if (tid < 6) tile_info[(size_t)blockIdx.x * 6 + tid] = map_at(tile_map, tid) | (s_cnt[tid] << 8);
}

// entry state and first word index of every tile.  One CTA; thread t owns a contiguous chunk of tiles.  A chunk's
// composite -- for each of the six entry states: the exit state (a Map6) and the number of words -- is built in
// registers from whole tile records (six independent loads per tile, no load depends on a previous one), the chunks'
// composites are scanned across the block with shuffles, and every thread replays its chunk from the true entry state.
// (The first version followed each entry state through its chunk with dependent loads and chained the chunks with one
// thread: 112 us per stream for 19,000 tiles, a sixth of the whole encode.)
struct Comp6 {
    Map6 s;           // exit state per entry state
    uint64_t c[6];    // words per entry state (a set of more than 2^32 entries is possible in 180 GB)
};
__device__ __forceinline__ uint64_t pick6(const uint64_t (&c)[6], uint32_t i) {
    return i == 0 ? c[0] : i == 1 ? c[1] : i == 2 ? c[2] : i == 3 ? c[3] : i == 4 ? c[4] : c[5];
}
// a then b
__device__ __forceinline__ Comp6 comp_then(const Comp6& a, const Comp6& b) {
    Comp6 r;
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = a.c[q] + pick6(b.c, map_at(a.s, q));
    r.s = map_then(a.s, b.s);
    return r;
}
__device__ __forceinline__ Comp6 comp_id() {
    Comp6 r;
    r.s = map_id();
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = 0;
    return r;
}
__device__ __forceinline__ Comp6 comp_shfl_up(const Comp6& a, int o) {
    Comp6 r;
    r.s = map_shfl_up(a.s, o);
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = __shfl_up_sync(0xffffffffu, (unsigned long long)a.c[q], o);
    return r;
}
// a tile's record, and the composite of a thread's chunk of tiles, with 32-bit word counts (a chunk holds far fewer than
// 2^32 values): half the selects and adds of the 64-bit form, which only the scan across the block needs
struct Comp6s {
    Map6 s;
    uint32_t c[6];
};
__device__ __forceinline__ uint32_t pick6s(const uint32_t (&c)[6], uint32_t i) {
    return i == 0 ? c[0] : i == 1 ? c[1] : i == 2 ? c[2] : i == 3 ? c[3] : i == 4 ? c[4] : c[5];
}
__device__ __forceinline__ Comp6s comps_of_tile(const uint32_t* __restrict__ info) {
    Comp6s r;
    uint32_t v[6];
#pragma unroll
    for (int q = 0; q < 6; q++) v[q] = info[q];
    r.s.lo = (v[0] & 0xffu) | ((v[1] & 0xffu) << 8) | ((v[2] & 0xffu) << 16) | ((v[3] & 0xffu) << 24);
    r.s.hi = (v[4] & 0xffu) | ((v[5] & 0xffu) << 8);
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = v[q] >> 8;
    return r;
}
__device__ __forceinline__ Comp6s comps_then(const Comp6s& a, const Comp6s& b) {
    Comp6s r;
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = a.c[q] + pick6s(b.c, map_at(a.s, q));
    r.s = map_then(a.s, b.s);
    return r;
}
__device__ __forceinline__ Comp6 comp_of_tile(const uint32_t* __restrict__ info) {
    Comp6 r;
    uint32_t v[6];
#pragma unroll
    for (int q = 0; q < 6; q++) v[q] = info[q];
    r.s.lo = (v[0] & 0xffu) | ((v[1] & 0xffu) << 8) | ((v[2] & 0xffu) << 16) | ((v[3] & 0xffu) << 24);
    r.s.hi = (v[4] & 0xffu) | ((v[5] & 0xffu) << 8);
#pragma unroll
    for (int q = 0; q < 6; q++) r.c[q] = v[q] >> 8;
    return r;
}

// range_map (may be null): [r] = exit state, [6 + r] = number of words, for every entry state r of the whole range
__global__ void __launch_bounds__(1024)
enc_scan_kernel(const uint32_t* __restrict__ tile_info, uint32_t tiles, uint32_t entry0, uint8_t* __restrict__ entry,
                uint64_t* __restrict__ woff, uint64_t* __restrict__ total_words, uint64_t* __restrict__ range_map) {
    __shared__ Comp6 s_warp[33];
    const unsigned t = threadIdx.x, l = t & 31, wid = t >> 5;
    const uint32_t chunk = (tiles + 1023) / 1024;
    const uint32_t t0 = min(tiles, t * chunk), t1 = min(tiles, t0 + chunk);
    Comp6 mine = comp_id();
    if (t0 < t1) {   // the next tile's record is requested before this one is composed
        Comp6s acc = comps_of_tile(tile_info + (size_t)t0 * 6);
        Comp6s nxt = acc;
        if (t0 + 1 < t1) nxt = comps_of_tile(tile_info + (size_t)(t0 + 1) * 6);
        for (uint32_t i = t0 + 1; i < t1; i++) {
            const Comp6s cur = nxt;
            if (i + 1 < t1) nxt = comps_of_tile(tile_info + (size_t)(i + 1) * 6);
            acc = comps_then(acc, cur);
        }
        mine.s = acc.s;
#pragma unroll
        for (int q = 0; q < 6; q++) mine.c[q] = acc.c[q];
    }
    // inclusive scan over the block
    Comp6 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Comp6 prev = comp_shfl_up(inc, o);
        if (l >= (unsigned)o) inc = comp_then(prev, inc);
    }
    if (l == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Comp6 winc = s_warp[l];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Comp6 prev = comp_shfl_up(winc, o);
            if (l >= (unsigned)o) winc = comp_then(prev, winc);
        }
        Comp6 wex = comp_shfl_up(winc, 1);
        if (l == 0) wex = comp_id();
        s_warp[l] = wex;
        if (l == 31) s_warp[32] = winc;
    }
    __syncthreads();
    Comp6 excl = comp_shfl_up(inc, 1);
    if (l == 0) excl = comp_id();
    excl = comp_then(s_warp[wid], excl);
    if (t == 0) {
        const Comp6 all = s_warp[32];
        if (range_map) {
#pragma unroll
            for (int q = 0; q < 6; q++) { range_map[q] = map_at(all.s, q); range_map[6 + q] = all.c[q]; }
        }
        *total_words = pick6(all.c, entry0);
    }
    // replay my chunk from the state in which the walk from entry0 reaches it
    uint32_t st = map_at(excl.s, entry0);
    uint64_t off = pick6(excl.c, entry0);
    for (uint32_t i = t0; i < t1; i++) {
        const Comp6s ti = comps_of_tile(tile_info + (size_t)i * 6);
        entry[i] = (uint8_t)st;
        woff[i] = off;
        off += pick6s(ti.c, st);
        st = map_at(ti.s, st);
    }
}

template <typename T>
__global__ void __launch_bounds__(EN_THREADS)
enc_emit_kernel(const T* __restrict__ vals, uint64_t n, int delta, const EncRange rg, const uint8_t* __restrict__ entry,
                const uint64_t* __restrict__ woff, uint64_t* __restrict__ words, unsigned int* __restrict__ err) {
    __shared__ __align__(16) uint8_t scls[EN_TILE + 16];
    __shared__ uint8_t s_lut[68];
    __shared__ uint64_t sv[EN_SV_SIZE];
    __shared__ Map6 s_warp[EN_THREADS / 32 + 1];
    __shared__ uint32_t s_scan[EN_THREADS / 32 + 1];
    const unsigned tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * EN_TILE;
    enc_prepare<T>(vals, n, delta != 0, base, rg, scls, s_lut, sv, err);
    const int lim = (int)min((uint64_t)EN_TILE, n - base);
    const int m = max(0, min(EN_PER, lim - (int)tid * EN_PER));
    const uint64_t j8 = enc_j8(scls);
    const ThreadWalk w = thread_walk(j8, m);
    Map6 tile_map;
    const Map6 excl = block_map_excl_scan(w.map, s_warp, &tile_map);
    // the state in which the walk from the tile's true entry state reaches this thread; its word starts
    const int st = (int)map_at(excl, (uint32_t)entry[blockIdx.x]);
    uint32_t tot;
    uint32_t rank = block_excl_scan<EN_THREADS, uint32_t, false>(map_at(w.cnt, (uint32_t)st), s_scan, &tot);
    const uint64_t wbase = woff[blockIdx.x];
    int p = st;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        if (p < m) {
            const int g = (int)((j8 >> (8 * p)) & 0xffu);
            const int wd = cw_width(g);
            const int q0 = (int)tid * EN_PER + p;
            uint64_t word = 0;
            for (int q = g - 1; q >= 0; q--) word = (word << wd) | sv[EN_SV(q0 + q)];
            words[wbase + rank] = (word << 4) | (uint64_t)g;
            rank++;
            p += g;
        }
    }
}

// ------------------------------------------------------------------------------------------- decode
static constexpr int DE_THREADS = 128;
static constexpr int DE_PER = 4;
static constexpr int DE_TILE = DE_THREADS * DE_PER;   // 512 words -> at most 3072 values

__device__ __forceinline__ uint64_t word_sum(uint64_t w, int g) {
    const int wd = cw_width(g);
    const uint64_t msk = (1ull << wd) - 1ull;
    uint64_t s = 0;
    w >>= 4;
    for (int q = 0; q < g; q++) { s += w & msk; w >>= wd; }
    return s;
}

__global__ void __launch_bounds__(DE_THREADS)
dec_tile_kernel(const uint64_t* __restrict__ words, uint64_t nw, uint32_t* __restrict__ tile_cnt,
                uint64_t* __restrict__ tile_sum, unsigned int* __restrict__ err) {
    __shared__ uint32_t s_c[DE_THREADS / 32];
    __shared__ uint64_t s_s[DE_THREADS / 32];
    const unsigned tid = threadIdx.x;
    uint32_t c = 0;
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < DE_PER; j++) {
        const uint64_t i = (uint64_t)blockIdx.x * DE_TILE + j * DE_THREADS + tid;
        if (i < nw) {
            const uint64_t w = words[i];
            const int g = (int)(w & 15);
            if (g < 1 || g > 6) atomicExch(err, 1u);   // codec64.py:128 KeyError on an unknown tag
            else { c += g; sum += word_sum(w, g); }
        }
    }
    c = warp_sum(c);
    sum = warp_sum(sum);
    if ((tid & 31) == 0) { s_c[tid >> 5] = c; s_s[tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < DE_THREADS / 32; q++) { c += s_c[q]; sum += s_s[q]; }
        tile_cnt[blockIdx.x] = c;
        tile_sum[blockIdx.x] = sum;
    }
}

__global__ void __launch_bounds__(1024)
dec_scan_kernel(const uint32_t* __restrict__ tile_cnt, const uint64_t* __restrict__ tile_sum, uint32_t tiles,
                uint64_t* __restrict__ cnt_off, uint64_t* __restrict__ sum_off, uint64_t* __restrict__ total) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    uint64_t ccarry = 0, scarry = 0;
    for (uint32_t b0 = 0; b0 < tiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t c = (i < tiles) ? tile_cnt[i] : 0, s = (i < tiles) ? tile_sum[i] : 0;
        uint64_t ctot, stot;
        const uint64_t cex = block_excl_scan<1024, uint64_t>(c, sm, &ctot);
        const uint64_t sex = block_excl_scan<1024, uint64_t>(s, sm, &stot);
        if (i < tiles) { cnt_off[i] = ccarry + cex; sum_off[i] = scarry + sex; }
        ccarry += ctot;
        scarry += stot;
    }
    if (threadIdx.x == 0) *total = ccarry;
}

template <typename T>
__global__ void __launch_bounds__(DE_THREADS)
dec_emit_kernel(const uint64_t* __restrict__ words, uint64_t nw, int delta, const uint64_t* __restrict__ cnt_off,
                const uint64_t* __restrict__ sum_off, T* __restrict__ out, unsigned int* __restrict__ err) {
    __shared__ uint64_t sval[DE_TILE * 6];
    __shared__ uint32_t s_scan[DE_THREADS / 32 + 1];
    __shared__ uint64_t s_scan64[DE_THREADS / 32 + 1];
    const unsigned tid = threadIdx.x;
    // thread t owns the 4 consecutive words 4t .. 4t+3 of the tile
    uint64_t w[DE_PER];
    uint32_t c = 0;
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < DE_PER; j++) {
        const uint64_t i = (uint64_t)blockIdx.x * DE_TILE + tid * DE_PER + j;
        w[j] = (i < nw) ? words[i] : 0ull;
        const int g = (int)(w[j] & 15);
        if (g >= 1 && g <= 6) { c += g; sum += word_sum(w[j], g); } else w[j] = 0;
    }
    uint32_t ctot;
    uint64_t stot;
    uint32_t o = block_excl_scan<DE_THREADS, uint32_t>(c, s_scan, &ctot);
    uint64_t acc = block_excl_scan<DE_THREADS, uint64_t>(sum, s_scan64, &stot);
    if (delta) acc += sum_off[blockIdx.x];
#pragma unroll
    for (int j = 0; j < DE_PER; j++) {
        const int g = (int)(w[j] & 15);
        if (g) {
            const int wd = cw_width(g);
            const uint64_t msk = (1ull << wd) - 1ull;
            uint64_t x = w[j] >> 4;
            for (int q = 0; q < g; q++) {
                const uint64_t v = x & msk;
                x >>= wd;
                if (delta) { acc += v; sval[o++] = acc; } else sval[o++] = v;
            }
        }
    }
    __syncthreads();
    const uint64_t dst = cnt_off[blockIdx.x];
    for (uint32_t i = tid; i < ctot; i += DE_THREADS) {
        const uint64_t v = sval[i];
        if (sizeof(T) == 4 && v > 0xffffffffull) atomicExch(err, 2u);
        out[dst + i] = (T)v;
    }
}

// ------------------------------------------------------------------------------------------- host side
// Encoding in two phases, so that several GPUs can encode consecutive ranges of ONE stream: encode_plan computes the
// per-tile maps and the range's own map (exit state and word count for each of the six entry states); the caller
// chains the maps of the ranges on the host and hands every range its true entry state; encode_emit packs the words.
struct EncodePlan {
    uint32_t tiles = 0;
    size_t n = 0;
    bool delta = false;
    EncRange rg;
    DBuf<uint32_t> info;
    DBuf<uint8_t> entry;
    DBuf<uint64_t> woff;      // [tiles] word offsets, then [0] total words, [1] error flag, [2..13] range map
    uint32_t map_exit[6];
    uint64_t map_words[6];
};

// [0] total words, [1] error flag, [2..13] range map -- behind the word offsets, on a 16-byte boundary
static inline uint64_t* enc_tail(EncodePlan* p) { return p->woff.get() + (((size_t)p->tiles + 1) & ~(size_t)1); }

template <typename T>
static void encode_plan(Ctx* c, const T* d_vals, size_t n, bool delta, const EncRange& rg, EncodePlan* p) {
    p->n = n;
    p->delta = delta;
    p->rg = rg;
    p->tiles = (uint32_t)div_up(n, EN_TILE);
    if (n == 0) {   // an empty range passes its entry state on
        for (int r = 0; r < 6; r++) { p->map_exit[r] = r; p->map_words[r] = 0; }
        return;
    }
    p->info.alloc(c, (size_t)p->tiles * 6);
    p->entry.alloc(c, p->tiles);
    p->woff.alloc(c, (size_t)p->tiles + 16);
    uint64_t* total = enc_tail(p);
    unsigned int* err = reinterpret_cast<unsigned int*>(total + 1);
    ZB_CUDA(dev_memset(c, total, 0, 16));
    enc_tile_kernel<T><<<p->tiles, EN_THREADS, 0, c->stream>>>(d_vals, n, delta ? 1 : 0, rg, p->info.get(), err);
    ZB_LAUNCH_CHECK(c);
    enc_scan_kernel<<<1, 1024, 0, c->stream>>>(p->info.get(), p->tiles, 0u, p->entry.get(), p->woff.get(), total, total + 2);
    ZB_LAUNCH_CHECK(c);
}

// the range's map (synchronises); only needed when ranges are chained
static void encode_plan_map(Ctx* c, EncodePlan* p) {
    if (p->n == 0) return;
    uint64_t* total = enc_tail(p);
    ZB_CUDA(read_back(c, total + 2, 96));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < 6; r++) { p->map_exit[r] = (uint32_t)c->h_scalars[r]; p->map_words[r] = c->h_scalars[6 + r]; }
}

// the same for the two streams of a set with ONE synchronisation (every host round trip of a step is ~15 us of idle GPU)
static void encode_plan_maps2(Ctx* c, EncodePlan* a, EncodePlan* b) {
    if (a->n == 0 || b->n == 0) { encode_plan_map(c, a); encode_plan_map(c, b); return; }
    // the maps sit 16 bytes into 16-byte aligned tails: copy [total, err, map] = 112 bytes of each plan
    copy_small_to_host(c, enc_tail(a), 0, 112);
    copy_small_to_host(c, enc_tail(b), 112, 112);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t* ha = reinterpret_cast<const uint64_t*>(c->h_big);
    const uint64_t* hb = reinterpret_cast<const uint64_t*>(c->h_big + 112);
    for (int r = 0; r < 6; r++) {
        a->map_exit[r] = (uint32_t)ha[2 + r]; a->map_words[r] = ha[8 + r];
        b->map_exit[r] = (uint32_t)hb[2 + r]; b->map_words[r] = hb[8 + r];
    }
}

// words (device, capacity n); entry_state = 0 for a whole stream; returns the number of words written
template <typename T>
static size_t encode_emit(Ctx* c, const T* d_vals, EncodePlan* p, uint32_t entry_state, uint64_t* d_words) {
    if (p->n == 0) return 0;
    uint64_t* total = enc_tail(p);
    unsigned int* err = reinterpret_cast<unsigned int*>(total + 1);
    if (entry_state != 0) {   // the plan's scan ran with entry state 0
        enc_scan_kernel<<<1, 1024, 0, c->stream>>>(p->info.get(), p->tiles, entry_state, p->entry.get(), p->woff.get(), total, nullptr);
        ZB_LAUNCH_CHECK(c);
    }
    enc_emit_kernel<T><<<p->tiles, EN_THREADS, 0, c->stream>>>(d_vals, p->n, p->delta ? 1 : 0, p->rg, p->entry.get(), p->woff.get(), d_words, err);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, total, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 1)[0] != 0)
        ZB_FAIL(ZB_E_RANGE, "codec64: value or k-mer gap needs more than 60 bits (reference: IndexError, codec64.py:93-99)");
    return (size_t)c->h_scalars[0];
}

// the two streams of a set (entry state 0), emitted back to back with ONE synchronisation
static void encode_emit2(Ctx* c, const uint64_t* d_k, EncodePlan* pk, uint64_t* kw, size_t* nk, const uint32_t* d_c, EncodePlan* pc,
                         uint64_t* cw, size_t* nc) {
    uint64_t* tk = enc_tail(pk);
    uint64_t* tc = enc_tail(pc);
    enc_emit_kernel<uint64_t><<<pk->tiles, EN_THREADS, 0, c->stream>>>(d_k, pk->n, 1, pk->rg, pk->entry.get(), pk->woff.get(), kw,
                                                                      reinterpret_cast<unsigned int*>(tk + 1));
    ZB_LAUNCH_CHECK(c);
    enc_emit_kernel<uint32_t><<<pc->tiles, EN_THREADS, 0, c->stream>>>(d_c, pc->n, 0, pc->rg, pc->entry.get(), pc->woff.get(), cw,
                                                                      reinterpret_cast<unsigned int*>(tc + 1));
    ZB_LAUNCH_CHECK(c);
    copy_small_to_host(c, tk, 0, 16);
    copy_small_to_host(c, tc, 16, 16);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t* h = reinterpret_cast<const uint64_t*>(c->h_big);
    if ((uint32_t)h[1] != 0 || (uint32_t)h[3] != 0)
        ZB_FAIL(ZB_E_RANGE, "codec64: value or k-mer gap needs more than 60 bits (reference: IndexError, codec64.py:93-99)");
    *nk = (size_t)h[0];
    *nc = (size_t)h[2];
}

// vals (device, n values) -> words (device, capacity n); returns the number of words
template <typename T>
static size_t encode_dev(Ctx* c, const T* d_vals, size_t n, bool delta, uint64_t* d_words) {
    if (n == 0) return 0;
    EncodePlan p;
    EncRange rg;
    memset(&rg, 0, sizeof rg);
    encode_plan<T>(c, d_vals, n, delta, rg, &p);
    return encode_emit<T>(c, d_vals, &p, 0, d_words);
}

struct DecodePlan {
    uint32_t tiles = 0;
    DBuf<uint32_t> tcnt;
    DBuf<uint64_t> tsum, coff, soff;
    size_t n = 0;
};

// pass 1 of a decode: number of values in the stream (validates the tags)
static void decode_plan(Ctx* c, const uint64_t* d_words, size_t nw, DecodePlan* p) {
    p->n = 0;
    if (nw == 0) return;
    p->tiles = (uint32_t)div_up(nw, DE_TILE);
    p->tcnt.alloc(c, p->tiles);
    p->tsum.alloc(c, p->tiles);
    p->coff.alloc(c, (size_t)p->tiles + 2);
    p->soff.alloc(c, p->tiles);
    uint64_t* total = p->coff.get() + p->tiles;
    unsigned int* err = reinterpret_cast<unsigned int*>(p->coff.get() + p->tiles + 1);
    ZB_CUDA(dev_memset(c, total, 0, 16));
    dec_tile_kernel<<<p->tiles, DE_THREADS, 0, c->stream>>>(d_words, nw, p->tcnt.get(), p->tsum.get(), err);
    ZB_LAUNCH_CHECK(c);
    dec_scan_kernel<<<1, 1024, 0, c->stream>>>(p->tcnt.get(), p->tsum.get(), p->tiles, p->coff.get(), p->soff.get(), total);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, total, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 1)[0] != 0)
        ZB_FAIL(ZB_E_FORMAT, "codec64: corrupt stream (tag outside 1..6; reference: KeyError, codec64.py:128)");
    p->n = (size_t)c->h_scalars[0];
}

template <typename T>
static void decode_emit(Ctx* c, const uint64_t* d_words, size_t nw, bool delta, DecodePlan* p, T* d_out) {
    if (nw == 0 || p->n == 0) return;
    unsigned int* err = reinterpret_cast<unsigned int*>(p->coff.get() + p->tiles + 1);
    dec_emit_kernel<T><<<p->tiles, DE_THREADS, 0, c->stream>>>(d_words, nw, delta ? 1 : 0, p->coff.get(), p->soff.get(), d_out, err);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, err, 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (reinterpret_cast<uint32_t*>(c->h_scalars)[0] != 0)
        ZB_FAIL(ZB_E_RANGE, "count exceeds 2^32-1 (reference: array('I') OverflowError, kmerize.py:374)");
}

}  // namespace zb

using namespace zb;

extern "C" {

int zb_encode_u64_stream(int device, const uint64_t* vals, size_t n, int delta, uint64_t* words, size_t* n_words) {
    ZB_TRY
    if ((n && (!vals || !words)) || !n_words) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    *n_words = 0;
    if (n == 0) return ZB_OK;
    DBuf<uint64_t> dv(c, n), dw(c, n);
    ZB_CUDA(cudaMemcpyAsync(dv.get(), vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    const size_t nw = encode_dev<uint64_t>(c, dv.get(), n, delta != 0, dw.get());
    ZB_CUDA(cudaMemcpyAsync(words, dw.get(), nw * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    *n_words = nw;
    ZB_CATCH
}

int zb_decode_u64_stream(int device, const uint64_t* words, size_t n_words, int delta, uint64_t* out, size_t* n) {
    ZB_TRY
    if ((n_words && !words) || !n) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    *n = 0;
    if (n_words == 0) return ZB_OK;
    DBuf<uint64_t> dw(c, n_words);
    ZB_CUDA(cudaMemcpyAsync(dw.get(), words, n_words * 8, cudaMemcpyHostToDevice, c->stream));
    DecodePlan p;
    decode_plan(c, dw.get(), n_words, &p);
    *n = p.n;
    if (out && p.n) {
        DBuf<uint64_t> dout(c, p.n);
        decode_emit<uint64_t>(c, dw.get(), n_words, delta != 0, &p, dout.get());
        ZB_CUDA(cudaMemcpyAsync(out, dout.get(), p.n * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    ZB_CATCH
}

int zb_set_encode_sizes(const zb_set* s, size_t* n_kmer_words, size_t* n_count_words) {
    // upper bounds (one value per word); exact sizes come back from zb_set_encode
    size_t n = 0;
    int rc = zb_set_size(s, &n);
    if (rc) return rc;
    if (n_kmer_words) *n_kmer_words = n;
    if (n_count_words) *n_count_words = n;
    return ZB_OK;
}

int zb_set_encode(const zb_set* s, uint64_t* kmer_words, size_t* n_kmer_words, uint64_t* count_words, size_t* n_count_words) {
    ZB_TRY
    if (!s || !n_kmer_words || !n_count_words) ZB_FAIL(ZB_E_ARG, "null argument");
    *n_kmer_words = *n_count_words = 0;
    if (s->n == 0) return ZB_OK;
    if (!kmer_words || !count_words) ZB_FAIL(ZB_E_ARG, "null argument");
    zb_words* w = nullptr;
    if (int rc = zb_set_encode_dev(s, &w)) return rc;
    *n_kmer_words = w->nk;
    *n_count_words = w->nc;
    const int rc = zb_words_fetch(w, kmer_words, count_words);
    zb_words_free(w);
    return rc;
    ZB_CATCH
}

// the two packed streams of a set, left in HBM: codec64 kernels only, no PCIe traffic
int zb_set_encode_dev(const zb_set* s, zb_words** out) {
    ZB_TRY
    if (!s || !out) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_words* w = new zb_words();
    w->c = c;
    try {
        EngineLock el(c, ENG_SM);
        Stage st(c, "encode");
        if (s->n) {
            // plan first (tile maps + scan): it knows the number of words, so the streams are allocated at their real size
            // (k-mer gaps of a 10 M set are ~27 bits: two per word; counts: six per word)
            EncRange rg;
            memset(&rg, 0, sizeof rg);
            if (s->wide.get()) {
                // counts beyond 2^32-1 (a merge of very deep sets): both streams through the 64-bit encoder, one after the other
                w->kw.alloc(c, s->n);
                w->cw.alloc(c, s->n);
                w->nk = encode_dev<uint64_t>(c, s->k.get(), s->n, true, w->kw.get());
                w->nc = encode_dev<uint64_t>(c, s->wide.get(), s->n, false, w->cw.get());
                *out = w;
                return ZB_OK;
            }
            EncodePlan pk, pc;
            encode_plan<uint64_t>(c, s->k.get(), s->n, true, rg, &pk);
            encode_plan<uint32_t>(c, s->cnt.get(), s->n, false, rg, &pc);
            encode_plan_maps2(c, &pk, &pc);
            w->kw.alloc(c, pk.map_words[0] + 1);
            w->cw.alloc(c, pc.map_words[0] + 1);
            encode_emit2(c, s->k.get(), &pk, w->kw.get(), &w->nk, s->cnt.get(), &pc, w->cw.get(), &w->nc);
        }
    } catch (...) {
        delete w;
        throw;
    }
    *out = w;
    ZB_CATCH
}

struct zb_encplan {
    const zb_set* s;
    EncodePlan pk, pc;
};

int zb_set_encode_plan(const zb_set* s, uint64_t prev_kmer, const uint64_t* next_kmers, const uint32_t* next_counts, int n_next,
                       zb_encplan** out, uint64_t kmer_map[12], uint64_t count_map[12]) {
    ZB_TRY
    if (!s || !out || !kmer_map || !count_map || n_next < 0 || n_next > 5 || (n_next && (!next_kmers || !next_counts)))
        ZB_FAIL(ZB_E_ARG, "bad argument");
    if (s->wide.get()) ZB_FAIL(ZB_E_RANGE, "a set with counts beyond 2^32-1 is encoded in one piece (zb_set_encode_dev)");
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_encplan* p = new zb_encplan();
    p->s = s;
    try {
        EngineLock el(c, ENG_SM);
        Stage st(c, "encode");
        EncRange rk, rc;
        memset(&rk, 0, sizeof rk);
        memset(&rc, 0, sizeof rc);
        rk.first_base = prev_kmer;
        rk.n_halo = rc.n_halo = (uint32_t)n_next;
        for (int i = 0; i < n_next; i++) { rk.halo[i] = next_kmers[i]; rc.halo[i] = next_counts[i]; }
        encode_plan<uint64_t>(c, s->k.get(), s->n, true, rk, &p->pk);
        encode_plan<uint32_t>(c, s->cnt.get(), s->n, false, rc, &p->pc);
        encode_plan_map(c, &p->pk);
        encode_plan_map(c, &p->pc);
    } catch (...) {
        delete p;
        throw;
    }
    for (int r = 0; r < 6; r++) {
        kmer_map[r] = p->pk.map_exit[r]; kmer_map[6 + r] = p->pk.map_words[r];
        count_map[r] = p->pc.map_exit[r]; count_map[6 + r] = p->pc.map_words[r];
    }
    *out = p;
    ZB_CATCH
}

int zb_set_encode_emit(zb_encplan* p, int kmer_entry_state, int count_entry_state, zb_words** out) {
    ZB_TRY
    if (!p || !out || kmer_entry_state < 0 || kmer_entry_state > 5 || count_entry_state < 0 || count_entry_state > 5)
        ZB_FAIL(ZB_E_ARG, "bad argument");
    const zb_set* s = p->s;
    Ctx* c = s->c;
    ZB_CUDA(cudaSetDevice(c->device));
    zb_words* w = new zb_words();
    w->c = c;
    try {
        EngineLock el(c, ENG_SM);
        Stage st(c, "encode");
        if (s->n) {
            w->kw.alloc(c, p->pk.map_words[kmer_entry_state] + 1);
            w->nk = encode_emit<uint64_t>(c, s->k.get(), &p->pk, (uint32_t)kmer_entry_state, w->kw.get());
            w->cw.alloc(c, p->pc.map_words[count_entry_state] + 1);
            w->nc = encode_emit<uint32_t>(c, s->cnt.get(), &p->pc, (uint32_t)count_entry_state, w->cw.get());
            if (w->nk != p->pk.map_words[kmer_entry_state] || w->nc != p->pc.map_words[count_entry_state])
                ZB_FAIL(ZB_E_CUDA, "codec64: the emitted word count differs from the plan's");
        }
    } catch (...) {
        delete w;
        delete p;
        throw;
    }
    delete p;
    *out = w;
    ZB_CATCH
}

int zb_encplan_free(zb_encplan* p) {
    ZB_TRY
    if (p) {
        cudaSetDevice(p->s->c->device);
        delete p;
    }
    ZB_CATCH
}

int zb_words_sizes(const zb_words* w, size_t* n_kmer_words, size_t* n_count_words) {
    if (!w) { zb::set_error("null argument"); return ZB_E_ARG; }
    if (n_kmer_words) *n_kmer_words = w->nk;
    if (n_count_words) *n_count_words = w->nc;
    return ZB_OK;
}

int zb_words_dev_ptrs(const zb_words* w, const uint64_t** d_kmer_words, const uint64_t** d_count_words) {
    if (!w) { zb::set_error("null argument"); return ZB_E_ARG; }
    if (d_kmer_words) *d_kmer_words = w->kw.get();
    if (d_count_words) *d_count_words = w->cw.get();
    return ZB_OK;
}

int zb_words_fetch(const zb_words* w, uint64_t* kmer_words, uint64_t* count_words) {
    ZB_TRY
    if (!w) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = w->c;
    ZB_CUDA(cudaSetDevice(c->device));
    EngineLock el(c, ENG_D2H);
    Stage st(c, "d2h_words");
    if (w->nk && kmer_words) copy_chunked(c, kmer_words, w->kw.get(), w->nk * 8, cudaMemcpyDeviceToHost);
    if (w->nc && count_words) copy_chunked(c, count_words, w->cw.get(), w->nc * 8, cudaMemcpyDeviceToHost);
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_words_free(zb_words* w) {
    ZB_TRY
    if (w) {
        cudaSetDevice(w->c->device);
        delete w;
    }
    ZB_CATCH
}

// words already on the device (e.g. staged there by the I/O threads) -> set
int zb_set_from_streams_dev(int device, const uint64_t* d_kmer_words, size_t n_kmer_words, const uint64_t* d_count_words,
                            size_t n_count_words, zb_set** out) {
    ZB_TRY
    if (!out || (n_kmer_words && !d_kmer_words)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    Stage st(c, "decode");
    DecodePlan pk, pc;
    decode_plan(c, d_kmer_words, n_kmer_words, &pk);
    if (d_count_words) {
        decode_plan(c, d_count_words, n_count_words, &pc);
        if (pc.n != pk.n)
            ZB_FAIL(ZB_E_FORMAT, "k-mer and count streams differ in length (%zu vs %zu)", pk.n, pc.n);   // files.py:182 assert
    }
    // empty template set, then fill its arrays in place
    zb_set* s = nullptr;
    if (int rc = zb_set_from_host(device, nullptr, nullptr, 0, &s)) return rc;
    zb_set* v = s;
    try {
        v->k.alloc(c, pk.n);
        v->cnt.alloc(c, pk.n);
        v->n = pk.n;
        decode_emit<uint64_t>(c, d_kmer_words, n_kmer_words, true, &pk, v->k.get());
        if (d_count_words) {
            try {
                decode_emit<uint32_t>(c, d_count_words, n_count_words, false, &pc, v->cnt.get());
            } catch (const zb::Fail& f) {
                if (f.code != ZB_E_RANGE) throw;
                // counts beyond 2^32-1 (the file of a merge of very deep sets, merge.py:145-146): a WIDE set
                ZB_CUDA(dev_memset(c, pc.coff.get() + pc.tiles + 1, 0, 4));
                v->wide.alloc(c, pk.n);
                decode_emit<uint64_t>(c, d_count_words, n_count_words, false, &pc, v->wide.get());
                zb_set_finish_wide(v);
            }
        } else if (pk.n) {
            fill_u32(c, v->cnt.get(), pk.n, 1u);
        }
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        zb_set_free(s);
        throw;
    }
    *out = s;
    ZB_CATCH
}

int zb_set_from_streams(int device, const uint64_t* kmer_words, size_t n_kmer_words, const uint64_t* count_words,
                        size_t n_count_words, zb_set** out) {
    ZB_TRY
    if (!out || (n_kmer_words && !kmer_words)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    DBuf<uint64_t> dkw(c, n_kmer_words), dcw;
    if (n_kmer_words) ZB_CUDA(cudaMemcpyAsync(dkw.get(), kmer_words, n_kmer_words * 8, cudaMemcpyHostToDevice, c->stream));
    if (count_words) {
        dcw.alloc(c, n_count_words);
        if (n_count_words) ZB_CUDA(cudaMemcpyAsync(dcw.get(), count_words, n_count_words * 8, cudaMemcpyHostToDevice, c->stream));
    }
    return zb_set_from_streams_dev(device, dkw.get(), n_kmer_words, count_words ? dcw.get() : nullptr, n_count_words, out);
    ZB_CATCH
}

}  // extern "C"
