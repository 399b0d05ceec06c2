// extract.cu -- 2-bit packing and canonical k-mer extraction from a dense base-code stream.
//
// Replaces zotmer/library/basics.py:303-347 (kmersList(K, seq, True)): every window of k consecutive
// valid bases yields the forward k-mer x and its reverse complement xb.  Because count(x) ==
// count(rc(x)) for every k-mer the reference ever emits, the device counts only the CANONICAL key
// min(x, xb) (half the sort volume) and mirrors the distinct result afterwards (setops.cu
// mirror_keys); the emitted set is the reference's both-strand multiset, bit for bit.
//
// Input layout (produced by parse.cu): one byte per base, 0..3 = A C G T/U, 4 = break (invalid byte
// or record boundary), preceded by 32 break bytes and padded with break bytes to a whole tile.
// One CTA handles 4096 codes: a single 1-D bulk async copy (TMA, UBLKCP) brings the tile plus its
// 32-code halo into shared memory; each thread packs 16 codes into a little-endian 2-bit word and a
// 16-bit invalid mask; neighbours' words come from shared memory; all 16 windows of a thread are
// then plain shifts of a 96-bit register window (no per-base rolling, no divergence).
#include "kernels.h"

namespace zb {

static constexpr int EX_THREADS = 256;
static constexpr int EX_PER = 16;
static_assert(EX_THREADS * EX_PER == EXTRACT_TILE, "tile shape");

__device__ __forceinline__ uint64_t rev2_64(uint64_t x) {
    uint64_t y = __brevll(x);
    return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
}

// 4 byte codes (0..4) -> 8 bits of 2-bit codes (first byte lowest) and 4 invalid bits
__device__ __forceinline__ void pack4(uint32_t v, uint32_t& w, uint32_t& m) {
    uint32_t t = v & 0x03030303u;
    t = (t | (t >> 6)) & 0x000f000fu;
    t = (t | (t >> 12)) & 0xffu;
    uint32_t u = (v >> 2) & 0x01010101u;
    u = (u | (u >> 7) | (u >> 14) | (u >> 21)) & 0xfu;
    w = t;
    m = u;
}
__device__ __forceinline__ void pack16(uint4 c, uint32_t& w, uint32_t& m) {
    uint32_t w0, w1, w2, w3, m0, m1, m2, m3;
    pack4(c.x, w0, m0);
    pack4(c.y, w1, m1);
    pack4(c.z, w2, m2);
    pack4(c.w, w3, m3);
    w = w0 | (w1 << 8) | (w2 << 16) | (w3 << 24);
    m = m0 | (m1 << 4) | (m2 << 8) | (m3 << 12);
}

// owner of a canonical k-mer among nranks GPUs (multi-GPU exchange, further down): high bits of an invertible 64-bit mix
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
__device__ __forceinline__ int owner_of(uint64_t key, int nranks) {
    return (int)__umul64hi(mix64(key), (uint64_t)nranks);
}

// OWNERS: the kernel also tallies how many of its keys each of `nranks` owners will get (owner_counts[nranks], global), so
// that the exchange needs no counting pass over the 1 GB of keys afterwards (0.25 ms of every multi-GPU step)
// HIST: the kernel also tallies the digits the coming sort will look at (SortPre: up to 4 digits of at most 8 bits) in
// shared memory and adds them to hist[passes][256] -- the sort then needs no pass over the keys for its histograms
struct ExHist {
    int passes;
    int shift[4];
    uint32_t mask[4];
    uint32_t* hist;
};

// HIST: 0 = no histograms; 2 / 3 = exactly that many digits, all in the upper 32 bits of the key (the top-bit passes of a
// k >= 17 batch: one 32-bit shift per digit, no predicates); 4 = the general form (up to 4 digits anywhere in the key).
// The tallies were 22 % of this kernel's instructions in the general form (profiles/r02_ncu_extract_hist.txt).
template <bool OWNERS, int HIST>
__global__ void __launch_bounds__(EX_THREADS)
extract_kernel(int k, const uint8_t* __restrict__ codes, uint64_t* __restrict__ out,
               unsigned long long* __restrict__ counter, int nranks, unsigned long long* __restrict__ owner_counts, const ExHist eh) {
    __shared__ __align__(128) uint8_t s_codes[EXTRACT_TILE + 32];
    __shared__ uint32_t s_w[EX_THREADS + 2];
    __shared__ uint32_t s_m[EX_THREADS + 2];
    __shared__ __align__(16) uint64_t s_keys[EXTRACT_TILE];
    __shared__ uint32_t s_scan[EX_THREADS / 32 + 1];
    __shared__ unsigned long long s_base;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_own[OWNERS ? 64 : 1];
    __shared__ uint32_t s_hist[HIST ? 4 * 256 : 1];

    const unsigned tid = threadIdx.x;
    if (OWNERS && tid < 64) s_own[tid] = 0;
    if (HIST) {
#pragma unroll
        for (int p = 0; p < 4; p++) s_hist[p * 256 + tid] = 0;
    }
    const uint8_t* src = codes + (size_t)blockIdx.x * EXTRACT_TILE - 32;  // 16-byte aligned by construction
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, EXTRACT_TILE + 32);
        bulk_g2s(s_codes, src, EXTRACT_TILE + 32, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    uint32_t w, m;
    pack16(*reinterpret_cast<const uint4*>(s_codes + 32 + tid * 16), w, m);
    s_w[tid + 2] = w;
    s_m[tid + 2] = m;
    if (tid < 2) {
        uint32_t hw, hm;
        pack16(*reinterpret_cast<const uint4*>(s_codes + tid * 16), hw, hm);
        s_w[tid] = hw;
        s_m[tid] = hm;
    }
    __syncthreads();
    const uint32_t w2 = s_w[tid], w1 = s_w[tid + 1];  // bases -32..-17, -16..-1 relative to my first base
    const uint32_t m2 = s_m[tid], m1 = s_m[tid + 1];

    // 96-bit little-endian window: base i of the 48 sits at bits [2i, 2i+2)
    const uint64_t lo = ((uint64_t)w1 << 32) | w2;
    const uint64_t hi = w;
    const int s = 2 * (33 - k);  // 2..64: uniform pre-shift so that window j starts at bit 2j
    uint64_t plo, phi;
    if (s >= 64) { plo = hi; phi = 0; }
    else { plo = (lo >> s) | (hi << (64 - s)); phi = hi >> s; }
    const uint64_t inv48 = ((uint64_t)m << 32) | ((uint64_t)m1 << 16) | m2;
    const uint64_t pinv = inv48 >> (33 - k);
    const uint64_t kmask = (k == 32) ? 0xffffffffull : ((1ull << k) - 1ull);
    const uint64_t msk = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const int fsh = 64 - 2 * k;

    uint64_t key[EX_PER];
    uint32_t vbits = 0;
#pragma unroll
    for (int j = 0; j < EX_PER; j++) {
        const uint64_t win = (j == 0) ? plo : ((plo >> (2 * j)) | (phi << (64 - 2 * j)));
        const bool ok = ((pinv >> j) & kmask) == 0;
        const uint64_t rcv = (~win) & msk;          // reverse complement: complement of the LE window
        const uint64_t fwd = rev2_64(win) >> fsh;   // forward: first base most significant
        key[j] = fwd < rcv ? fwd : rcv;
        vbits |= (ok ? 1u : 0u) << j;
    }
    uint32_t tot;
    uint32_t off = block_excl_scan<EX_THREADS, uint32_t>(__popc(vbits), s_scan, &tot);
#pragma unroll
    for (int j = 0; j < EX_PER; j++) {
        if ((vbits >> j) & 1u) s_keys[off++] = key[j];
    }
    if (tid == 0) s_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
    __syncthreads();
    const unsigned long long base = s_base;
    // up to 8 owners: a thread tallies its ~13 keys in two registers of four 16-bit fields (no atomics, no indexing),
    // the warp adds them up; more owners: one shared-memory atomic per key
    uint64_t alo = 0, ahi = 0;
    for (uint32_t i = tid; i < tot; i += EX_THREADS) {
        const uint64_t kx = s_keys[i];
        out[base + i] = kx;
        if (HIST == 2 || HIST == 3) {
            const uint32_t hi = (uint32_t)(kx >> 32);
#pragma unroll
            for (int p = 0; p < HIST; p++) atomicAdd(&s_hist[p * 256 + ((hi >> (eh.shift[p] - 32)) & eh.mask[p])], 1u);
        } else if (HIST) {
#pragma unroll
            for (int p = 0; p < 4; p++)
                if (p < eh.passes) atomicAdd(&s_hist[p * 256 + ((uint32_t)(kx >> eh.shift[p]) & eh.mask[p])], 1u);
        }
        if (OWNERS) {
            const int o = owner_of(kx, nranks);
            if (nranks <= 8) {
                const uint64_t one = 1ull << (16 * (o & 3));
                if (o < 4) alo += one; else ahi += one;
            } else {
                atomicAdd(&s_own[o], 1u);
            }
        }
    }
    if (HIST) {
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; p++) {
            if (p < eh.passes) {
                const uint32_t v = s_hist[p * 256 + tid];
                if (v) atomicAdd(eh.hist + p * 256 + tid, v);
            }
        }
    }
    if (OWNERS) {
        if (nranks <= 8) {
            alo = warp_sum(alo);
            ahi = warp_sum(ahi);
            if (lane_id() == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t a = (uint32_t)(alo >> (16 * q)) & 0xffffu, b = (uint32_t)(ahi >> (16 * q)) & 0xffffu;
                    if (a) atomicAdd(&s_own[q], a);
                    if (b) atomicAdd(&s_own[4 + q], b);
                }
            }
        }
        __syncthreads();
        if ((int)tid < nranks && s_own[tid]) atomicAdd(owner_counts + tid, (unsigned long long)s_own[tid]);
    }
}

void extract_canonical(Ctx* c, int k, const uint8_t* codes, size_t n, uint64_t* out, unsigned long long* d_count, int nranks,
                       unsigned long long* d_owner_counts, const SortPre* pre) {
    if (n == 0) return;
    const unsigned tiles = (unsigned)div_up(n, EXTRACT_TILE);
    ExHist eh;
    memset(&eh, 0, sizeof eh);
    if (nranks > 1 && nranks <= 64 && d_owner_counts) {
        extract_kernel<true, 0><<<tiles, EX_THREADS, 0, c->stream>>>(k, codes, out, d_count, nranks, d_owner_counts, eh);
    } else if (pre && pre->d_hist && pre->passes >= 1 && pre->passes <= 4) {
        eh.passes = pre->passes;
        for (int p = 0; p < pre->passes; p++) {
            if (pre->bits[p] > 8) ZB_FAIL(ZB_E_ARG, "extract: digits wider than 8 bits");
            eh.shift[p] = pre->shift[p];
            eh.mask[p] = (1u << pre->bits[p]) - 1u;
        }
        eh.hist = pre->d_hist;
        bool upper = true;
        for (int p = 0; p < pre->passes; p++) upper = upper && pre->shift[p] >= 32;
        if (upper && pre->passes == 2) extract_kernel<false, 2><<<tiles, EX_THREADS, 0, c->stream>>>(k, codes, out, d_count, 0, nullptr, eh);
        else if (upper && pre->passes == 3) extract_kernel<false, 3><<<tiles, EX_THREADS, 0, c->stream>>>(k, codes, out, d_count, 0, nullptr, eh);
        else extract_kernel<false, 4><<<tiles, EX_THREADS, 0, c->stream>>>(k, codes, out, d_count, 0, nullptr, eh);
    } else {
        extract_kernel<false, 0><<<tiles, EX_THREADS, 0, c->stream>>>(k, codes, out, d_count, 0, nullptr, eh);
    }
    ZB_LAUNCH_CHECK(c);
}

// ---------------------------------------------------------------------------------------------
// capture mode (`zot kmerize -C BAITS`, kmerize.py:478-483 and :507-517): `if found: buf.addList(xs)` -- a record
// contributes ALL its k-mers iff one of them is a bait k-mer.  The bait set holds both strands of every bait k-mer, so
// "x in B or rc(x) in B" is "canonical(x) in B".  Records are the stretches between codes 5 (parse_* with
// mark_records); a record that is not captured is overwritten with 4s, and the ordinary extractor then runs on
// what is left.  A rarely used mode: plain kernels, one thread per base.
// ---------------------------------------------------------------------------------------------
static constexpr int CP_THREADS = 256;
static constexpr int CP_PER = 16;
static constexpr int CP_TILE = CP_THREADS * CP_PER;

__global__ void __launch_bounds__(CP_THREADS)
capture_hit_kernel(int k, const uint8_t* __restrict__ codes, uint64_t n, const uint64_t* __restrict__ baits, uint64_t nbaits,
                   uint8_t* __restrict__ hit) {
    const uint64_t p = (uint64_t)blockIdx.x * CP_THREADS + threadIdx.x;
    if (p >= n) return;
    bool ok = true;
    uint64_t fwd = 0, rc = 0;
    for (int i = 0; i < k; i++) {   // the stream is padded with 4s behind n, so p + i never leaves it
        const uint32_t cd = codes[p + i];
        ok &= cd < 4u;
        fwd = (fwd << 2) | (cd & 3u);
        rc |= (uint64_t)(3u - (cd & 3u)) << (2 * i);
    }
    uint8_t h = 0;
    if (ok) {
        const uint64_t x = min(fwd, rc);
        uint64_t lo = 0, hi = nbaits;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (baits[mid] < x) lo = mid + 1; else hi = mid;
        }
        h = (lo < nbaits && baits[lo] == x) ? 1 : 0;
    }
    hit[p] = h;
}

__global__ void __launch_bounds__(CP_THREADS)
capture_count_kernel(const uint8_t* __restrict__ codes, uint64_t n, uint32_t* __restrict__ tile_seps) {
    __shared__ uint32_t sm[CP_THREADS / 32 + 1];
    const uint64_t p0 = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)threadIdx.x * CP_PER;
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < CP_PER; j++) cnt += (p0 + j < n && codes[p0 + j] == 5) ? 1u : 0u;
    uint32_t tot;
    block_excl_scan<CP_THREADS, uint32_t>(cnt, sm, &tot);
    if (threadIdx.x == 0) tile_seps[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
capture_scan_kernel(const uint32_t* __restrict__ tile_seps, uint32_t tiles, uint64_t* __restrict__ tile_base) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < tiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = (i < tiles) ? tile_seps[i] : 0;
        uint64_t tot;
        const uint64_t ex = block_excl_scan<1024, uint64_t>(v, sm, &tot);
        if (i < tiles) tile_base[i] = carry + ex;
        carry += tot;
    }
}

// BLANK = false: found[record] = 1 for every record with a hit; BLANK = true: records without one become 4s.
// record of position p = number of separators before p.
template <bool BLANK>
__global__ void __launch_bounds__(CP_THREADS)
capture_apply_kernel(uint8_t* __restrict__ codes, uint64_t n, const uint64_t* __restrict__ tile_base,
                     const uint8_t* __restrict__ hit, uint8_t* __restrict__ found) {
    __shared__ uint32_t sm[CP_THREADS / 32 + 1];
    const uint64_t p0 = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)threadIdx.x * CP_PER;
    uint8_t cd[CP_PER];
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < CP_PER; j++) {
        cd[j] = (p0 + j < n) ? codes[p0 + j] : (uint8_t)4;
        cnt += (cd[j] == 5) ? 1u : 0u;
    }
    uint32_t tot;
    uint64_t rec = tile_base[blockIdx.x] + block_excl_scan<CP_THREADS, uint32_t>(cnt, sm, &tot);
#pragma unroll
    for (int j = 0; j < CP_PER; j++) {
        if (p0 + j < n) {
            if (BLANK) {
                if (cd[j] != 5 && !found[rec]) codes[p0 + j] = 4;
            } else if (hit[p0 + j]) {
                found[rec] = 1;
            }
        }
        rec += (cd[j] == 5) ? 1u : 0u;
    }
}

void capture_records(Ctx* c, int k, uint8_t* codes, size_t n, const uint64_t* baits, size_t nbaits) {
    if (n == 0) return;
    const uint32_t tiles = (uint32_t)div_up(n, CP_TILE);
    DBuf<uint8_t> hit(c, n);
    DBuf<uint32_t> tile_seps(c, tiles);
    DBuf<uint64_t> tile_base(c, tiles);
    DBuf<uint8_t> found(c, n + 2);   // at most one record per code
    ZB_CUDA(dev_memset(c, found.get(), 0, n + 2));
    Stage st(c, "capture");
    capture_hit_kernel<<<(unsigned)div_up(n, CP_THREADS), CP_THREADS, 0, c->stream>>>(k, codes, n, baits, nbaits, hit.get());
    ZB_LAUNCH_CHECK(c);
    capture_count_kernel<<<tiles, CP_THREADS, 0, c->stream>>>(codes, n, tile_seps.get());
    ZB_LAUNCH_CHECK(c);
    capture_scan_kernel<<<1, 1024, 0, c->stream>>>(tile_seps.get(), tiles, tile_base.get());
    ZB_LAUNCH_CHECK(c);
    capture_apply_kernel<false><<<tiles, CP_THREADS, 0, c->stream>>>(codes, n, tile_base.get(), hit.get(), found.get());
    ZB_LAUNCH_CHECK(c);
    capture_apply_kernel<true><<<tiles, CP_THREADS, 0, c->stream>>>(codes, n, tile_base.get(), hit.get(), found.get());
    ZB_LAUNCH_CHECK(c);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU routing: owner(key) = floor(mix64(key) * nranks / 2^64) -- the high bits of an
// invertible 64-bit mix, so ownership is uniform even for low-complexity sequence.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
owner_count_kernel(const uint64_t* __restrict__ keys, uint64_t n, int nranks, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int sc[64];
    if (threadIdx.x < 64) sc[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&sc[owner_of(keys[i], nranks)], 1u);
    __syncthreads();
    if (threadIdx.x < nranks && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
owner_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t n, int nranks, unsigned long long* __restrict__ cursor,
                      uint64_t* __restrict__ out) {
    // per-CTA: count per owner, reserve ranges with one atomic per owner, then place.
    __shared__ unsigned int sc[64];
    __shared__ unsigned long long sb[64];
    constexpr int PER = 8;
    const uint64_t base = (uint64_t)blockIdx.x * blockDim.x * PER;
    if (threadIdx.x < 64) sc[threadIdx.x] = 0;
    __syncthreads();
    uint64_t kx[PER];
    int ow[PER];
    unsigned int rk[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const uint64_t i = base + (uint64_t)j * blockDim.x + threadIdx.x;
        ow[j] = -1;
        if (i < n) {
            kx[j] = keys[i];
            ow[j] = owner_of(kx[j], nranks);
            rk[j] = atomicAdd(&sc[ow[j]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < nranks) sb[threadIdx.x] = sc[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)sc[threadIdx.x]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; j++)
        if (ow[j] >= 0) out[sb[ow[j]] + rk[j]] = kx[j];
}

// Fused routing + transfer: every key goes straight to dst.p[owner] -- the owner's receive buffer, mapped into
// this process through CUDA IPC, so for a remote owner the stores travel over NVLink as they are issued.  A CTA
// groups its 2048 keys by owner in shared memory first, so every (CTA, owner) run leaves as consecutive,
// fully used 128-byte lines.  `cursor[r]` counts what this rank has written to owner r so far.
static constexpr int RT_THREADS = 256;
int g_route_per = 8;   // keys per thread of a routing tile (ZB_ROUTE_PER: 8, or 16 = runs twice as long, half as many reservations)

// RESERVE: nobody has told this rank where its runs go -- a CTA reserves its run in the owner's buffer with ONE
// system-scope atomicAdd on the owner's cursor word (rcur.p[o], in the owner's memory: over NVLink for a remote owner),
// so the ranks need neither the owner-count pass over the keys nor the all-gather of the count matrix before they can
// route.  The order of the runs in a receive buffer then depends on timing; the owner sorts the buffer anyway.  A
// reservation past `cap` keys raises *err and its keys are not stored.  `cursor` (local) tallies what went to whom.
template <bool RESERVE, int RT_PER>
__global__ void __launch_bounds__(RT_THREADS)
route_p2p_kernel(const uint64_t* __restrict__ keys, uint64_t n, int nranks, PeerPtrs dst,
                 unsigned long long* __restrict__ cursor, PeerPtrs rcur, unsigned long long cap,
                 unsigned int* __restrict__ err) {
    constexpr int RT_TILE = RT_THREADS * RT_PER;
    __shared__ uint32_t sc[64];                 // keys per owner in this tile, then exclusive start
    __shared__ unsigned long long sb[64];       // my run's position in the owner's buffer
    __shared__ uint64_t sk[RT_TILE];
    __shared__ uint8_t so[RT_TILE];
    const unsigned tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * RT_TILE;
    if (tid < 64) sc[tid] = 0;
    __syncthreads();
    uint64_t kx[RT_PER];
    int ow[RT_PER];
    uint32_t rk[RT_PER];
#pragma unroll
    for (int j = 0; j < RT_PER; j++) {
        const uint64_t i = base + (uint64_t)j * RT_THREADS + tid;
        ow[j] = -1;
        if (i < n) {
            kx[j] = keys[i];
            ow[j] = owner_of(kx[j], nranks);
            rk[j] = atomicAdd(&sc[ow[j]], 1u);
        }
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the (<= 64) per-owner counts, one reservation per owner
        const uint32_t c0 = sc[tid], c1 = sc[tid + 32];
        const uint32_t i0 = warp_incl_scan(c0);
        const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31);
        const uint32_t i1 = warp_incl_scan(c1);
        if (RESERVE) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int o = (int)tid + 32 * half;
                const uint32_t cn = half ? c1 : c0;
                if (o < nranks && cn) {
                    unsigned long long at = atomicAdd_system(reinterpret_cast<unsigned long long*>(rcur.p[o]), (unsigned long long)cn);
                    atomicAdd(&cursor[o], (unsigned long long)cn);
                    if (at + cn > cap) {
                        atomicExch(err, 1u);
                        at = ~0ull;
                    }
                    sb[o] = at;
                }
            }
        } else {
            if ((int)tid < nranks && c0) sb[tid] = atomicAdd(&cursor[tid], (unsigned long long)c0);
            if ((int)tid + 32 < nranks && c1) sb[tid + 32] = atomicAdd(&cursor[tid + 32], (unsigned long long)c1);
        }
        sc[tid] = i0 - c0;
        sc[tid + 32] = t0 + i1 - c1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RT_PER; j++) {
        if (ow[j] >= 0) {
            const uint32_t p = sc[ow[j]] + rk[j];
            sk[p] = kx[j];
            so[p] = (uint8_t)ow[j];
        }
    }
    __syncthreads();
    const uint32_t cnt = (uint32_t)min((uint64_t)RT_TILE, n - base);
    for (uint32_t p = tid; p < cnt; p += RT_THREADS) {
        const int o = so[p];
        if (RESERVE && sb[o] == ~0ull) continue;
        dst.p[o][sb[o] + (p - sc[o])] = sk[p];
    }
}

void route_p2p(Ctx* c, const uint64_t* keys, size_t n, int nranks, const PeerPtrs& dst, unsigned long long* d_cursor,
               const PeerPtrs* rcur, unsigned long long cap, unsigned int* d_err, cudaStream_t stream) {
    if (n == 0) return;
    if (nranks > 64) ZB_FAIL(ZB_E_ARG, "route_p2p: nranks > 64");
    if (!stream) stream = c->stream;
    if (rcur && g_route_per == 16)
        route_p2p_kernel<true, 16><<<(unsigned)div_up(n, (size_t)RT_THREADS * 16), RT_THREADS, 0, stream>>>(keys, n, nranks, dst, d_cursor, *rcur, cap, d_err);
    else if (rcur)
        route_p2p_kernel<true, 8><<<(unsigned)div_up(n, (size_t)RT_THREADS * 8), RT_THREADS, 0, stream>>>(keys, n, nranks, dst, d_cursor, *rcur, cap, d_err);
    else if (g_route_per == 16)
        route_p2p_kernel<false, 16><<<(unsigned)div_up(n, (size_t)RT_THREADS * 16), RT_THREADS, 0, stream>>>(keys, n, nranks, dst, d_cursor, dst, 0ull, nullptr);
    else
        route_p2p_kernel<false, 8><<<(unsigned)div_up(n, (size_t)RT_THREADS * 8), RT_THREADS, 0, stream>>>(keys, n, nranks, dst, d_cursor, dst, 0ull, nullptr);
    ZB_LAUNCH_CHECK(c);
}

void bucket_count(Ctx* c, const uint64_t* keys, size_t n, int nranks, unsigned long long* d_counts) {
    if (n == 0) return;
    if (nranks > 64) ZB_FAIL(ZB_E_ARG, "bucket_count: nranks > 64");
    int blocks = (int)std::min<size_t>((size_t)c->sm_count * 8, div_up(n, 256));
    owner_count_kernel<<<blocks, 256, 0, c->stream>>>(keys, n, nranks, d_counts);
    ZB_LAUNCH_CHECK(c);
}

void bucket_scatter(Ctx* c, const uint64_t* keys, size_t n, int nranks, unsigned long long* d_cursor, uint64_t* out) {
    if (n == 0) return;
    owner_scatter_kernel<<<(unsigned)div_up(n, 256 * 8), 256, 0, c->stream>>>(keys, n, nranks, d_cursor, out);
    ZB_LAUNCH_CHECK(c);
}

}  // namespace zb
