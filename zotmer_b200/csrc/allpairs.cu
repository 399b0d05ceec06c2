// allpairs.cu -- |X n Y| for ALL pairs of a collection of sorted k-mer sets: the work behind
// `zot dist` (zotmer/commands/dist.py:145-168 -> library/dist.py:241-265 split()) and `zot jaccard -a`
// (commands/jaccard.py:148-166 -> :31-54), where the reference walks two sorted arrays once per pair (and once
// more per measure).
//
// A pair-at-a-time merge reads 8 (|X| + |Y|) bytes of HBM per pair.  Here the key space is cut into NB buckets
// by the top bits of the key (bucket b of every set holds the same key range, so intersections never cross
// buckets) and the sets into blocks of 8; a CTA takes one pair of blocks (64 set pairs, 28 on the diagonal) and
// a run of 16 buckets.  Per bucket the 8 slices (~0.3 K keys each) of one block are hashed into shared memory
// and the 8 slices of the other block are streamed past them.  A slice is fetched once per 8 partner sets, and
// CTAs are ordered bucket-major, so at any time the resident CTAs touch the same few buckets of all sets (a few
// MB: L2 hits).  HBM sees every set about once; the kernel is bound by shared-memory probe throughput, not by
// HBM -- which is why its set-pairs/s can exceed the naive 8 (|X| + |Y|) B/pair HBM roofline (SURVEY.md 8d
// allows that and asks to say so).
//
// Skewed key spaces: if a slice does not fit its table the kernel raises a flag; the host retries with 16x more
// buckets and finally falls back to the pair-at-a-time kernel (setops.cu pairs_abc).
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace zb {

static constexpr int AP_S = 8;              // sets per block
static constexpr int AP_THREADS = 256;
static constexpr int AP_WARPS = AP_THREADS / 32;
static constexpr int AP_SLOTS = 1024;       // hash slots per staged slice (8 tables = 64 KB -> 3 CTAs per SM)
static constexpr int AP_CAP = 704;          // longest slice a table takes (load <= 0.69)
static constexpr int AP_FILT = 4096;        // filter bits per slice
static constexpr int AP_BPC = 16;           // buckets per CTA (one atomic per pair per CTA)
static_assert(AP_WARPS == AP_S, "one warp per set of a block");
#define AP_EMPTY 0xffffffffffffffffull

// boff[s][b] = first index of set s whose key >> shift is >= b   (b = 0 .. NB)
__global__ void bucket_offsets_kernel(const SetRef* __restrict__ sets, int nsets, int shift, uint32_t NB,
                                      uint32_t* __restrict__ boff) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nsets * (NB + 1)) return;
    const int s = (int)(idx / (NB + 1));
    const uint32_t b = (uint32_t)(idx % (NB + 1));
    const SetRef X = sets[s];
    uint64_t lo = 0, hi = X.n;
    if (b < NB) {
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((X.k[mid] >> shift) < (uint64_t)b) lo = mid + 1; else hi = mid;
        }
    } else {
        lo = X.n;
    }
    boff[idx] = (uint32_t)lo;
}

// upper-triangular (diagonal included) block pair number t -> (bi, bj), bi <= bj < nblk
__device__ __forceinline__ void tile_to_blocks(uint32_t t, uint32_t nblk, uint32_t& bi, uint32_t& bj) {
    // row bi starts at bi * nblk - bi (bi - 1) / 2
    double x = (2.0 * nblk + 1.0 - sqrt((2.0 * nblk + 1.0) * (2.0 * nblk + 1.0) - 8.0 * (double)t)) * 0.5;
    uint32_t r = (uint32_t)x;
    if (r >= nblk) r = nblk - 1;
    while (r > 0 && (uint64_t)r * nblk - (uint64_t)r * (r - 1) / 2 > t) r--;
    while ((uint64_t)(r + 1) * nblk - (uint64_t)(r + 1) * r / 2 <= t) r++;
    bi = r;
    bj = r + (t - (uint32_t)((uint64_t)r * nblk - (uint64_t)r * (r - 1) / 2));
}

// One CTA = one pair of blocks (bi, bj) x AP_BPC consecutive buckets.  Per bucket: warp a hashes slice a of block
// bi into shared-memory table a; then warp w streams slice w of block bj from L2 and probes every table:
// hits[a] counts |set (bi, a) n set (bj, w)| within the bucket.  Probes are independent of each other, so the
// shared-memory latency is hidden by instruction-level parallelism (a merge is one long dependent chain).
__global__ void __launch_bounds__(AP_THREADS, 3)
allpairs_kernel(const SetRef* __restrict__ sets, int nsets, const uint32_t* __restrict__ boff, uint32_t NB, uint32_t nblk,
                uint32_t tile_begin, uint32_t ntiles, unsigned long long* __restrict__ abc, unsigned int* __restrict__ overflow) {
    extern __shared__ __align__(16) uint64_t tab[];   // [AP_S][AP_SLOTS]
    __shared__ uint32_t s_lo[2 * AP_S][AP_BPC + 1];
    // membership filter in front of the tables: bit (f, a) is set iff slice a holds a key whose filter hash is f.
    // One 32-byte row = the same 32 filter positions of all 8 slices, so a key tests all 8 slices with two 16-byte
    // loads; only slices whose bit is set (real members + ~7 % false positives) are probed in their table.
    __shared__ __align__(16) uint32_t s_filt[AP_FILT / 32][AP_S];
    __shared__ unsigned int s_ones;                    // bit a: slice a holds the key 2^64-1 (= the empty marker)
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = tile_begin + blockIdx.x % ntiles;
    const uint32_t b0 = (blockIdx.x / ntiles) * AP_BPC;
    uint32_t bi, bj;
    tile_to_blocks(tile, nblk, bi, bj);
    const bool diag = (bi == bj);
    const int steps = (int)min((uint32_t)AP_BPC, NB - b0);

    for (int idx = tid; idx < 2 * AP_S * (AP_BPC + 1); idx += AP_THREADS) {
        const int sl = idx / (AP_BPC + 1), q = idx % (AP_BPC + 1);
        const int si = (sl < AP_S) ? (int)bi * AP_S + sl : (int)bj * AP_S + sl - AP_S;
        s_lo[sl][q] = (si < nsets) ? boff[(size_t)si * (NB + 1) + min(b0 + q, NB)] : 0u;
    }
    const int ja = (int)bi * AP_S + (int)warp;   // the set whose slices I insert
    const int jb = (int)bj * AP_S + (int)warp;   // the set whose slices I stream
    const uint64_t* Ak = (ja < nsets) ? sets[ja].k : nullptr;
    const uint64_t* Bk = (jb < nsets) ? sets[jb].k : nullptr;
    // tables I have to probe: sets of block bi that exist and (on the diagonal) come before mine
    const int na = min((int)AP_S, max(0, nsets - (int)bi * AP_S));
    const int nprobe = (jb < nsets) ? (diag ? min(na, (int)warp) : na) : 0;
    uint32_t hits[AP_S];
#pragma unroll
    for (int a = 0; a < AP_S; a++) hits[a] = 0;
    uint64_t* mytab = tab + warp * AP_SLOTS;

    for (int q = 0; q < steps; q++) {
        __syncthreads();   // offsets loaded (first step) / every probe of the previous step is done
        {
            const uint4 e4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
#pragma unroll
            for (int r = 0; r < AP_S * AP_SLOTS / 2 / AP_THREADS; r++) reinterpret_cast<uint4*>(tab)[r * AP_THREADS + tid] = e4;
            if (tid == 0) s_ones = 0;
            for (int r = tid; r < AP_FILT / 32 * AP_S / 4; r += AP_THREADS) reinterpret_cast<uint4*>(&s_filt[0][0])[r] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        {
            const uint32_t lo = s_lo[warp][q], len = s_lo[warp][q + 1] - lo;
            if (len > AP_CAP) {
                if (lane == 0) atomicExch(overflow, 1u);
            } else {
                for (uint32_t i = lane; i < len; i += 32) {
                    const uint64_t key = __ldg(Ak + lo + i);
                    if (key == AP_EMPTY) { atomicOr(&s_ones, 1u << warp); continue; }
                    const uint64_t hh = key * 0x9E3779B97F4A7C15ull;
                    uint32_t h = (uint32_t)(hh >> 54);
                    const uint32_t f = (uint32_t)(hh >> 20) & (AP_FILT - 1);   // filter hash: other bits than the slot
                    atomicOr(&s_filt[f >> 5][warp], 1u << (f & 31));
                    while (atomicCAS(reinterpret_cast<unsigned long long*>(&mytab[h]), AP_EMPTY, key) != AP_EMPTY)
                        h = (h + 1) & (AP_SLOTS - 1);
                }
            }
        }
        __syncthreads();
        if (nprobe) {
            const uint32_t lo = s_lo[(diag ? 0 : AP_S) + warp][q], len = s_lo[(diag ? 0 : AP_S) + warp][q + 1] - lo;
            const unsigned ones = s_ones;
            // two keys per round, both loads in flight before the first probe
            for (uint32_t i = lane; i < len; i += 64) {
                const bool two = i + 32 < len;
                uint64_t kk[2];
                kk[0] = __ldg(Bk + lo + i);
                kk[1] = two ? __ldg(Bk + lo + i + 32) : AP_EMPTY;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint64_t key = kk[u];
                    if (key == AP_EMPTY) {   // the empty marker itself (k = 32, all T) or the missing second key
                        if (u == 0 || two) {
#pragma unroll
                            for (int a = 0; a < AP_S; a++) hits[a] += (a < nprobe) ? ((ones >> a) & 1u) : 0u;
                        }
                        continue;
                    }
                    const uint64_t hh = key * 0x9E3779B97F4A7C15ull;
                    const uint32_t h0 = (uint32_t)(hh >> 54);
                    const uint32_t f = (uint32_t)(hh >> 20) & (AP_FILT - 1);
                    const uint4 f0 = *reinterpret_cast<const uint4*>(&s_filt[f >> 5][0]);
                    const uint4 f1 = *reinterpret_cast<const uint4*>(&s_filt[f >> 5][4]);
                    const uint32_t fw[AP_S] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                    uint32_t cand = 0;   // slices whose filter bit is set
#pragma unroll
                    for (int a = 0; a < AP_S; a++) cand |= ((fw[a] >> (f & 31)) & 1u) << a;
                    cand &= (1u << nprobe) - 1u;
#pragma unroll
                    for (int a = 0; a < AP_S; a++) {
                        if ((cand >> a) & 1u) {
                            const uint64_t* t = tab + a * AP_SLOTS;
                            uint32_t h = h0;
                            while (true) {
                                const uint64_t v = t[h];
                                if (v == key) { hits[a]++; break; }
                                if (v == AP_EMPTY) break;
                                h = (h + 1) & (AP_SLOTS - 1);
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < AP_S; a++) {
        const uint32_t h = warp_sum(hits[a]);
        if (lane == 0 && h && a < nprobe) {
            const uint64_t ia = (uint64_t)bi * AP_S + a;
            const uint64_t p = ia * (2ull * nsets - ia - 1) / 2 + ((uint64_t)jb - ia - 1);
            atomicAdd(&abc[3 * p], (unsigned long long)h);
        }
    }
}

// returns false when a bucket overflowed the staging budget (result unusable)
static bool allpairs_try(Ctx* c, const SetRef* d_sets, int nsets, uint32_t NB, int shift, uint32_t nblk, uint32_t tile_begin,
                         uint32_t ntiles, uint64_t* d_abc) {
    DBuf<uint32_t> boff(c, (size_t)nsets * (NB + 1) + 1);
    unsigned int* d_ovf = reinterpret_cast<unsigned int*>(boff.get() + (size_t)nsets * (NB + 1));
    ZB_CUDA(dev_memset(c, d_ovf, 0, 4));
    {
        const uint64_t tot = (uint64_t)nsets * (NB + 1);
        bucket_offsets_kernel<<<(unsigned)div_up(tot, 256), 256, 0, c->stream>>>(d_sets, nsets, shift, NB, boff.get());
        ZB_LAUNCH_CHECK(c);
    }
    const size_t smem = (size_t)AP_S * AP_SLOTS * 8;
    ZB_CUDA(cudaFuncSetAttribute(allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t groups = div_up(NB, AP_BPC);
    // keep one launch below 2^31 CTAs: split the bucket groups
    const uint64_t max_groups = std::max<uint64_t>(1, 0x7fffffffull / ntiles);
    if (groups > max_groups) ZB_FAIL(ZB_E_ARG, "allpairs: %u block pairs x %llu bucket groups exceed one launch; shard the tiles",
                                     ntiles, (unsigned long long)groups);
    Stage st(c, "allpairs");
    allpairs_kernel<<<(unsigned)(groups * ntiles), AP_THREADS, smem, c->stream>>>(
        d_sets, nsets, boff.get(), NB, nblk, tile_begin, ntiles, reinterpret_cast<unsigned long long*>(d_abc), d_ovf);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, d_ovf, 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return reinterpret_cast<uint32_t*>(c->h_scalars)[0] == 0;
}

uint64_t allpairs_tiles(int nsets) {
    const uint64_t nblk = div_up((size_t)nsets, AP_S);
    return nblk * (nblk + 1) / 2;
}

void allpairs_abc(Ctx* c, const std::vector<SetRef>& refs, uint64_t tile_begin, uint64_t tile_end, uint64_t* abc_host) {
    const int nsets = (int)refs.size();
    const uint64_t npairs = (uint64_t)nsets * (nsets - 1) / 2;
    if (npairs == 0) return;
    const uint32_t nblk = (uint32_t)div_up((size_t)nsets, AP_S);
    const uint64_t all_tiles = allpairs_tiles(nsets);
    if (tile_end == 0 || tile_end > all_tiles) tile_end = all_tiles;
    memset(abc_host, 0, npairs * 3 * 8);
    if (tile_begin >= tile_end) return;
    const uint32_t ntiles = (uint32_t)(tile_end - tile_begin);

    DBuf<SetRef> d_refs(c, nsets);
    ZB_CUDA(cudaMemcpyAsync(d_refs.get(), refs.data(), nsets * sizeof(SetRef), cudaMemcpyHostToDevice, c->stream));
    // key range: the sets are sorted, so the largest key is the largest last element
    uint64_t maxkey = 0, max_n = 0;
    {
        std::vector<uint64_t> last(nsets, 0);
        for (int i = 0; i < nsets; i++) {
            max_n = std::max<uint64_t>(max_n, refs[i].n);
            if (refs[i].n) ZB_CUDA(cudaMemcpyAsync(&last[i], refs[i].k + refs[i].n - 1, 8, cudaMemcpyDeviceToHost, c->stream));
        }
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < nsets; i++) maxkey = std::max(maxkey, last[i]);
    }
    const int key_bits = maxkey ? 64 - __builtin_clzll(maxkey) : 1;
    DBuf<uint64_t> d_abc(c, npairs * 3);
    bool ok = false;
    int lgNB = 4;
    while ((max_n >> lgNB) > 384 && lgNB < 22) lgNB++;   // <= 384 keys of the largest set per bucket on average
    for (int attempt = 0; attempt < 2 && !ok; attempt++, lgNB += 4) {
        if (lgNB > key_bits) lgNB = key_bits;
        if (lgNB > 24) break;
        const uint32_t NB = 1u << lgNB;
        if (div_up(NB, AP_BPC) * (uint64_t)ntiles > 0x7fffffffull) break;
        ZB_CUDA(dev_memset(c, d_abc.get(), 0, npairs * 3 * 8));
        ok = allpairs_try(c, d_refs.get(), nsets, NB, key_bits - lgNB, nblk, (uint32_t)tile_begin, ntiles, d_abc.get());
        if (lgNB == key_bits) break;
    }
    if (ok) {
        ZB_CUDA(cudaMemcpyAsync(abc_host, d_abc.get(), npairs * 3 * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
    }
    // pairs of the requested tiles: fill |X \ Y|, |Y \ X| (and, after an overflow, everything pair by pair)
    std::vector<uint32_t> I, J;
    for (uint64_t t = tile_begin; t < tile_end; t++) {
        // same enumeration as tile_to_blocks
        uint64_t r = 0, start = 0;
        while (start + (nblk - r) <= t) { start += nblk - r; r++; }
        const uint64_t bi = r, bj = r + (t - start);
        for (uint64_t i = bi * AP_S; i < std::min<uint64_t>((bi + 1) * AP_S, nsets); i++)
            for (uint64_t j = std::max(bj * AP_S, i + 1); j < std::min<uint64_t>((bj + 1) * AP_S, nsets); j++) {
                I.push_back((uint32_t)i);
                J.push_back((uint32_t)j);
            }
    }
    if (!ok) {
        // skewed key space: pair-at-a-time merge path (setops.cu)
        std::vector<uint64_t> tmp(I.size() * 3);
        pairs_abc_host(c, refs, I.data(), J.data(), I.size(), tmp.data());
        for (size_t q = 0; q < I.size(); q++) {
            const uint64_t i = I[q], j = J[q];
            const uint64_t p = i * (2ull * nsets - i - 1) / 2 + (j - i - 1);
            abc_host[3 * p] = tmp[3 * q];
        }
    }
    for (size_t q = 0; q < I.size(); q++) {
        const uint64_t i = I[q], j = J[q];
        const uint64_t p = i * (2ull * nsets - i - 1) / 2 + (j - i - 1);
        const uint64_t a = abc_host[3 * p];
        abc_host[3 * p + 1] = refs[i].n - a;
        abc_host[3 * p + 2] = refs[j].n - a;
    }
}

}  // namespace zb
