// nwaymerge.cu -- N-way union of sorted counted k-mer sets with the counts summed, in ONE pass over the inputs:
// the kernel behind `zot merge` (zotmer/commands/merge.py:26-86 merge() = pairwise merge of two (x, c) streams,
// :94-163 _kmerRadixBlockStream + mergeNinto = N-way union over key-range blocks with a heap and a dict).
//
// The reference cuts the key space into 4096 radix blocks and merges the inputs block by block; so does this,
// with blocks small enough for shared memory:
//
//   1. the key space is cut into NB = 2^cb buckets by the top cb bits of the key, cb chosen so that ALL inputs
//      together hold <= 4096 entries per bucket (~2000 on average).  bm_offsets_kernel finds where every bucket
//      starts in every input (one streaming pass over the keys); bm_starts_kernel adds them up to the
//      bucket's position in the (virtual) concatenation and takes the largest bucket.
//   2. bm_merge_kernel: one CTA per bucket gathers the bucket's slice of every input (contiguous runs) into
//      shared memory, inserts every key into a shared-memory hash table (one CAS; the first entry of a key is its
//      "head", every other entry adds its count to the head's sum), orders the heads by the next 11 key bits with a
//      shared-memory counting sort and by the whole key inside those (tiny) groups, and writes the bucket's
//      distinct (key, count) run.  No CTA waits for another one.
//   3. a one-CTA scan of the per-bucket head counts and a compaction kernel place the runs (as segsort.cu does).
//
// Algorithmic bytes: 12 B per input entry read + 12 B per output entry staged, re-read and written; the offsets
// table is 4 B per (bucket, input).  Fallback (skewed keys whose buckets do not fit, > 1024 inputs): the caller
// uses the weighted sort_count of the concatenation.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace zb {

static constexpr int BM_PER = 8;                        // entries per thread: a bucket holds THREADS x 8 entries of all inputs together
static constexpr int BM_MAXSETS = 1024;
// per shape (THREADS = 512 / 256): hash slots 2 x capacity (load <= 0.5), groups of the in-bucket counting sort = THREADS x 4
int g_merge_cfg = 0;                                    // ZB_MERGE_CFG: 0 = 512 threads (default), 1 = 256 threads (measured: slower)
#define BM_EMPTY 0xffffffffu

struct KCRef {
    const uint64_t* k;
    const uint32_t* c;
    uint64_t n;
};

// off[b * nsets + i] = first index of input i whose key has top bits >= b   (b = 0 .. nb; row nb = sizes).
// One streaming pass over the keys (blockIdx.y = input): element j closes every bucket between its predecessor's and
// its own.  (A binary search per (bucket, input) -- 17 M searches over 5 GB of keys for 64 bacterial sets -- took
// 7.9 ms: nearly every probe misses the TLB.  This reads 8 B per entry, coalesced.)  `off` is zeroed by the caller
// (rows of empty inputs stay zero).
__global__ void __launch_bounds__(256)
bm_offsets_kernel(const KCRef* __restrict__ sets, int nsets, int shift, uint32_t nb, uint32_t* __restrict__ off, uint64_t key_base) {
    bucket_offsets_body(sets, nsets, shift, nb, off, key_base);
}

// start[b] = sum over inputs of off[b][i] (position of bucket b in the virtual concatenation); one warp per bucket
__global__ void __launch_bounds__(256)
bm_starts_kernel(const uint32_t* __restrict__ off, int nsets, uint32_t nb, uint64_t* __restrict__ start) {
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b > nb) return;
    uint64_t s = 0;
    for (int i = lane_id(); i < nsets; i += 32) s += off[(size_t)b * nsets + i];
    s = warp_sum(s);
    if (lane_id() == 0) start[b] = s;
}

__global__ void __launch_bounds__(256)
bm_maxsize_kernel(const uint64_t* __restrict__ start, uint32_t nb, unsigned long long* __restrict__ maxsize) {
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    unsigned long long v = (b < nb) ? (unsigned long long)(start[b + 1] - start[b]) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane_id() == 0 && v) atomicMax(maxsize, v);
}

// One CTA per bucket.  Shared memory: keys 32 KB + hash table 32 KB + count sums 16 KB + slice prefixes 8 KB + group
// sizes 8 KB = 96 KB -> 2 CTAs per SM.
// BY_SLICE: a warp copies whole slices (few inputs: a thread finds the slice of its entry by bisection instead).
// Two shapes (ZB_MERGE_CFG): THREADS = 512, buckets of <= 4096 entries, 2 CTAs per SM (96 KB each) -- the default; THREADS =
// 256, buckets of <= 2048 entries, 4 CTAs per SM (52 KB each).  Four independent CTAs were what ap_bucket_kernel wanted; here
// they lose: 64 bacterial sets 18.6 -> 40.0 ms (twice the buckets: offsets 3.1 -> 8.2 ms, buckets 9.4 -> 25.2 ms -- with 64
// slices per bucket the per-bucket set-up, not the per-entry work, is what doubles), the human-scale compaction 257 -> 282 ms
// (gpurun_out/r2_test14.log and the run after it).
template <bool BY_SLICE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 4)
bm_merge_kernel(const KCRef* __restrict__ sets, int nsets, const uint32_t* __restrict__ off, const uint64_t* __restrict__ start,
                int fine_shift, uint32_t fine_mask, uint64_t key_base, uint64_t* __restrict__ tmp_k, uint32_t* __restrict__ tmp_c,
                uint32_t* __restrict__ tile_heads, unsigned int* __restrict__ err) {
    constexpr int BM_THREADS = THREADS;
    constexpr int BM_CAP = THREADS * BM_PER;
    constexpr int BM_HASH = 2 * BM_CAP;
    constexpr int BM_HASH_BITS = (THREADS == 512) ? 13 : 12;
    constexpr int BM_FINE = THREADS * 4;
    constexpr int BM_SETS_PER = BM_MAXSETS / THREADS;
    extern __shared__ __align__(16) unsigned char bm_raw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(bm_raw);                  // [BM_CAP] gathered keys
    uint32_t* table = reinterpret_cast<uint32_t*>(sk + BM_CAP);          // [BM_HASH] position of a key's head
    uint32_t* wsum = table + BM_HASH;                                    // [BM_CAP] count of an entry; sum of counts at a head
    uint32_t* spre = wsum + BM_CAP;                                      // [BM_MAXSETS + 1] slice starts inside the bucket
    uint32_t* soff = spre + BM_MAXSETS + 1;                              // [BM_MAXSETS] slice starts inside the inputs
    uint32_t* hist = soff + BM_MAXSETS;                                  // [BM_FINE + 1]
    uint64_t* hs = reinterpret_cast<uint64_t*>(table);                   // [BM_CAP] heads grouped by fine digit (after the dedupe)
    uint32_t* hc = reinterpret_cast<uint32_t*>(sk);                      // [BM_CAP] their counts            (after the dedupe)
    __shared__ uint32_t s_scan[BM_THREADS / 32 + 1];

    const unsigned tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint64_t s0 = start[b];
    const int m = (int)(start[b + 1] - s0);
    if (m == 0) {
        if (tid == 0) tile_heads[b] = 0;
        return;
    }
    {
        const uint4 e4 = make_uint4(BM_EMPTY, BM_EMPTY, BM_EMPTY, BM_EMPTY);
#pragma unroll
        for (int j = 0; j < BM_HASH / 4 / BM_THREADS; j++) reinterpret_cast<uint4*>(table)[j * BM_THREADS + tid] = e4;
#pragma unroll
        for (int j = 0; j < BM_FINE / BM_THREADS; j++) hist[j * BM_THREADS + tid] = 0;
    }
    // ---- where every input's slice of this bucket starts (in the input, and inside the bucket)
    {
        uint32_t len[BM_SETS_PER];
        uint32_t tot = 0;
#pragma unroll
        for (int u = 0; u < BM_SETS_PER; u++) {
            const int i = (int)tid * BM_SETS_PER + u;
            uint32_t o0 = 0, o1 = 0;
            if (i < nsets) {
                o0 = __ldg(off + (size_t)b * nsets + i);
                o1 = __ldg(off + (size_t)(b + 1) * nsets + i);
                soff[i] = o0;
            }
            len[u] = o1 - o0;
            tot += len[u];
        }
        uint32_t all;
        uint32_t ex = block_excl_scan<BM_THREADS, uint32_t, false>(tot, s_scan, &all);
#pragma unroll
        for (int u = 0; u < BM_SETS_PER; u++) {
            const int i = (int)tid * BM_SETS_PER + u;
            if (i < nsets) spre[i] = ex;
            ex += len[u];
        }
        if (tid == 0) spre[nsets] = (uint32_t)m;
    }
    __syncthreads();

    // ---- gather the slices: entry q of the bucket = element q - spre[i] of input i's slice
    if (BY_SLICE) {
        const int lane = (int)(tid & 31), warp = (int)(tid >> 5);
        constexpr int NW = BM_THREADS / 32;
        constexpr int U = 4;   // slices a warp copies at a time: 2 U loads in flight per lane
        for (int i0 = warp; i0 < nsets; i0 += U * NW) {
            uint32_t q0[U], len[U];
            const uint64_t* pk[U];
            const uint32_t* pc[U];
            uint32_t mx = 0;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * NW;
                q0[u] = 0; len[u] = 0; pk[u] = nullptr; pc[u] = nullptr;
                if (i < nsets) {
                    q0[u] = spre[i];
                    len[u] = spre[i + 1] - q0[u];
                    pk[u] = sets[i].k + soff[i];
                    pc[u] = sets[i].c + soff[i];
                }
                mx = max(mx, len[u]);
            }
            for (uint32_t e = lane; e < mx; e += 32) {
                uint64_t kv[U];
                uint32_t cv[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (e < len[u]) { kv[u] = __ldg(pk[u] + e); cv[u] = __ldg(pc[u] + e); }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (e < len[u]) { sk[q0[u] + e] = kv[u]; wsum[q0[u] + e] = cv[u]; }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < BM_PER; j++) {
            const int q = j * BM_THREADS + (int)tid;
            if (q < m) {
                int lo = 0, hi = nsets - 1;   // first input whose slice ends beyond q
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (spre[mid + 1] <= (uint32_t)q) lo = mid + 1; else hi = mid;
                }
                const uint64_t src = (uint64_t)soff[lo] + ((uint32_t)q - spre[lo]);
                sk[q] = __ldg(sets[lo].k + src);
                wsum[q] = __ldg(sets[lo].c + src);
            }
        }
    }
    __syncthreads();

    // ---- dedupe: the first entry to claim a key's slot is its head; every other entry adds its count to the head's
    uint64_t kx[BM_PER];
    uint32_t headbits = 0;
#pragma unroll
    for (int j = 0; j < BM_PER; j++) {
        const int q = j * BM_THREADS + (int)tid;
        kx[j] = 0;
        if (q < m) {
            const uint64_t x = sk[q];
            kx[j] = x;
            uint32_t h = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> (64 - BM_HASH_BITS));
            while (true) {
                const uint32_t old = atomicCAS(&table[h], BM_EMPTY, (uint32_t)q);
                if (old == BM_EMPTY) { headbits |= 1u << j; break; }
                if (sk[old] == x) {
                    const uint32_t wt = wsum[q];   // nobody else touches the slot of an entry that is not a head
                    const uint32_t before = atomicAdd(&wsum[old], wt);
                    if (before + wt < before) atomicExch(err, 1u);
                    break;
                }
                h = (h + 1) & (BM_HASH - 1);
            }
        }
    }
    __syncthreads();

    // ---- heads: counting sort by the next key bits (rank inside a group = arrival order, fixed up below)
    uint32_t rd[BM_PER], wx[BM_PER];
#pragma unroll
    for (int j = 0; j < BM_PER; j++) {
        if ((headbits >> j) & 1u) {
            const uint32_t d = (uint32_t)((kx[j] - key_base) >> fine_shift) & fine_mask;
            rd[j] = atomicAdd(&hist[d], 1u) | (d << 16);
            wx[j] = wsum[j * BM_THREADS + tid];
        }
    }
    __syncthreads();
    {   // exclusive scan of the group sizes
        constexpr int GP = BM_FINE / BM_THREADS;
        uint32_t v[GP];
        uint32_t t = 0;
#pragma unroll
        for (int u = 0; u < GP; u++) { v[u] = hist[tid * GP + u]; t += v[u]; }
        uint32_t all;
        uint32_t ex = block_excl_scan<BM_THREADS, uint32_t, false>(t, s_scan, &all);
#pragma unroll
        for (int u = 0; u < GP; u++) { hist[tid * GP + u] = ex; ex += v[u]; }
        if (tid == 0) { hist[BM_FINE] = all; tile_heads[b] = all; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BM_PER; j++) {
        if ((headbits >> j) & 1u) {
            const uint32_t p = hist[rd[j] >> 16] + (rd[j] & 0xffffu);
            hs[p] = kx[j];
            hc[p] = wx[j];
        }
    }
    __syncthreads();

    // ---- every head ranks itself inside its group (one or two keys) and writes (key, count)
    const int H = (int)hist[BM_FINE];
    for (int p = (int)tid; p < H; p += BM_THREADS) {
        const uint64_t x = hs[p];
        const uint32_t d = (uint32_t)((x - key_base) >> fine_shift) & fine_mask;
        const int g0 = (int)hist[d], g1 = (int)hist[d + 1];
        int r = 0;
        const int g = g1 - g0;
        if (g > 1 && g <= 4) {   // the usual case, without a data-dependent loop
#pragma unroll
            for (int u = 0; u < 4; u++) r += (u < g && hs[min(g0 + u, g1 - 1)] < x) ? 1 : 0;
        } else if (g > 4) {
            for (int p2 = g0; p2 < g1; p2++) r += (hs[p2] < x) ? 1 : 0;
        }
        tmp_k[s0 + g0 + r] = x;
        tmp_c[s0 + g0 + r] = hc[p];
    }
}

__global__ void __launch_bounds__(1024) bm_scan_kernel(const uint32_t* __restrict__ tile_heads, uint32_t tiles,
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

// 8 buckets per CTA, one warp each (a bucket's run is ~1000 entries)
__global__ void __launch_bounds__(256)
bm_compact_kernel(const uint64_t* __restrict__ tmp_k, const uint32_t* __restrict__ tmp_c, const uint32_t* __restrict__ tile_heads,
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

bool merge_nway(Ctx* c, const std::vector<const uint64_t*>& ks, const std::vector<const uint32_t*>& cs,
                const std::vector<size_t>& ns, int key_bits, DBuf<uint64_t>* out_k, DBuf<uint32_t>* out_c, size_t* n_out,
                uint64_t key_base) {
    const int nsets = (int)ks.size();
    if (nsets < 1 || nsets > BM_MAXSETS) return false;
    size_t total = 0;
    for (int i = 0; i < nsets; i++) {
        if (ns[i] >= ((size_t)1 << 32)) return false;
        total += ns[i];
    }
    if (total == 0) { *n_out = 0; return true; }
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;

    std::vector<KCRef> refs(nsets);
    for (int i = 0; i < nsets; i++) refs[i] = KCRef{ks[i], cs[i], (uint64_t)ns[i]};
    DBuf<KCRef> d_refs(c, nsets);
    ZB_CUDA(cudaMemcpyAsync(d_refs.get(), refs.data(), nsets * sizeof(KCRef), cudaMemcpyHostToDevice, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // `refs` is pageable host memory

    if (const char* e = getenv("ZB_MERGE_CFG")) g_merge_cfg = atoi(e);
    const int THREADS = (g_merge_cfg == 0) ? 512 : 256;
    const size_t BM_CAP = (size_t)THREADS * BM_PER;
    const int BM_FINE = THREADS * 4, BM_FINE_BITS = (THREADS == 512) ? 11 : 10;
    const size_t BM_HASH = 2 * BM_CAP;
    // buckets: at most 2/3 of the capacity on average; more bits when the largest one does not fit
    int cb = 0;
    while (cb < key_bits && (total >> cb) > (size_t)BM_CAP * 2 / 3) cb++;
    DBuf<uint32_t> off;
    DBuf<uint64_t> start;
    uint32_t nb = 0;
    for (int attempt = 0;; attempt++) {
        if (cb > 30 || ((size_t)1 << cb) * (size_t)nsets > ((size_t)1 << 28)) return false;
        nb = 1u << cb;
        const size_t noff = (size_t)(nb + 1) * nsets;
        off.alloc(c, noff);
        start.alloc(c, (size_t)nb + 4);
        unsigned long long* d_max = reinterpret_cast<unsigned long long*>(start.get() + nb + 2);
        ZB_CUDA(dev_memset(c, d_max, 0, 8));
        {
            Stage st(c, "merge_offsets");
            ZB_CUDA(dev_memset(c, off.get(), 0, noff * 4));
            size_t nmax = 0;
            for (int i = 0; i < nsets; i++) nmax = std::max(nmax, ns[i]);
            const dim3 grid((unsigned)std::min<size_t>(div_up(nmax, 256 * 2 * 8), 65535), (unsigned)nsets);
            bm_offsets_kernel<<<grid, 256, 0, c->stream>>>(d_refs.get(), nsets, key_bits - cb, nb, off.get(), key_base);
            ZB_LAUNCH_CHECK(c);
            bm_starts_kernel<<<(unsigned)div_up((size_t)nb + 1, 8), 256, 0, c->stream>>>(off.get(), nsets, nb, start.get());
            ZB_LAUNCH_CHECK(c);
            bm_maxsize_kernel<<<(unsigned)div_up(nb, 256), 256, 0, c->stream>>>(start.get(), nb, d_max);
            ZB_LAUNCH_CHECK(c);
        }
        ZB_CUDA(read_back(c, d_max, 8));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        const size_t mx = (size_t)c->h_scalars[0];
        if (mx <= (size_t)BM_CAP) break;
        if (attempt >= 3 || cb >= key_bits) return false;   // badly skewed key space: the caller sorts instead
        int more = 1;
        while (((mx >> more) > (size_t)BM_CAP / 2) && more < 8) more++;
        cb = std::min(key_bits, cb + more);
    }
    const int fb = std::min(BM_FINE_BITS, key_bits - cb);
    const int fine_shift = key_bits - cb - fb;
    const uint32_t fine_mask = (1u << fb) - 1u;

    DBuf<uint64_t> tmp_k(c, total);
    DBuf<uint32_t> tmp_c(c, total);
    DBuf<uint32_t> tile_heads(c, nb);
    DBuf<uint64_t> tile_off(c, (size_t)nb + 4);
    uint64_t* totals = tile_off.get() + nb;
    unsigned int* err = reinterpret_cast<unsigned int*>(totals + 1);
    ZB_CUDA(dev_memset(c, totals, 0, 16));
    const size_t smem = (size_t)BM_CAP * 8 + (size_t)BM_HASH * 4 + (size_t)BM_CAP * 4 + (size_t)(2 * BM_MAXSETS + 1) * 4 +
                        (size_t)(BM_FINE + 1) * 4;
    {
        Stage st(c, "merge_buckets");
#define ZB_BM_LAUNCH(SLICE, T)                                                                                                  \
        do {                                                                                                                    \
            ZB_CUDA(cudaFuncSetAttribute(bm_merge_kernel<SLICE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
            bm_merge_kernel<SLICE, T><<<nb, T, smem, c->stream>>>(d_refs.get(), nsets, off.get(), start.get(), fine_shift,      \
                                                                   fine_mask, key_base, tmp_k.get(), tmp_c.get(), tile_heads.get(), err); \
        } while (0)
        if (nsets >= 16) { if (THREADS == 512) ZB_BM_LAUNCH(true, 512); else ZB_BM_LAUNCH(true, 256); }
        else { if (THREADS == 512) ZB_BM_LAUNCH(false, 512); else ZB_BM_LAUNCH(false, 256); }
#undef ZB_BM_LAUNCH
        ZB_LAUNCH_CHECK(c);
        bm_scan_kernel<<<1, 1024, 0, c->stream>>>(tile_heads.get(), nb, tile_off.get(), totals);
        ZB_LAUNCH_CHECK(c);
    }
    ZB_CUDA(read_back(c, totals, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const size_t nd = (size_t)c->h_scalars[0];
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 1)[0] != 0)
        ZB_FAIL(ZB_E_RANGE, "k-mer count exceeds 2^32-1 (reference: array('I') OverflowError, kmerize.py:374)");
    out_k->alloc(c, nd);
    out_c->alloc(c, nd);
    {
        Stage st(c, "merge_compact");
        bm_compact_kernel<<<(unsigned)div_up(nb, 8), 256, 0, c->stream>>>(tmp_k.get(), tmp_c.get(), tile_heads.get(), start.get(),
                                                                         tile_off.get(), nb, out_k->get(), out_c->get());
        ZB_LAUNCH_CHECK(c);
    }
    *n_out = nd;
    return true;
}

// cut[i * m + j] = first index of input i whose key is >= split[j]: one thread per (input, splitter)
__global__ void __launch_bounds__(128)
slab_cuts_kernel(const KCRef* __restrict__ sets, int nsets, const uint64_t* __restrict__ split, int m, uint64_t* __restrict__ cut) {
    const int i = blockIdx.x;
    const uint64_t* __restrict__ k = sets[i].k;
    const uint64_t n = sets[i].n;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const uint64_t x = split[j];
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (__ldg(k + mid) < x) lo = mid + 1; else hi = mid;
        }
        cut[(size_t)i * m + j] = lo;
    }
}

// The same, slab by slab of the key space, so that the bucket merge's staging (12 B per input entry of a slab) stays
// below `slab_entries` entries however large the inputs are: splitters are order statistics of the largest input, every
// input is cut at them (one lower_bound call per input), the slabs are merged one after the other -- bucketed by
// (key - slab base), so that no bucket lies outside its slab -- and laid end to end.
bool merge_nway_slabs(Ctx* c, const std::vector<const uint64_t*>& ks, const std::vector<const uint32_t*>& cs,
                      const std::vector<size_t>& ns, int key_bits, size_t slab_entries, DBuf<uint64_t>* out_k,
                      DBuf<uint32_t>* out_c, size_t* n_out) {
    const size_t nr = ks.size();
    size_t total = 0, big = 0;
    for (size_t i = 0; i < nr; i++) {
        total += ns[i];
        if (ns[i] > ns[big]) big = i;
    }
    const size_t S = std::max<size_t>(1, div_up(total, std::max<size_t>(slab_entries, 1)));
    if (S == 1) return merge_nway(c, ks, cs, ns, key_bits, out_k, out_c, n_out);
    std::vector<uint64_t> split;
    for (size_t j = 1; j < S; j++) {
        const size_t pos = (ns[big] * j) / S;
        ZB_CUDA(read_back(c, ks[big] + pos, 8));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        split.push_back(c->h_scalars[0]);
    }
    // every input is cut at every splitter by ONE kernel (a lower_bound call per input was 64 launches + round trips)
    std::vector<std::vector<uint64_t>> cut(nr, std::vector<uint64_t>(S + 1, 0));
    {
        const int m = (int)(S - 1);
        std::vector<KCRef> refs(nr);
        for (size_t i = 0; i < nr; i++) refs[i] = KCRef{ks[i], cs[i], (uint64_t)ns[i]};
        DBuf<KCRef> d_refs(c, nr);
        DBuf<uint64_t> d_io(c, (size_t)m + nr * (size_t)m);
        std::vector<uint64_t> h_cut(nr * (size_t)m);
        ZB_CUDA(cudaMemcpyAsync(d_refs.get(), refs.data(), nr * sizeof(KCRef), cudaMemcpyHostToDevice, c->stream));
        ZB_CUDA(cudaMemcpyAsync(d_io.get(), split.data(), (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
        slab_cuts_kernel<<<(unsigned)nr, 128, 0, c->stream>>>(d_refs.get(), (int)nr, d_io.get(), m, d_io.get() + m);
        ZB_LAUNCH_CHECK(c);
        ZB_CUDA(cudaMemcpyAsync(h_cut.data(), d_io.get() + m, h_cut.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < nr; i++) {
            for (int j = 0; j < m; j++) cut[i][j + 1] = h_cut[i * m + j];
            cut[i][S] = ns[i];
            for (size_t j = 1; j <= S; j++) cut[i][j] = std::max(cut[i][j], cut[i][j - 1]);
        }
    }
    struct Slab { DBuf<uint64_t> k; DBuf<uint32_t> c; size_t n = 0; };
    std::vector<Slab> slabs(S);
    const uint64_t top = (key_bits >= 64) ? ~0ull : ((1ull << key_bits) - 1ull);
    for (size_t j = 0; j < S; j++) {
        std::vector<const uint64_t*> sk;
        std::vector<const uint32_t*> sc;
        std::vector<size_t> sn;
        for (size_t i = 0; i < nr; i++) {
            const size_t b = cut[i][j], e = cut[i][j + 1];
            if (e > b) { sk.push_back(ks[i] + b); sc.push_back(cs[i] + b); sn.push_back(e - b); }
        }
        if (sk.empty()) continue;
        const uint64_t lo = (j == 0) ? 0ull : split[j - 1];
        const uint64_t hi = (j + 1 < S) ? split[j] : top;
        const uint64_t width = hi - lo;
        const int bits = width ? 64 - __builtin_clzll(width) : 1;
        if (!merge_nway(c, sk, sc, sn, bits, &slabs[j].k, &slabs[j].c, &slabs[j].n, lo)) return false;
    }
    size_t nd = 0;
    for (auto& sl : slabs) nd += sl.n;
    out_k->alloc(c, nd);
    out_c->alloc(c, nd);
    size_t o = 0;
    for (auto& sl : slabs) {
        if (sl.n) {
            ZB_CUDA(dev_copy(c, out_k->get() + o, sl.k.get(), sl.n * 8));
            ZB_CUDA(dev_copy(c, out_c->get() + o, sl.c.get(), sl.n * 4));
        }
        o += sl.n;
        sl.k.release();
        sl.c.release();
    }
    *n_out = nd;
    return true;
}

}  // namespace zb
