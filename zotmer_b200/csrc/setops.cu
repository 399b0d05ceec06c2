// setops.cu -- sorted k-mer-set algebra on the device: run-length count / reduce-by-key, merge-path
// merge, both-strand mirroring, trim (compaction), projection+dedup, count histogram / acgt tallies,
// and batched intersection cardinalities.  All kernels are HBM-streaming (no tensor-core work).
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace zb {

static constexpr int ST_THREADS = 256;
static constexpr int ST_ITEMS = 16;
static constexpr int ST_TILE = ST_THREADS * ST_ITEMS;  // 4096 elements per CTA

// ---------------------------------------------------------------------------------------------
// reduce_by_key: sorted keys (+ optional u32 weights) -> distinct keys, start offsets in the
// (weighted) prefix space.  Striped tile layout: item j of thread t is element j*THREADS + t, so
// every global access is a fully coalesced 2 KB row.  Chained scan carries (#heads, weight sum).
// Reference: zotmer/commands/kmerize.py:41-132 (merge = RLE of the sorted buffer, counts summed).
// ---------------------------------------------------------------------------------------------
// A CTA owns ST_ROUNDS consecutive 4096-element sub-tiles (16384 elements) and pays for ONE chained-scan
// step: phase 1 streams the keys once and keeps only head flags (bit masks) and per-(row,warp) cell
// counts; after the look-back, phase 2 re-reads just the keys at run heads (L2 hits) and writes them.
static constexpr int ST_ROUNDS = 4;
static constexpr int ST_BIG = ST_ROUNDS * ST_TILE;
static constexpr int ST_CELLS = ST_ROUNDS * ST_ITEMS * (ST_THREADS / 32);  // 512

template <bool WEIGHTED>
__global__ void __launch_bounds__(ST_THREADS, 4)
rbk_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ w, uint64_t n,
           uint64_t* __restrict__ out_k, uint64_t* __restrict__ out_start, uint64_t* __restrict__ st_heads,
           uint64_t* __restrict__ st_wsum, uint32_t* __restrict__ ticket, uint64_t* __restrict__ totals) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_head[ST_CELLS];                      // heads per (round,row,warp), then exclusive
    __shared__ uint64_t s_sum[WEIGHTED ? ST_CELLS : 1];        // weight per cell, then exclusive
    __shared__ uint64_t s_pref[2];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * ST_BIG;

    uint64_t headbits = 0;   // bit r*16+j
#pragma unroll 1
    for (int r = 0; r < ST_ROUNDS; r++) {
        unsigned hb = 0;
#pragma unroll
        for (int j = 0; j < ST_ITEMS; j++) {
            const uint64_t i = base + (uint64_t)(r * ST_ITEMS + j) * ST_THREADS + tid;
            const bool in = i < n;
            const uint64_t key = in ? __ldg(keys + i) : 0;
            uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
            if (lane == 0 && in && i > 0) prev = __ldg(keys + i - 1);
            const bool head = in && (i == 0 || prev != key);
            hb |= (head ? 1u : 0u) << j;
            const unsigned b = __ballot_sync(0xffffffffu, head);
            if (WEIGHTED) {
                const uint64_t ws = warp_sum<uint64_t>(in ? __ldg(w + i) : 0u);
                if (lane == 0) s_sum[(r * ST_ITEMS + j) * (ST_THREADS / 32) + warp] = ws;
            }
            if (lane == 0) s_head[(r * ST_ITEMS + j) * (ST_THREADS / 32) + warp] = __popc(b);
        }
        headbits |= (uint64_t)hb << (r * ST_ITEMS);
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PERL = ST_CELLS / 32;  // 16 consecutive cells per lane
        uint32_t hs = 0;
        uint64_t ss = 0;
#pragma unroll
        for (int q = 0; q < PERL; q++) {
            hs += s_head[lane * PERL + q];
            if (WEIGHTED) ss += s_sum[lane * PERL + q];
        }
        const uint32_t hi = warp_incl_scan(hs);
        const uint32_t htot = __shfl_sync(0xffffffffu, hi, 31);
        uint32_t he = hi - hs;
        uint64_t stot = 0, se = 0;
        if (WEIGHTED) {
            const uint64_t si = warp_incl_scan(ss);
            stot = __shfl_sync(0xffffffffu, si, 31);
            se = si - ss;
        }
#pragma unroll
        for (int q = 0; q < PERL; q++) {
            const uint32_t h = s_head[lane * PERL + q];
            s_head[lane * PERL + q] = he;
            he += h;
            if (WEIGHTED) {
                const uint64_t x = s_sum[lane * PERL + q];
                s_sum[lane * PERL + q] = se;
                se += x;
            }
        }
        const uint64_t ph = lookback_u64(st_heads, tile, htot);
        uint64_t pw = 0;
        if (WEIGHTED) pw = lookback_u64(st_wsum, tile, stot);
        if (lane == 0) {
            s_pref[0] = ph;
            s_pref[1] = pw;
            if (base + ST_BIG >= n) { totals[0] = ph + htot; totals[1] = WEIGHTED ? pw + stot : n; }
        }
    }
    __syncthreads();
    const uint64_t ph = s_pref[0], pw = s_pref[1];
    const unsigned lt = lanemask_lt();
#pragma unroll 1
    for (int r = 0; r < ST_ROUNDS; r++) {
        const unsigned hb = (unsigned)(headbits >> (r * ST_ITEMS)) & 0xffffu;
#pragma unroll
        for (int j = 0; j < ST_ITEMS; j++) {
            const uint64_t i = base + (uint64_t)(r * ST_ITEMS + j) * ST_THREADS + tid;
            const bool head = (hb >> j) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, head);
            const int cell = (r * ST_ITEMS + j) * (ST_THREADS / 32) + warp;
            uint64_t wi = 0;
            uint32_t wt = 0;
            if (WEIGHTED) {
                wt = (i < n) ? __ldg(w + i) : 0u;
                wi = warp_incl_scan<uint64_t>(wt);
            }
            if (head) {
                const uint64_t hidx = ph + s_head[cell] + __popc(b & lt);
                out_k[hidx] = __ldg(keys + i);
                // start of the run in the (weighted) prefix space; unweighted: simply its position
                out_start[hidx] = WEIGHTED ? (pw + s_sum[cell] + (wi - wt)) : i;
            }
        }
    }
}

__global__ void rbk_finish_kernel(const uint64_t* __restrict__ start, uint64_t n_out, uint64_t total,
                                  uint32_t* __restrict__ out_c, unsigned int* __restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const uint64_t e = (i + 1 < n_out) ? start[i + 1] : total;
    const uint64_t c = e - start[i];
    if (c > 0xffffffffull) atomicExch(err, 1u);
    out_c[i] = (uint32_t)c;
}

size_t reduce_by_key(Ctx* c, const uint64_t* keys, const uint32_t* w, size_t n, uint64_t* out_k, uint32_t* out_c) {
    if (n == 0) return 0;
    const uint32_t tiles = (uint32_t)div_up(n, ST_BIG);
    DBuf<uint64_t> status(c, (size_t)tiles * 2 + 4);
    DBuf<uint64_t> start(c, n);
    ZB_CUDA(dev_memset(c, status.get(), 0, ((size_t)tiles * 2 + 4) * 8));
    uint64_t* st_heads = status.get();
    uint64_t* st_wsum = status.get() + tiles;
    uint64_t* totals = status.get() + 2 * (size_t)tiles;      // [0]=heads [1]=weight
    uint32_t* ticket = reinterpret_cast<uint32_t*>(totals + 2);  // zeroed
    unsigned int* err = reinterpret_cast<unsigned int*>(totals + 3);
    if (w)
        rbk_kernel<true><<<tiles, ST_THREADS, 0, c->stream>>>(keys, w, n, out_k, start.get(), st_heads, st_wsum, ticket, totals);
    else
        rbk_kernel<false><<<tiles, ST_THREADS, 0, c->stream>>>(keys, w, n, out_k, start.get(), st_heads, st_wsum, ticket, totals);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, totals, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t n_out = c->h_scalars[0], total = c->h_scalars[1];
    rbk_finish_kernel<<<(unsigned)div_up(n_out, 256), 256, 0, c->stream>>>(start.get(), n_out, total, out_c, err);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, err, 4));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (*reinterpret_cast<uint32_t*>(c->h_scalars) != 0)
        ZB_FAIL(ZB_E_RANGE, "k-mer count exceeds 2^32-1 (reference: array('I') OverflowError, kmerize.py:374)");
    return (size_t)n_out;
}

// idx[i] = first position of `keys` (ascending) that is >= probe[i]
__global__ void lower_bound_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint64_t* __restrict__ probe,
                                   uint32_t m, uint64_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t y = probe[i];
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < y) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo;
}
void lower_bound(Ctx* c, const uint64_t* keys, size_t n, const uint64_t* h_probe, size_t m, uint64_t* h_idx) {
    if (m == 0) return;
    DBuf<uint64_t> d(c, 2 * m);
    ZB_CUDA(cudaMemcpyAsync(d.get(), h_probe, m * 8, cudaMemcpyHostToDevice, c->stream));
    lower_bound_kernel<<<(unsigned)div_up(m, 128), 128, 0, c->stream>>>(keys, n, d.get(), (uint32_t)m, d.get() + m);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(cudaMemcpyAsync(h_idx, d.get() + m, m * 8, cudaMemcpyDeviceToHost, c->stream));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
void fill_u32(Ctx* c, uint32_t* p, size_t n, uint32_t v) {
    if (n == 0) return;
    fill_u32_kernel<<<(unsigned)std::min<size_t>(div_up(n, 256), (size_t)c->sm_count * 16), 256, 0, c->stream>>>(p, n, v);
    ZB_LAUNCH_CHECK(c);
}

// ---------------------------------------------------------------------------------------------
// merge-path merge of two sorted (key,count) lists; ties: A first, so equal keys end up adjacent.
// Reference: zotmer/commands/merge.py:26-86 (two-pointer merge); the count sum is done by
// reduce_by_key on the merged stream.
// ---------------------------------------------------------------------------------------------
static constexpr int MG_THREADS = 256;
static constexpr int MG_ITEMS = 8;
static constexpr int MG_TILE = MG_THREADS * MG_ITEMS;  // 2048

// number of A elements among the first `diag` outputs
__device__ __forceinline__ uint64_t merge_path(const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb, uint64_t diag) {
    uint64_t lo = diag > nb ? diag - nb : 0, hi = diag < na ? diag : na;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        // take a[mid] before b[diag-1-mid] iff a[mid] <= b[...]
        if (a[mid] <= b[diag - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void merge_partition_kernel(const uint64_t* __restrict__ a, uint64_t na, const uint64_t* __restrict__ b,
                                       uint64_t nb, uint64_t* __restrict__ part, uint32_t nparts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nparts) return;
    uint64_t diag = (uint64_t)i * MG_TILE;
    if (diag > na + nb) diag = na + nb;
    part[i] = merge_path(a, na, b, nb, diag);
}

__global__ void __launch_bounds__(MG_THREADS)
merge_kernel(const uint64_t* __restrict__ ak, const uint32_t* __restrict__ ac, uint64_t na,
             const uint64_t* __restrict__ bk, const uint32_t* __restrict__ bc, uint64_t nb,
             const uint64_t* __restrict__ part, uint64_t* __restrict__ ok, uint32_t* __restrict__ oc) {
    __shared__ uint64_t sk[MG_TILE];
    __shared__ uint32_t sc[MG_TILE];
    const unsigned tid = threadIdx.x;
    const uint64_t d0 = (uint64_t)blockIdx.x * MG_TILE;
    const uint64_t total = na + nb;
    const uint64_t d1 = min(d0 + MG_TILE, total);
    const uint64_t a0 = part[blockIdx.x], a1 = part[blockIdx.x + 1];
    const uint64_t b0 = d0 - a0, b1 = d1 - a1;
    const uint32_t la = (uint32_t)(a1 - a0), lb = (uint32_t)(b1 - b0);
    for (uint32_t i = tid; i < la; i += MG_THREADS) { sk[i] = ak[a0 + i]; sc[i] = ac ? ac[a0 + i] : 1u; }
    for (uint32_t i = tid; i < lb; i += MG_THREADS) { sk[la + i] = bk[b0 + i]; sc[la + i] = bc ? bc[b0 + i] : 1u; }
    __syncthreads();
    const uint32_t n_loc = la + lb;
    const uint32_t diag = min(tid * MG_ITEMS, n_loc);
    uint32_t i = (uint32_t)merge_path(sk, la, sk + la, lb, diag);
    uint32_t j = diag - i;
    uint64_t rk[MG_ITEMS];
    uint32_t rc[MG_ITEMS];
#pragma unroll
    for (int q = 0; q < MG_ITEMS; q++) {
        const bool ha = i < la, hb = j < lb;
        const uint64_t av = ha ? sk[i] : 0, bv = hb ? sk[la + j] : 0;
        const bool ta = ha && (!hb || av <= bv);
        rk[q] = ta ? av : bv;
        rc[q] = (ha || hb) ? (ta ? sc[i] : sc[la + j]) : 0u;
        if (ta) i++; else j++;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MG_ITEMS; q++) {
        const uint32_t p = tid * MG_ITEMS + q;
        if (p < n_loc) { sk[p] = rk[q]; sc[p] = rc[q]; }
    }
    __syncthreads();
    for (uint32_t p = tid; p < n_loc; p += MG_THREADS) { ok[d0 + p] = sk[p]; oc[d0 + p] = sc[p]; }
}

void merge_pairs(Ctx* c, const uint64_t* ak, const uint32_t* ac, size_t na, const uint64_t* bk, const uint32_t* bc,
                 size_t nb, uint64_t* ok, uint32_t* oc) {
    const size_t total = na + nb;
    if (total == 0) return;
    const uint32_t tiles = (uint32_t)div_up(total, MG_TILE);
    DBuf<uint64_t> part(c, (size_t)tiles + 1);
    merge_partition_kernel<<<(unsigned)div_up((size_t)tiles + 1, 128), 128, 0, c->stream>>>(ak, na, bk, nb, part.get(), tiles + 1);
    ZB_LAUNCH_CHECK(c);
    merge_kernel<<<tiles, MG_THREADS, 0, c->stream>>>(ak, ac, na, bk, bc, nb, part.get(), ok, oc);
    ZB_LAUNCH_CHECK(c);
}

// ---------------------------------------------------------------------------------------------
// mirror: canonical counted set -> the reverse-complement half of the both-strand set.
// rc(k,x) = rev2(~x) >> (64-2k)  (zotmer/library/basics.py:115-121, bits.py:22-31).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rc64(uint64_t x, int k) {
    uint64_t y = __brevll(~x);                                                   // reverse all bits
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);  // un-swap inside pairs
    return y >> (64 - 2 * k);
}

// Odd k: no k-mer equals its own reverse complement, so the mirrored half has exactly n entries and
// is written in place order (no compaction, no atomics).
__global__ void __launch_bounds__(256) mirror_odd_kernel(int k, const uint64_t* __restrict__ ck,
                                                         const uint32_t* __restrict__ cc, uint64_t n,
                                                         uint64_t* __restrict__ rk, uint32_t* __restrict__ rcnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rk[i] = rc64(ck[i], k);
    rcnt[i] = cc[i];
}

// Even k: reverse-palindromes (rc(x) == x) stay single and get their count doubled (the reference
// emits x twice per window, kmerize.py both=True); everything else is appended through one atomic
// reservation per CTA (order is irrelevant, the mirrored half is sorted afterwards).
__global__ void __launch_bounds__(256) mirror_kernel(int k, const uint64_t* __restrict__ ck, uint32_t* __restrict__ cc,
                                                     uint64_t n, uint64_t* __restrict__ rk, uint32_t* __restrict__ rcnt,
                                                     unsigned long long* __restrict__ counter,
                                                     unsigned int* __restrict__ err) {
    __shared__ uint32_t s_scan[256 / 32 + 1];
    __shared__ unsigned long long s_base;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    uint64_t x = 0, r = 0;
    uint32_t cnt = 0;
    bool emit = false;
    if (in) {
        x = ck[i];
        cnt = cc[i];
        r = rc64(x, k);
        if (r == x) {
            if (cnt > 0x7fffffffu) atomicExch(err, 1u);
            cc[i] = cnt * 2u;
        } else {
            emit = true;
        }
    }
    uint32_t tot;
    const uint32_t off = block_excl_scan<256, uint32_t>(emit ? 1u : 0u, s_scan, &tot);
    if (threadIdx.x == 0) s_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
    __syncthreads();
    if (emit) {
        rk[s_base + off] = r;
        rcnt[s_base + off] = cnt;
    }
}

size_t mirror_keys(Ctx* c, int k, const uint64_t* ck, uint32_t* cc, size_t n, uint64_t* rk, uint32_t* rcnt) {
    if (n == 0) return 0;
    if (k & 1) {
        mirror_odd_kernel<<<(unsigned)div_up(n, 256), 256, 0, c->stream>>>(k, ck, cc, n, rk, rcnt);
        ZB_LAUNCH_CHECK(c);
        return n;
    }
    DBuf<unsigned long long> ctr(c, 2);
    ZB_CUDA(dev_memset(c, ctr.get(), 0, 16));
    mirror_kernel<<<(unsigned)div_up(n, 256), 256, 0, c->stream>>>(k, ck, cc, n, rk, rcnt, ctr.get(),
                                                                 reinterpret_cast<unsigned int*>(ctr.get() + 1));
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, ctr.get(), 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_scalars[1] != 0) ZB_FAIL(ZB_E_RANGE, "palindromic k-mer count exceeds 2^32-1");
    return (size_t)c->h_scalars[0];
}

// ---------------------------------------------------------------------------------------------
// order-preserving compaction (trim / project) with a chained scan.
// ---------------------------------------------------------------------------------------------
struct TrimOp {
    const uint32_t* cnt;
    uint64_t cmin, cmax;
    __device__ bool keep(const uint64_t*, uint64_t i) const {
        const uint64_t f = __ldg(cnt + i);
        return f >= cmin && (cmax == 0 || f <= cmax);
    }
    __device__ uint64_t key(uint64_t x) const { return x; }
};
struct ProjectOp {
    int shift;
    __device__ bool keep(const uint64_t* keys, uint64_t i) const {
        return i == 0 || (__ldg(keys + i - 1) >> shift) != (__ldg(keys + i) >> shift);
    }
    __device__ uint64_t key(uint64_t x) const { return x >> shift; }
};

// zotmer/library/basics.py:191-229 murmer(x, s): one 64-bit MurmurHash3 block + fmix64
__device__ __forceinline__ uint64_t murmer64(uint64_t x, uint64_t s) {
    uint64_t k = x * 0x87c37b91114253d5ull;
    k = (k << 31) | (k >> 33);
    k *= 0x4cf5ad432745937full;
    uint64_t h = s ^ k;
    h = (h << 27) | (h >> 37);
    h = h * 5ull + 0x52dce729ull;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
// Deterministic sub-sampling by k-mer hash.  The keep test is the reference's float expression, evaluated in
// IEEE double (int -> double round-to-nearest-even, correctly rounded division), so the decision is bit-exact:
//   mode 0  commands/sample.py:27-34  sampleD:  float(murmer(y, s) & 0xFFFFFFFFFF) / float(0xFFFFFFFFFF) < p
//   mode 1  library/basics.py:251-259 sub():    float(murmer(x, s)) / float(0x1FFFFFFFFFFFFFFF) < p   (kmerize -D)
struct SampleOp {
    int mode;
    uint64_t seed;
    double p;
    __device__ bool keep(const uint64_t* keys, uint64_t i) const {
        const uint64_t h = murmer64(__ldg(keys + i), seed);
        const double u = (mode == 0) ? __ull2double_rn(h & 0xFFFFFFFFFFull) / 1099511627775.0
                                     : __ull2double_rn(h) / 2305843009213693952.0;   // float(2^61 - 1) == 2^61
        return u < p;
    }
    __device__ uint64_t key(uint64_t x) const { return x; }
};
// keep the entries whose k-mer occurs in the sorted reference array (commands/project.py:18-40 project1/2)
struct RestrictOp {
    const uint64_t* ref;
    uint64_t nref;
    __device__ bool keep(const uint64_t* keys, uint64_t i) const {
        const uint64_t y = __ldg(keys + i);
        uint64_t lo = 0, hi = nref;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (__ldg(ref + mid) < y) lo = mid + 1; else hi = mid;
        }
        return lo < nref && __ldg(ref + lo) == y;
    }
    __device__ uint64_t key(uint64_t x) const { return x; }
};

template <typename Op, bool HAS_CNT>
__global__ void __launch_bounds__(ST_THREADS, 4)
compact_kernel(Op op, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cnt, uint64_t n,
               uint64_t* __restrict__ ok, uint32_t* __restrict__ oc, uint64_t* __restrict__ status,
               uint32_t* __restrict__ ticket, uint64_t* __restrict__ total) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_cell[ST_CELLS];
    __shared__ uint64_t s_pref;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * ST_BIG;
    uint64_t keepbits = 0;   // bit r*16+j
#pragma unroll 1
    for (int r = 0; r < ST_ROUNDS; r++) {
        unsigned kb = 0;
#pragma unroll
        for (int j = 0; j < ST_ITEMS; j++) {
            const uint64_t i = base + (uint64_t)(r * ST_ITEMS + j) * ST_THREADS + tid;
            const bool kp = (i < n) && op.keep(keys, i);
            kb |= (kp ? 1u : 0u) << j;
            const unsigned b = __ballot_sync(0xffffffffu, kp);
            if (lane == 0) s_cell[(r * ST_ITEMS + j) * (ST_THREADS / 32) + warp] = __popc(b);
        }
        keepbits |= (uint64_t)kb << (r * ST_ITEMS);
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PERL = ST_CELLS / 32;
        uint32_t hs = 0;
#pragma unroll
        for (int q = 0; q < PERL; q++) hs += s_cell[lane * PERL + q];
        const uint32_t hi = warp_incl_scan(hs);
        const uint32_t htot = __shfl_sync(0xffffffffu, hi, 31);
        uint32_t he = hi - hs;
#pragma unroll
        for (int q = 0; q < PERL; q++) {
            const uint32_t h = s_cell[lane * PERL + q];
            s_cell[lane * PERL + q] = he;
            he += h;
        }
        const uint64_t p = lookback_u64(status, tile, htot);
        if (lane == 0) {
            s_pref = p;
            if (base + ST_BIG >= n) *total = p + htot;
        }
    }
    __syncthreads();
    const uint64_t p0 = s_pref;
    const unsigned lt = lanemask_lt();
#pragma unroll 1
    for (int r = 0; r < ST_ROUNDS; r++) {
        const unsigned kb = (unsigned)(keepbits >> (r * ST_ITEMS)) & 0xffffu;
#pragma unroll
        for (int j = 0; j < ST_ITEMS; j++) {
            const uint64_t i = base + (uint64_t)(r * ST_ITEMS + j) * ST_THREADS + tid;
            const bool kp = (kb >> j) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                const uint64_t o = p0 + s_cell[(r * ST_ITEMS + j) * (ST_THREADS / 32) + warp] + __popc(b & lt);
                ok[o] = op.key(__ldg(keys + i));
                if (HAS_CNT) oc[o] = __ldg(cnt + i);
            }
        }
    }
}

template <typename Op, bool HAS_CNT>
static size_t run_compact(Ctx* c, Op op, const uint64_t* k, const uint32_t* cnt, size_t n, uint64_t* ok, uint32_t* oc) {
    if (n == 0) return 0;
    const uint32_t tiles = (uint32_t)div_up(n, ST_BIG);
    DBuf<uint64_t> status(c, (size_t)tiles + 2);
    ZB_CUDA(dev_memset(c, status.get(), 0, ((size_t)tiles + 2) * 8));
    uint64_t* total = status.get() + tiles;
    uint32_t* ticket = reinterpret_cast<uint32_t*>(status.get() + tiles + 1);
    compact_kernel<Op, HAS_CNT><<<tiles, ST_THREADS, 0, c->stream>>>(op, k, cnt, n, ok, oc, status.get(), ticket, total);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, total, 8));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    return (size_t)c->h_scalars[0];
}

size_t trim_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, uint64_t cmin, uint64_t cmax,
                  uint64_t* ok, uint32_t* oc) {
    TrimOp op{cnt, cmin, cmax};
    return run_compact<TrimOp, true>(c, op, k, cnt, n, ok, oc);
}

size_t sample_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, int mode, uint64_t seed, double p,
                    uint64_t* ok, uint32_t* oc) {
    SampleOp op{mode, seed, p};
    return run_compact<SampleOp, true>(c, op, k, cnt, n, ok, oc);
}

size_t restrict_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, const uint64_t* ref, size_t nref,
                      uint64_t* ok, uint32_t* oc) {
    RestrictOp op{ref, nref};
    return run_compact<RestrictOp, true>(c, op, k, cnt, n, ok, oc);
}

size_t project_keys(Ctx* c, const uint64_t* k, size_t n, int shift, uint64_t* ok) {
    ProjectOp op{shift};
    return run_compact<ProjectOp, false>(c, op, k, nullptr, n, ok, nullptr);
}

// ---------------------------------------------------------------------------------------------
// stats: acgt tallies + histogram of counts with first-occurrence index per distinct count.
// Small counts (< HBINS) use shared-memory bins; the (rare) larger ones go to an overflow list.
// Reference: kmerize.py:492-493,544-545; merge.py:88-92,158-159.
// ---------------------------------------------------------------------------------------------
static constexpr int HBINS = 2048;

__global__ void __launch_bounds__(256)
stats_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cnt, uint64_t n,
             unsigned long long* __restrict__ acgt /*[8]: weighted[4], plain[4]*/,
             unsigned long long* __restrict__ gh /*[HBINS]*/, unsigned long long* __restrict__ gfirst /*[HBINS]*/,
             unsigned long long* __restrict__ ovf_n, uint64_t* __restrict__ ovf_idx, uint32_t* __restrict__ ovf_cnt,
             uint64_t ovf_cap) {
    __shared__ uint32_t sh[HBINS];
    __shared__ unsigned long long sfirst[HBINS];
    __shared__ unsigned long long sacgt[8];
    for (int i = threadIdx.x; i < HBINS; i += blockDim.x) { sh[i] = 0; sfirst[i] = ~0ull; }
    if (threadIdx.x < 8) sacgt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long aw[4] = {0, 0, 0, 0};
    unsigned long long ap[4] = {0, 0, 0, 0};
    const uint64_t per_block = (uint64_t)blockDim.x * 16;
    // Shared-memory atomics are the bound of this kernel (one lane per two cycles per SM): the bin of a count is bumped
    // once per RUN of equal counts a thread sees (at 30x three k-mers in four of a read set are error k-mers with count
    // 1), and the first-occurrence index is only offered when it is lower than what the bin already holds -- the bins
    // only ever decrease, and blocks walk upwards, so almost every offer after a bin's first is refused by a plain load
    // (the 64-bit shared atomicMin is a CAS loop: 81 ms for the 1.25 G entries of the human-scale set before this).
    uint32_t run_f = 0xffffffffu, run_n = 0;
    volatile unsigned long long* vfirst = sfirst;
    for (uint64_t b0 = (uint64_t)blockIdx.x * per_block; b0 < n; b0 += (uint64_t)gridDim.x * per_block) {
#pragma unroll 4
        for (int j = 0; j < 16; j++) {
            const uint64_t i = b0 + (uint64_t)j * blockDim.x + threadIdx.x;
            if (i < n) {
                const uint64_t x = keys[i];
                const uint32_t f = cnt ? cnt[i] : 1u;
                const int bb = (int)(x & 3);
#pragma unroll
                for (int q = 0; q < 4; q++) { if (bb == q) { aw[q] += f; ap[q] += 1; } }
                if (f < HBINS) {
                    if (f == run_f) {
                        run_n++;
                    } else {
                        if (run_n) atomicAdd(&sh[run_f], run_n);
                        run_f = f;
                        run_n = 1;
                        if ((unsigned long long)i < vfirst[f]) atomicMin(&sfirst[f], (unsigned long long)i);
                    }
                } else {
                    const unsigned long long s = atomicAdd(ovf_n, 1ull);
                    if (s < ovf_cap) { ovf_idx[s] = i; ovf_cnt[s] = f; }
                }
            }
        }
    }
    if (run_n) atomicAdd(&sh[run_f], run_n);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned long long a = warp_sum(aw[q]), p = warp_sum(ap[q]);
        if (lane_id() == 0) { atomicAdd(&sacgt[q], a); atomicAdd(&sacgt[4 + q], p); }
    }
    __syncthreads();
    if (threadIdx.x < 8) atomicAdd(&acgt[threadIdx.x], sacgt[threadIdx.x]);
    for (int i = threadIdx.x; i < HBINS; i += blockDim.x) {
        if (sh[i]) { atomicAdd(&gh[i], (unsigned long long)sh[i]); atomicMin(&gfirst[i], sfirst[i]); }
    }
}

void set_stats(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, uint64_t acgt_w[4], uint64_t acgt_p[4],
               uint64_t* total, std::vector<std::pair<uint64_t, uint64_t>>* hist, std::vector<uint64_t>* first_idx) {
    for (int q = 0; q < 4; q++) acgt_w[q] = acgt_p[q] = 0;
    *total = 0;
    hist->clear();
    if (first_idx) first_idx->clear();
    if (n == 0) return;
    size_t ovf_cap = std::min<size_t>(n, 1u << 20);
    Stage st(c, "stats");
    for (int attempt = 0; attempt < 2; attempt++) {
        DBuf<unsigned long long> d(c, 8 + 2 * HBINS + 2);
        DBuf<uint64_t> oidx(c, ovf_cap);
        DBuf<uint32_t> ocnt(c, ovf_cap);
        ZB_CUDA(dev_memset(c, d.get(), 0, (8 + HBINS) * 8));
        ZB_CUDA(dev_memset(c, d.get() + 8 + HBINS, 0xff, HBINS * 8));
        ZB_CUDA(dev_memset(c, d.get() + 8 + 2 * HBINS, 0, 8));
        int blocks = (int)std::min<size_t>((size_t)c->sm_count * 8, div_up(n, 256 * 16));
        stats_kernel<<<blocks, 256, 0, c->stream>>>(k, cnt, n, d.get(), d.get() + 8, d.get() + 8 + HBINS,
                                                    d.get() + 8 + 2 * HBINS, oidx.get(), ocnt.get(), ovf_cap);
        ZB_LAUNCH_CHECK(c);
        // read back by a kernel into mapped pinned memory: a copy-engine transfer of these 33 KB would queue behind the
        // bulk copies other host threads have in flight
        std::vector<unsigned long long> h(8 + 2 * HBINS + 2);
        ZB_CUDA(read_back_big(c, d.get(), h.size() * 8));
        ZB_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(h.data(), c->h_big, h.size() * 8);
        const size_t novf = (size_t)h[8 + 2 * HBINS];
        if (novf > ovf_cap) { ovf_cap = novf; continue; }  // rare: rerun with an exact-size list
        for (int q = 0; q < 4; q++) { acgt_w[q] = h[q]; acgt_p[q] = h[4 + q]; *total += h[q]; }
        // (first index, count value, frequency)
        struct E { uint64_t first, val, freq; };
        std::vector<E> es;
        for (int b = 0; b < HBINS; b++)
            if (h[8 + b]) es.push_back(E{h[8 + HBINS + b], (uint64_t)b, h[8 + b]});
        if (novf) {
            std::vector<uint64_t> oi(novf);
            std::vector<uint32_t> oc(novf);
            ZB_CUDA(cudaMemcpy(oi.data(), oidx.get(), novf * 8, cudaMemcpyDeviceToHost));
            ZB_CUDA(cudaMemcpy(oc.data(), ocnt.get(), novf * 4, cudaMemcpyDeviceToHost));
            std::vector<std::pair<uint32_t, uint64_t>> pr(novf);
            for (size_t i = 0; i < novf; i++) pr[i] = {oc[i], oi[i]};
            std::sort(pr.begin(), pr.end());
            for (size_t i = 0; i < novf;) {
                size_t j = i;
                while (j < novf && pr[j].first == pr[i].first) j++;
                es.push_back(E{pr[i].second, pr[i].first, (uint64_t)(j - i)});
                i = j;
            }
        }
        std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.first < b.first; });
        for (auto& e : es) {
            hist->push_back({e.val, e.freq});
            if (first_idx) first_idx->push_back(e.first);
        }
        return;
    }
    ZB_FAIL(ZB_E_CUDA, "set_stats: overflow list kept growing");
}

// ---------------------------------------------------------------------------------------------
// batched |X n Y| by merge path.  One CTA per 2048-element slice of a pair's merged order.
// Reference: zotmer/library/dist.py:241-265 (split), commands/jaccard.py:31-54.
// abc[3p+0] accumulates the intersection size; the caller derives the two differences.
// ---------------------------------------------------------------------------------------------
static constexpr int IX_THREADS = 256;
static constexpr int IX_ITEMS = 16;
static constexpr int IX_TILE = IX_THREADS * IX_ITEMS;  // 4096

__global__ void __launch_bounds__(IX_THREADS)
pairs_abc_kernel(const SetRef* __restrict__ sets, const uint32_t* __restrict__ I, const uint32_t* __restrict__ J,
                 const uint64_t* __restrict__ tile_start /*[npairs+1]*/, uint32_t npairs,
                 unsigned long long* __restrict__ abc) {
    __shared__ uint64_t sk[IX_TILE + 1];
    __shared__ uint32_t s_pair;
    __shared__ uint64_t s_part[2];
    __shared__ uint32_t s_warp[IX_THREADS / 32];
    const unsigned tid = threadIdx.x;
    const uint64_t t = blockIdx.x;
    if (tid == 0) {
        uint32_t lo = 0, hi = npairs;  // last pair with tile_start <= t
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (tile_start[mid] <= t) lo = mid; else hi = mid; }
        s_pair = lo;
    }
    __syncthreads();
    const uint32_t p = s_pair;
    const SetRef X = sets[I[p]], Y = sets[J[p]];
    const uint64_t total = X.n + Y.n;
    const uint64_t d0 = (t - tile_start[p]) * IX_TILE;
    const uint64_t d1 = min(d0 + IX_TILE, total);
    if (tid < 2) s_part[tid] = merge_path(X.k, X.n, Y.k, Y.n, tid == 0 ? d0 : d1);
    __syncthreads();
    const uint64_t a0 = s_part[0], a1 = s_part[1];
    const uint64_t b0 = d0 - a0, b1 = d1 - a1;
    const uint32_t la = (uint32_t)(a1 - a0), lb = (uint32_t)(b1 - b0);
    // sk[0] = A[a0-1] (halo: a match whose A half fell into the previous slice), then A, then B
    if (tid == 0) sk[0] = (a0 > 0) ? X.k[a0 - 1] : ~0ull;
    for (uint32_t i = tid; i < la; i += IX_THREADS) sk[1 + i] = X.k[a0 + i];
    for (uint32_t i = tid; i < lb; i += IX_THREADS) sk[1 + la + i] = Y.k[b0 + i];
    __syncthreads();
    const uint64_t* A = sk + 1;
    const uint64_t* B = sk + 1 + la;
    const uint32_t n_loc = la + lb;
    const uint32_t diag = min(tid * IX_ITEMS, n_loc);
    uint32_t i = (uint32_t)merge_path(A, la, B, lb, diag);
    uint32_t j = diag - i;
    uint32_t hits = 0;
    const bool halo_valid = a0 > 0;
#pragma unroll
    for (int q = 0; q < IX_ITEMS; q++) {
        const bool ha = i < la, hb = j < lb;
        if (ha || hb) {
            const uint64_t av = ha ? A[i] : 0, bv = hb ? B[j] : 0;
            const bool ta = ha && (!hb || av <= bv);
            if (ta) {
                i++;
            } else {
                // consuming B[j]: it matches iff the A element consumed just before it is equal
                const bool prev_ok = (i > 0) || halo_valid;
                if (prev_ok && A[(int)i - 1] == bv) hits++;
                j++;
            }
        }
    }
    hits = warp_sum(hits);
    if ((tid & 31) == 0) s_warp[tid >> 5] = hits;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < IX_THREADS / 32; w++) s += s_warp[w];
        if (s) atomicAdd(&abc[3 * (size_t)p], (unsigned long long)s);
    }
}

void pairs_abc(Ctx* c, const SetRef* d_sets, const uint32_t* d_I, const uint32_t* d_J, size_t npairs, uint64_t* d_abc,
               uint64_t total_tiles) {
    // d_abc layout: [3*npairs] zeroed by caller, followed by tile_start[npairs+1] prepared by caller
    if (npairs == 0 || total_tiles == 0) return;
    const uint64_t* tile_start = d_abc + 3 * npairs;
    if (total_tiles > 0x7fffffffull) ZB_FAIL(ZB_E_ARG, "pairs_abc: batch too large");
    pairs_abc_kernel<<<(unsigned)total_tiles, IX_THREADS, 0, c->stream>>>(
        d_sets, d_I, d_J, tile_start, (uint32_t)npairs, reinterpret_cast<unsigned long long*>(d_abc));
    ZB_LAUNCH_CHECK(c);
}

}  // namespace zb
