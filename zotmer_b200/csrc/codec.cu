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
//   enc_tile_kernel   per 2048-value tile: J[] in shared memory, then six threads walk the tile, one per entry
//                     state -> (exit state, number of words) for each entry state
//   enc_scan_kernel   composes the per-tile maps in order -> entry state and first word index of every tile
//   enc_emit_kernel   re-derives J[], one thread walks from the tile's real entry state and marks the word
//                     starts; every start then packs its word
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
template <typename T>
__device__ __forceinline__ void enc_prepare(const T* __restrict__ vals, uint64_t n, bool delta, uint64_t base,
                                            uint8_t* sbl, uint8_t* sJ, uint64_t* sv, unsigned int* err) {
    const unsigned tid = threadIdx.x;
    for (int p = tid; p < EN_TILE + 5; p += EN_THREADS) {
        const uint64_t i = base + p;
        uint64_t v = 0;
        int bl = 255;   // past the end of the data: never fits, so no group runs over the end
        if (i < n) {
            const uint64_t x = (uint64_t)vals[i];
            v = (delta && i > 0) ? x - (uint64_t)vals[i - 1] : x;
            bl = bitlen64(v);
        }
        sbl[p] = (uint8_t)bl;
        if (sv) sv[p] = v;
    }
    __syncthreads();
    for (int p = tid; p < EN_TILE; p += EN_THREADS) {
        int g = 0, mw = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            mw = max(mw, (int)sbl[p + q]);
            if (g == q && mw <= cw_width(q + 1)) g = q + 1;
        }
        if (g == 0) {   // a value (or gap) wider than 60 bits: an error when it is real data
            if (base + p < n) atomicExch(err, 1u);
            g = 1;      // keeps every walk moving
        }
        sJ[p] = (uint8_t)g;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(EN_THREADS)
enc_tile_kernel(const T* __restrict__ vals, uint64_t n, int delta, uint32_t* __restrict__ tile_info /*[tiles][6]*/,
                unsigned int* __restrict__ err) {
    __shared__ uint8_t sbl[EN_TILE + 8];
    __shared__ uint8_t sJ[EN_TILE];
    const uint64_t base = (uint64_t)blockIdx.x * EN_TILE;
    enc_prepare<T>(vals, n, delta != 0, base, sbl, sJ, nullptr, err);
    if (threadIdx.x < 6) {
        const int lim = (int)min((uint64_t)EN_TILE, n - base);
        int p = threadIdx.x, cnt = 0;   // entry state r: the next word starts r positions into the tile
        while (p < lim) {
            p += sJ[p];
            cnt++;
        }
        tile_info[(size_t)blockIdx.x * 6 + threadIdx.x] = (uint32_t)(p - lim) | ((uint32_t)cnt << 8);
    }
}

// entry state and first word index of every tile.  One CTA: every thread composes a contiguous chunk of tiles
// for all six entry states, thread 0 chains the chunks, every thread then replays its chunk.
__global__ void __launch_bounds__(1024)
enc_scan_kernel(const uint32_t* __restrict__ tile_info, uint32_t tiles, uint8_t* __restrict__ entry,
                uint64_t* __restrict__ woff, uint64_t* __restrict__ total_words) {
    __shared__ uint8_t c_exit[1024][6];
    __shared__ uint32_t c_cnt[1024][6];   // words in a chunk of tiles (a chunk holds far fewer than 2^32 values)
    __shared__ uint32_t c_state[1024];
    __shared__ uint64_t c_off[1024];
    const unsigned t = threadIdx.x;
    const uint32_t chunk = (tiles + 1023) / 1024;
    const uint32_t t0 = min(tiles, t * chunk), t1 = min(tiles, t0 + chunk);
    for (int r = 0; r < 6; r++) {
        uint32_t st = r, cnt = 0;
        for (uint32_t i = t0; i < t1; i++) {
            const uint32_t info = tile_info[(size_t)i * 6 + st];
            st = info & 0xffu;
            cnt += info >> 8;
        }
        c_exit[t][r] = (uint8_t)st;
        c_cnt[t][r] = cnt;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t st = 0;
        uint64_t off = 0;
        for (int i = 0; i < 1024; i++) {
            c_state[i] = st;
            c_off[i] = off;
            off += c_cnt[i][st];
            st = c_exit[i][st];
        }
        *total_words = off;
    }
    __syncthreads();
    uint32_t st = c_state[t];
    uint64_t off = c_off[t];
    for (uint32_t i = t0; i < t1; i++) {
        entry[i] = (uint8_t)st;
        woff[i] = off;
        const uint32_t info = tile_info[(size_t)i * 6 + st];
        st = info & 0xffu;
        off += info >> 8;
    }
}

template <typename T>
__global__ void __launch_bounds__(EN_THREADS)
enc_emit_kernel(const T* __restrict__ vals, uint64_t n, int delta, const uint8_t* __restrict__ entry,
                const uint64_t* __restrict__ woff, uint64_t* __restrict__ words, unsigned int* __restrict__ err) {
    __shared__ uint8_t sbl[EN_TILE + 8];
    __shared__ uint8_t sJ[EN_TILE];
    __shared__ uint64_t sv[EN_TILE + 8];
    __shared__ uint32_t sstart[EN_TILE / 32];
    __shared__ uint32_t s_scan[EN_THREADS / 32 + 1];
    const unsigned tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * EN_TILE;
    if (tid < EN_TILE / 32) sstart[tid] = 0;
    enc_prepare<T>(vals, n, delta != 0, base, sbl, sJ, sv, err);
    const int lim = (int)min((uint64_t)EN_TILE, n - base);
    if (tid == 0) {
        int p = entry[blockIdx.x];
        while (p < lim) {
            sstart[p >> 5] |= 1u << (p & 31);
            p += sJ[p];
        }
    }
    __syncthreads();
    // thread t owns positions 8t .. 8t+7 = byte t of the start mask
    const uint32_t mine = (sstart[tid >> 2] >> ((tid & 3) * 8)) & 0xffu;
    uint32_t tot;
    uint32_t rank = block_excl_scan<EN_THREADS, uint32_t, false>(__popc(mine), s_scan, &tot);
    uint32_t m = mine;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int p = tid * EN_PER + b;
        const int g = sJ[p];
        const int wd = cw_width(g);
        uint64_t word = 0;
        for (int q = g - 1; q >= 0; q--) word = (word << wd) | sv[p + q];
        words[woff[blockIdx.x] + rank] = (word << 4) | (uint64_t)g;
        rank++;
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
// vals (device, n values) -> words (device, capacity n); returns the number of words
template <typename T>
static size_t encode_dev(Ctx* c, const T* d_vals, size_t n, bool delta, uint64_t* d_words) {
    if (n == 0) return 0;
    const uint32_t tiles = (uint32_t)div_up(n, EN_TILE);
    DBuf<uint32_t> info(c, (size_t)tiles * 6);
    DBuf<uint8_t> entry(c, tiles);
    DBuf<uint64_t> woff(c, (size_t)tiles + 2);
    uint64_t* total = woff.get() + tiles;
    unsigned int* err = reinterpret_cast<unsigned int*>(woff.get() + tiles + 1);
    ZB_CUDA(dev_memset(c, total, 0, 16));
    enc_tile_kernel<T><<<tiles, EN_THREADS, 0, c->stream>>>(d_vals, n, delta ? 1 : 0, info.get(), err);
    ZB_LAUNCH_CHECK(c);
    enc_scan_kernel<<<1, 1024, 0, c->stream>>>(info.get(), tiles, entry.get(), woff.get(), total);
    ZB_LAUNCH_CHECK(c);
    enc_emit_kernel<T><<<tiles, EN_THREADS, 0, c->stream>>>(d_vals, n, delta ? 1 : 0, entry.get(), woff.get(), d_words, err);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, total, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    if (reinterpret_cast<uint32_t*>(c->h_scalars + 1)[0] != 0)
        ZB_FAIL(ZB_E_RANGE, "codec64: value or k-mer gap needs more than 60 bits (reference: IndexError, codec64.py:93-99)");
    return (size_t)c->h_scalars[0];
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
    const zb_set* v = s;
    Ctx* c = v->c;
    ZB_CUDA(cudaSetDevice(c->device));
    *n_kmer_words = *n_count_words = 0;
    if (v->n == 0) return ZB_OK;
    if (!kmer_words || !count_words) ZB_FAIL(ZB_E_ARG, "null argument");
    DBuf<uint64_t> dw(c, v->n);
    Stage st(c, "encode");
    size_t nw = encode_dev<uint64_t>(c, v->k.get(), v->n, true, dw.get());
    ZB_CUDA(cudaMemcpyAsync(kmer_words, dw.get(), nw * 8, cudaMemcpyDeviceToHost, c->stream));
    *n_kmer_words = nw;
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    nw = encode_dev<uint32_t>(c, v->cnt.get(), v->n, false, dw.get());
    ZB_CUDA(cudaMemcpyAsync(count_words, dw.get(), nw * 8, cudaMemcpyDeviceToHost, c->stream));
    *n_count_words = nw;
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    ZB_CATCH
}

int zb_set_from_streams(int device, const uint64_t* kmer_words, size_t n_kmer_words, const uint64_t* count_words,
                        size_t n_count_words, zb_set** out) {
    ZB_TRY
    if (!out || (n_kmer_words && !kmer_words)) ZB_FAIL(ZB_E_ARG, "null argument");
    Ctx* c = ctx_for(device);
    Stage st(c, "decode");
    DBuf<uint64_t> dkw(c, n_kmer_words), dcw;
    if (n_kmer_words) ZB_CUDA(cudaMemcpyAsync(dkw.get(), kmer_words, n_kmer_words * 8, cudaMemcpyHostToDevice, c->stream));
    if (count_words) {
        dcw.alloc(c, n_count_words);
        if (n_count_words) ZB_CUDA(cudaMemcpyAsync(dcw.get(), count_words, n_count_words * 8, cudaMemcpyHostToDevice, c->stream));
    }
    DecodePlan pk, pc;
    decode_plan(c, dkw.get(), n_kmer_words, &pk);
    if (count_words) {
        decode_plan(c, dcw.get(), n_count_words, &pc);
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
        decode_emit<uint64_t>(c, dkw.get(), n_kmer_words, true, &pk, v->k.get());
        if (count_words) {
            decode_emit<uint32_t>(c, dcw.get(), n_count_words, false, &pc, v->cnt.get());
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

}  // extern "C"
