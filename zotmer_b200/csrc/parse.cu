// parse.cu -- FASTQ / FASTA text -> dense base-code stream, entirely on the device.
//
// Replaces zotmer/library/file.py:19-36 (readFasta), :38-52 (readFastq) as driven by
// zotmer/library/reads.py:86-125.  Output: one byte per sequence character that the reference's
// k-mer extractor would see, 0..3 for AaCcGgTtUu (basics.py:42-46), 4 for anything else, plus one
// 4 ("break") between records so that no window spans two records.
//
// FASTQ semantics reproduced: lines are split at '\n' only; every 4 lines form a record, line 1 of
// each group is the sequence; a trailing group of fewer than 4 lines is dropped (file.py:51 is never
// true); strip() only removes bytes that are invalid bases anyway, so it needs no special handling.
//
// FASTA semantics reproduced: a line whose first non-blank byte is '>' starts a record; everything
// before the first header is ignored; sequence lines are strip()ped and JOINED, i.e. a maximal
// whitespace run that contains a '\n' (or touches either end of the file) vanishes, while a
// whitespace run inside a line is an ordinary invalid byte.  Whitespace = " \t\n\r\v\f" (py2 strip).
//
// FASTQ: three dependency-free kernels (per-tile counts, one-CTA scan, emit).  FASTA: single pass, a chained scan
// ("decoupled look-back") carries the line state and the output offset from tile to tile.
#include "kernels.h"
#include "fasta_rules.cuh"

namespace zb {

static constexpr int PA_THREADS = 256;
static constexpr int PA_ROWS = 4;
static constexpr int PA_TILE = PA_THREADS * 16 * PA_ROWS;  // 16 KB of text per CTA
static constexpr int PA_WARPS = PA_THREADS / 32;

__device__ __forceinline__ uint4 load16(const uint8_t* raw, uint64_t off, uint64_t n, uint32_t fill) {
    uint4 v;
    if (off + 16 <= n) {
        v = *reinterpret_cast<const uint4*>(raw + off);
    } else {
        uint32_t w[4] = {fill, fill, fill, fill};
        for (int b = 0; b < 16; b++)
            if (off + b < n) w[b >> 2] = (w[b >> 2] & ~(0xffu << (8 * (b & 3)))) | ((uint32_t)raw[off + b] << (8 * (b & 3)));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}
__device__ __forceinline__ uint32_t count_eq(const uint4& v, uint32_t pat) {
    return (__popc(__vcmpeq4(v.x, pat)) + __popc(__vcmpeq4(v.y, pat)) + __popc(__vcmpeq4(v.z, pat)) +
            __popc(__vcmpeq4(v.w, pat))) >> 3;
}

// exclusive scan, in position order, over the [ROWS][WARPS] cells held in shared memory (by warp 0)
__device__ __forceinline__ uint32_t cells_excl_scan(uint32_t* cells /*[ROWS*WARPS]*/) {
    // ROWS*WARPS == 32: one cell per lane
    static_assert(PA_ROWS * PA_WARPS == 32, "cell grid");
    const unsigned l = lane_id();
    const uint32_t v = cells[l];
    const uint32_t inc = warp_incl_scan(v);
    cells[l] = inc - v;
    return __shfl_sync(0xffffffffu, inc, 31);
}

// ------------------------------------------------------------------------------------- FASTQ
// 16 bytes -> 16-bit mask of the bytes equal to the (replicated) pattern
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t pat) {
    const uint32_t x = w ^ pat;
    const uint32_t z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;   // 0x80 where the byte is zero
    return (((z >> 7) * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4& v, uint32_t pat) {
    return eq_mask4(v.x, pat) | (eq_mask4(v.y, pat) << 4) | (eq_mask4(v.z, pat) << 8) | (eq_mask4(v.w, pat) << 12);
}
// 4 text bytes -> 4 base codes (0..3, 4 = not AaCcGgTtUu), branch-free:
// (ch >> 1) & 3 maps A,C,T/U,G -> 0,1,2,3 (case-insensitive); x ^ (x >> 1) swaps 2 <-> 3; the code is
// valid iff the case-folded byte equals "ACGT"[code] (byte permute as a 4-entry table) or is 'U'.
__device__ __forceinline__ uint32_t codes4(uint32_t w) {
    const uint32_t x = (w >> 1) & 0x03030303u;
    const uint32_t c = x ^ ((x >> 1) & 0x01010101u);
    uint32_t sel = (c | (c >> 4)) & 0x00ff00ffu;
    sel = (sel | (sel >> 8)) & 0xffffu;
    const uint32_t expect = __byte_perm(0x54474341u, 0u, sel);
    const uint32_t u = w & 0xdfdfdfdfu;
    uint32_t d = u ^ expect;
    uint32_t ok = ~(((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u;
    d = u ^ 0x55555555u;
    ok |= ~(((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u;
    const uint32_t keep = (ok >> 7) * 3u;                       // 0x03 per valid byte
    return (c & keep) | ((~ok & 0x80808080u) >> 5);             // 0x04 per invalid byte
}

// FASTQ in three dependency-free kernels (no CTA ever waits for another one):
//   fq_count_kernel  per 16 KB tile: number of newlines, and the number of bytes that lie on lines whose index
//                    INSIDE the tile is 0, 1, 2, 3 (mod 4) -- whichever of the four classes turns out to be "line 1
//                    of a record" depends on the line number at the start of the tile, i.e. on a prefix sum
//   fq_scan_kernel   one CTA: line number at the start of every tile -> emitted bytes of every tile -> output offset
//   fastq_kernel     classifies, converts and compacts its tile straight to its final place
// (the first version carried line number and output offset with two chained scans inside one kernel: 38 % of its
// stall samples sat behind those look-backs, profiles/r01_kmerize_step.md)
struct FqTile {
    uint32_t nl;
    uint32_t cls[4];
};

__global__ void __launch_bounds__(PA_THREADS)
fq_count_kernel(const uint8_t* __restrict__ raw, uint64_t n, FqTile* __restrict__ tiles) {
    __shared__ uint32_t s_nl[PA_ROWS * PA_WARPS];
    __shared__ uint32_t s_cls[4];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * PA_TILE;
    if (tid < 4) s_cls[tid] = 0;
    uint32_t nlm[PA_ROWS], nlx[PA_ROWS], lim[PA_ROWS];
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        const uint64_t off = base + (uint64_t)r * (PA_THREADS * 16) + tid * 16;
        const uint4 v = (off < n) ? load16(raw, off, n, 0) : make_uint4(0, 0, 0, 0);
        lim[r] = (off >= n) ? 0u : (off + 16 > n ? (uint32_t)(n - off) : 16u);   // bytes of the chunk inside the text
        nlm[r] = eq_mask16(v, 0x0a0a0a0au) & ((1u << lim[r]) - 1u);
        const uint32_t c = __popc(nlm[r]);
        const uint32_t inc = warp_incl_scan(c);
        nlx[r] = inc - c;
        if (lane == 31) s_nl[r * PA_WARPS + warp] = inc;
    }
    __syncthreads();
    uint32_t tot = 0;
    if (warp == 0) tot = cells_excl_scan(s_nl);
    __syncthreads();
    uint32_t cls[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        uint32_t line = s_nl[r * PA_WARPS + warp] + nlx[r];   // line index inside the tile at the chunk's first byte
        uint32_t rem = nlm[r], start = 0;
        while (start < lim[r]) {
            const uint32_t nxt = rem ? (uint32_t)(__ffs(rem) - 1) : lim[r] - 1u;   // last byte of this line inside the chunk
            const uint32_t len = nxt + 1u - start;
#pragma unroll
            for (int q = 0; q < 4; q++) cls[q] += ((line & 3u) == (uint32_t)q) ? len : 0u;
            if (!rem) break;
            rem &= rem - 1u;
            start = nxt + 1u;
            line++;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t c = warp_sum(cls[q]);
        if (lane == 0 && c) atomicAdd(&s_cls[q], c);
    }
    __syncthreads();
    if (tid == 0) {
        FqTile t;
        t.nl = tot;
        for (int q = 0; q < 4; q++) t.cls[q] = s_cls[q];
        tiles[blockIdx.x] = t;
    }
}

// scalars[0] = number of newlines, scalars[1] = emitted bytes (written by fastq_kernel), scalars[2] = last byte of the text
__global__ void __launch_bounds__(1024)
fq_scan_kernel(const FqTile* __restrict__ tiles, uint32_t ntiles, uint64_t* __restrict__ line0, uint64_t* __restrict__ out0,
               unsigned long long* __restrict__ scalars, const uint8_t* __restrict__ raw, uint64_t n) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = (i < ntiles) ? tiles[i].nl : 0;
        uint64_t tot;
        const uint64_t ex = block_excl_scan<1024, uint64_t>(v, sm, &tot);
        if (i < ntiles) line0[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        scalars[0] = carry;
        scalars[2] = raw[n - 1];   // the host wants to know whether the text ends in a newline
    }
    __syncthreads();   // line0[] of this block's own earlier writes is visible to the whole block
    carry = 0;
    for (uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        // sequence lines are those with global index = 1 (mod 4): local index = 1 - line0 (mod 4)
        const uint64_t v = (i < ntiles) ? tiles[i].cls[(1u - (uint32_t)line0[i]) & 3u] : 0;
        uint64_t tot;
        const uint64_t ex = block_excl_scan<1024, uint64_t>(v, sm, &tot);
        if (i < ntiles) out0[i] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(PA_THREADS)
fastq_kernel(const uint8_t* __restrict__ raw, uint64_t n, const unsigned long long* __restrict__ scalars,
             uint8_t* __restrict__ codes, const uint64_t* __restrict__ tile_line0, const uint64_t* __restrict__ tile_out0,
             unsigned long long* __restrict__ total_out, bool mark) {
    __shared__ uint32_t s_nl[PA_ROWS * PA_WARPS];
    __shared__ uint32_t s_em[PA_ROWS * PA_WARPS];
    __shared__ uint32_t s_tot;
    __shared__ __align__(16) uint8_t s_stage[PA_TILE + 32];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t base = (uint64_t)tile * PA_TILE;
    // complete records only: lines = newlines (+1 if the text does not end in '\n')
    const uint64_t lines = scalars[0] + ((n > 0 && raw[n - 1] != '\n') ? 1 : 0);
    const uint64_t max_line = (lines >> 2) << 2;
    const uint64_t line0 = tile_line0[tile];
    const uint64_t out0 = tile_out0[tile];

    uint4 v[PA_ROWS];
    uint32_t nlm[PA_ROWS];  // newline mask of my 16 bytes
    uint32_t nlx[PA_ROWS];  // newlines before my 16 bytes inside the warp-row
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        const uint64_t off = base + (uint64_t)r * (PA_THREADS * 16) + tid * 16;
        v[r] = (off < n) ? load16(raw, off, n, 0) : make_uint4(0, 0, 0, 0);
        nlm[r] = eq_mask16(v[r], 0x0a0a0a0au);
        if (off + 16 > n) nlm[r] &= (off < n) ? ((1u << (n - off)) - 1u) : 0u;
        const uint32_t c = __popc(nlm[r]);
        const uint32_t inc = warp_incl_scan(c);
        nlx[r] = inc - c;
        if (lane == 31) s_nl[r * PA_WARPS + warp] = inc;
    }
    __syncthreads();
    if (warp == 0) cells_excl_scan(s_nl);
    __syncthreads();

    // emitted bytes: those on line 1 (mod 4) of a complete record, the terminating '\n' included
    uint32_t em[PA_ROWS], emx[PA_ROWS];
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        const uint64_t off = base + (uint64_t)r * (PA_THREADS * 16) + tid * 16;
        uint64_t line = line0 + s_nl[r * PA_WARPS + warp] + nlx[r];
        uint32_t rem = nlm[r], E = 0, start = 0;
        while (true) {
            const uint32_t nxt = rem ? (uint32_t)(__ffs(rem) - 1) : 15u;   // last byte of this line inside the chunk
            if (((line & 3) == 1) && (line < max_line)) E |= ((2u << nxt) - 1u) & ~((1u << start) - 1u);
            if (!rem) break;
            rem &= rem - 1u;
            start = nxt + 1u;
            line++;
            if (start > 15u) break;
        }
        if (off + 16 > n) E &= (off < n) ? ((1u << (n - off)) - 1u) : 0u;
        em[r] = E;
        const uint32_t m = __popc(E);
        const uint32_t inc = warp_incl_scan(m);
        emx[r] = inc - m;
        if (lane == 31) s_em[r * PA_WARPS + warp] = inc;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t tot = cells_excl_scan(s_em);
        if (lane == 0) {
            s_tot = tot;
            // offsets come from counts that ignore the dropped partial record at the very end of the text; every byte
            // that IS emitted sits before the dropped ones, so its offset is exact and the largest end is the total
            if (tot) atomicMax(total_out, (unsigned long long)(out0 + tot));
        }
    }
    __syncthreads();
    const uint32_t tot = s_tot;
    const uint32_t mis = (uint32_t)((uintptr_t)(codes + out0) & 15u);   // stage congruent to the global address
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        uint32_t E = em[r];
        if (E) {
            uint32_t cw[4] = {codes4(v[r].x), codes4(v[r].y), codes4(v[r].z), codes4(v[r].w)};
            if (mark) {   // the '\n' that ends a sequence line becomes 5 = "a record ends here" (capture mode)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t mm = (nlm[r] >> (4 * i)) & 0xfu;
                    cw[i] |= (mm & 1u) | ((mm & 2u) << 7) | ((mm & 4u) << 14) | ((mm & 8u) << 21);
                }
            }
            uint8_t* dst = s_stage + mis + s_em[r * PA_WARPS + warp] + emx[r];
            if (E == 0xffffu) {
#pragma unroll
                for (int b = 0; b < 16; b++) dst[b] = (uint8_t)(cw[b >> 2] >> (8 * (b & 3)));
            } else {
                uint32_t q = 0;
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    if ((E >> b) & 1u) dst[q++] = (uint8_t)(cw[b >> 2] >> (8 * (b & 3)));
                }
            }
        }
    }
    __syncthreads();
    // copy out: whole 16-byte vectors where possible, single bytes at the ragged ends
    uint8_t* gdst = codes + out0 - mis;   // 16-byte aligned
    const uint32_t lo = mis, hi = mis + tot;
    const uint32_t v0 = (lo + 15u) & ~15u, v1 = hi & ~15u;
    if (v0 <= v1) {
        for (uint32_t i = v0 + tid * 16; i < v1; i += PA_THREADS * 16)
            *reinterpret_cast<uint4*>(gdst + i) = *reinterpret_cast<const uint4*>(s_stage + i);
        for (uint32_t i = lo + tid; i < v0 && i < hi; i += PA_THREADS) gdst[i] = s_stage[i];
        for (uint32_t i = v1 + tid; i < hi; i += PA_THREADS) gdst[i] = s_stage[i];
    } else {
        for (uint32_t i = lo + tid; i < hi; i += PA_THREADS) gdst[i] = s_stage[i];
    }
}

void parse_fastq(Ctx* c, const uint8_t* raw, size_t n, uint8_t* codes, size_t* n_codes, uint64_t* n_records, bool mark_records) {
    *n_codes = 0;
    *n_records = 0;
    if (n == 0) return;
    const uint32_t tiles = (uint32_t)div_up(n, PA_TILE);
    DBuf<FqTile> info(c, tiles);
    DBuf<uint64_t> st(c, (size_t)tiles * 2 + 3);
    uint64_t* line0 = st.get();
    uint64_t* out0 = st.get() + tiles;
    unsigned long long* scalars = reinterpret_cast<unsigned long long*>(st.get() + 2 * (size_t)tiles);   // [0] newlines [1] codes [2] last byte
    ZB_CUDA(dev_memset(c, scalars, 0, 24));
    fq_count_kernel<<<tiles, PA_THREADS, 0, c->stream>>>(raw, n, info.get());
    ZB_LAUNCH_CHECK(c);
    fq_scan_kernel<<<1, 1024, 0, c->stream>>>(info.get(), tiles, line0, out0, scalars, raw, n);
    ZB_LAUNCH_CHECK(c);
    fastq_kernel<<<tiles, PA_THREADS, 0, c->stream>>>(raw, n, scalars, codes, line0, out0, scalars + 1, mark_records);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, scalars, 24));   // (a 1-byte cudaMemcpy of the last byte would queue behind other threads' bulk copies)
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    const uint8_t last = (uint8_t)c->h_scalars[2];
    const uint64_t lines = c->h_scalars[0] + (last != '\n' ? 1 : 0);
    *n_records = lines / 4;
    *n_codes = (size_t)c->h_scalars[1];
}

// ------------------------------------------------------------------------------------- FASTA
// (line-state rules: fasta_rules.cuh)
#define FST_AGG 0x40000000u
#define FST_PFX 0x80000000u

__global__ void __launch_bounds__(PA_THREADS)
fasta_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint8_t* __restrict__ codes, uint32_t* __restrict__ st_state,
             uint64_t* __restrict__ st_out, uint32_t* __restrict__ ticket, uint64_t* __restrict__ total_out,
             unsigned long long* __restrict__ n_records, uint32_t brk) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_cell[PA_ROWS * PA_WARPS];   // inclusive transfer map per (row, warp), then exclusive
    __shared__ uint32_t s_bcell[PA_ROWS * PA_WARPS];  // backward (gb, pb) per cell
    __shared__ uint32_t s_em[PA_ROWS * PA_WARPS];
    __shared__ uint32_t s_in;    // tile entry state (f,h,s)
    __shared__ uint32_t s_bin;   // B carry entering the tile from the right
    __shared__ uint64_t s_pref;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * PA_TILE;

    FaMasks mk[PA_ROWS];
    uint4 v[PA_ROWS];
    uint32_t texcl[PA_ROWS];  // composition of the pieces before mine inside my warp-row
    uint32_t bexcl[PA_ROWS];  // backward: composition of the pieces after mine inside my warp-row (bit0=gb, bit1=pb)
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        const uint64_t off = base + (uint64_t)r * (PA_THREADS * 16) + tid * 16;
        // bytes past the end behave like blanks touching the end of the text
        v[r] = (off < n) ? load16(raw, off, n, 0x20202020u) : make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        mk[r] = fa_masks(v[r]);
        const uint32_t t = fa_summary(mk[r]);
        // forward inclusive scan by composition
        uint32_t inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc = fa_compose(u, inc);
        }
        uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
        texcl[r] = (lane == 0) ? FA_IDENT : ex;
        if (lane == 31) s_cell[r * PA_WARPS + warp] = inc;
        // backward: b_left = gb | pb & b_right ; gb = leading blank run holds '\n', pb = all blank & no '\n'
        const FaPiece p0 = fa_piece(mk[r], false, false, false);
        uint32_t bs = ((p0.B & 1u) ? 1u : 0u) | (((t & FA_PF) ? 1u : 0u) << 1);
        uint32_t binc = bs;  // inclusive from the right
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_down_sync(0xffffffffu, binc, o);
            if (lane + o < 32) {
                // mine (left) then u (right):  g = g1 | p1&g2 ; p = p1&p2
                const uint32_t g = (binc & 1u) | (((binc >> 1) & 1u) & (u & 1u));
                const uint32_t pp = ((binc >> 1) & 1u) & ((u >> 1) & 1u);
                binc = g | (pp << 1);
            }
        }
        uint32_t bex = __shfl_down_sync(0xffffffffu, binc, 1);
        bexcl[r] = (lane == 31) ? 2u : bex;  // identity = (g=0,p=1)
        if (lane == 0) s_bcell[r * PA_WARPS + warp] = binc;
    }
    __syncthreads();
    if (warp == 0) {
        // ---- forward: scan the 32 cells, then chain with the previous tiles
        const uint32_t t = s_cell[lane];
        uint32_t inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc = fa_compose(u, inc);
        }
        const uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
        s_cell[lane] = (lane == 0) ? FA_IDENT : ex;
        const uint32_t tile_map = __shfl_sync(0xffffffffu, inc, 31);
        // ---- backward cells
        const uint32_t bt = s_bcell[lane];
        uint32_t binc = bt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_down_sync(0xffffffffu, binc, o);
            if (lane + o < 32) {
                const uint32_t g = (binc & 1u) | (((binc >> 1) & 1u) & (u & 1u));
                const uint32_t pp = ((binc >> 1) & 1u) & ((u >> 1) & 1u);
                binc = g | (pp << 1);
            }
        }
        const uint32_t bex = __shfl_down_sync(0xffffffffu, binc, 1);
        s_bcell[lane] = (lane == 31) ? 2u : bex;
        if (lane == 0) {
            // chained look-back with function composition (sequential walk; predecessors resolve fast)
            uint32_t st_in;
            if (tile == 0) {
                st_in = 1u;  // f = 1 at the start of the text, no header yet
            } else {
                st_volatile_u32(st_state + tile, FST_AGG | tile_map);
                uint32_t comp = FA_IDENT;
                int64_t p = (int64_t)tile - 1;
                while (true) {
                    uint32_t x;
                    do { x = ld_volatile_u32(st_state + p); } while ((x & (FST_AGG | FST_PFX)) == 0);
                    if (x & FST_PFX) { st_in = fa_apply(comp, x & 7u); break; }
                    comp = fa_compose(x & 0x7fu, comp);
                    p--;
                }
            }
            st_volatile_u32(st_state + tile, FST_PFX | fa_apply(tile_map, st_in));
            s_in = st_in;
            // ---- B carry from the right of the tile: does the blank run that starts right after the
            // tile reach a '\n' or the end of the text?  Direct forward scan (runs are short).
            uint64_t q = base + PA_TILE;
            uint32_t bin = 1u;
            while (q < n) {
                const uint8_t ch = raw[q];
                if (ch == '\n') { bin = 1u; break; }
                if (!(ch == ' ' || (ch >= 9 && ch <= 13))) { bin = 0u; break; }
                q++;
            }
            s_bin = bin;
        }
    }
    __syncthreads();
    const uint32_t tile_in = s_in;
    const uint32_t tile_bin = s_bin;

    uint64_t acc[PA_ROWS];
    uint32_t cnt[PA_ROWS], emx[PA_ROWS];
    uint32_t nrec = 0;
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        const uint64_t off = base + (uint64_t)r * (PA_THREADS * 16) + tid * 16;
        const uint32_t st = fa_apply(fa_compose(s_cell[r * PA_WARPS + warp], texcl[r]), tile_in);
        // backward carry: my right neighbours inside the warp-row, then the cells to the right, then the tile's
        const uint32_t bc = s_bcell[r * PA_WARPS + warp];
        const uint32_t bright_cells = (bc & 1u) | (((bc >> 1) & 1u) & tile_bin);
        const uint32_t b_in = (bexcl[r] & 1u) | (((bexcl[r] >> 1) & 1u) & bright_cells);
        const FaPiece p = fa_piece(mk[r], st & 1u, st & 2u, b_in);
        // bytes that come after the first header: everything if s already set, else from the first HS bit on
        uint32_t live = (st & 4u) ? 0xffffu : (p.HS ? (0xffffu & ~((p.HS & (0u - p.HS)) - 1u)) : 0u);
        const uint32_t skip = mk[r].W & (p.F | p.B);
        // emit: header start -> one break; header text -> nothing; vanished blanks -> nothing; else code
        const uint32_t emit = live & ((p.HS) | (~p.HIN & ~skip & 0xffffu));
        uint64_t a = 0;
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const bool in = off + b < n;
            if (in && ((emit >> b) & 1u)) {
                const uint32_t cd = ((p.HS >> b) & 1u) ? brk : code_of(byte_of(v[r], b));
                a |= (uint64_t)cd << (4 * m);
                m++;
            }
        }
        // records: header starts inside the text
        uint32_t hs_valid = p.HS;
        if (off + 16 > n) hs_valid &= (off < n) ? ((1u << (n - off)) - 1u) : 0u;
        nrec += __popc(hs_valid);
        acc[r] = a;
        cnt[r] = m;
        const uint32_t inc = warp_incl_scan(m);
        emx[r] = inc - m;
        if (lane == 31) s_em[r * PA_WARPS + warp] = inc;
    }
    nrec = warp_sum(nrec);
    if (lane == 0 && nrec) atomicAdd(n_records, (unsigned long long)nrec);
    __syncthreads();
    if (warp == 0) {
        const uint32_t tot = cells_excl_scan(s_em);
        const uint64_t p = lookback_u64(st_out, tile, tot);
        if (lane == 0) {
            s_pref = p;
            if (base + PA_TILE >= n) *total_out = p + tot;
        }
    }
    __syncthreads();
    const uint64_t out0 = s_pref;
#pragma unroll
    for (int r = 0; r < PA_ROWS; r++) {
        uint64_t o = out0 + s_em[r * PA_WARPS + warp] + emx[r];
        uint64_t a = acc[r];
        for (uint32_t q = 0; q < cnt[r]; q++) {
            codes[o + q] = (uint8_t)(a & 0xf);
            a >>= 4;
        }
    }
}

void parse_fasta(Ctx* c, const uint8_t* raw, size_t n, uint8_t* codes, size_t* n_codes, uint64_t* n_records, bool mark_records) {
    *n_codes = 0;
    *n_records = 0;
    if (n == 0) return;
    const uint32_t tiles = (uint32_t)div_up(n, PA_TILE);
    DBuf<uint64_t> st(c, (size_t)tiles * 2 + 4);
    ZB_CUDA(dev_memset(c, st.get(), 0, ((size_t)tiles * 2 + 4) * 8));
    uint64_t* st_out = st.get();
    uint32_t* st_state = reinterpret_cast<uint32_t*>(st.get() + tiles);
    uint64_t* total = st.get() + 2 * (size_t)tiles;
    unsigned long long* nrec = reinterpret_cast<unsigned long long*>(st.get() + 2 * (size_t)tiles + 1);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(st.get() + 2 * (size_t)tiles + 2);
    fasta_kernel<<<tiles, PA_THREADS, 0, c->stream>>>(raw, n, codes, st_state, st_out, ticket, total, nrec, mark_records ? 5u : 4u);
    ZB_LAUNCH_CHECK(c);
    ZB_CUDA(read_back(c, total, 16));
    ZB_CUDA(cudaStreamSynchronize(c->stream));
    *n_codes = (size_t)c->h_scalars[0];
    *n_records = c->h_scalars[1];
}

}  // namespace zb
