// fasta_rules.cuh -- byte classification rules shared by the FASTQ/FASTA parse kernels.
// Pure functions on 16-byte pieces; compiled for the device by parse.cu and for the HOST by
// tests/host/parse_rules_host.cpp, which checks them against the oracle without a GPU.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD static inline
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace zb {

ZB_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

ZB_HD uint32_t code_of(uint32_t ch) {
    // AaCcGgTtUu -> 0..3, else 4.  Branch-free on the folded-to-upper character.
    const uint32_t u = ch & 0xdfu;  // fold case for letters
    uint32_t c = 4u;
    c = (u == 'A') ? 0u : c;
    c = (u == 'C') ? 1u : c;
    c = (u == 'G') ? 2u : c;
    c = (u == 'T' || u == 'U') ? 3u : c;
    // 0xdf folding maps e.g. 'a'(0x61)->'A' but also 0xe1 -> 0xc1: not a letter, fine; it maps
    // nothing that is not already [A-Za-z] onto 'A','C','G','T','U' except bytes >= 0x80 with bit 5
    // set whose fold lands in 0xc1.. (never equal to 0x41..0x55).  So exact.
    return c;
}

ZB_HD uint32_t byte_of(const uint4& v, int b) {
    const uint32_t w = (b < 4) ? v.x : (b < 8) ? v.y : (b < 12) ? v.z : v.w;
    return (w >> (8 * (b & 3))) & 0xffu;
}
// Forward line state after a byte: f = the whitespace run containing it holds a '\n' (or touches
// the start of the text) at or before it; h = inside a header line; s = a header has been seen.
// A segment acts on (f,h,s) as   f' = gf | pf&f,  h' = gh | ph&h | qh&f,  s' = gs | s | qs&f
// and these maps are closed under composition, so they can be scanned.
#define FA_GF 1u
#define FA_PF 2u
#define FA_GH 4u
#define FA_PH 8u
#define FA_QH 16u
#define FA_GS 32u
#define FA_QS 64u
#define FA_IDENT (FA_PF | FA_PH)

ZB_HD uint32_t fa_compose(uint32_t a, uint32_t b) {  // a first, then b
    const bool gf1 = a & FA_GF, pf1 = a & FA_PF, gh1 = a & FA_GH, ph1 = a & FA_PH, qh1 = a & FA_QH, gs1 = a & FA_GS, qs1 = a & FA_QS;
    const bool gf2 = b & FA_GF, pf2 = b & FA_PF, gh2 = b & FA_GH, ph2 = b & FA_PH, qh2 = b & FA_QH, gs2 = b & FA_GS, qs2 = b & FA_QS;
    uint32_t r = 0;
    if (gf2 || (pf2 && gf1)) r |= FA_GF;
    if (pf2 && pf1) r |= FA_PF;
    if (gh2 || (ph2 && gh1) || (qh2 && gf1)) r |= FA_GH;
    if (ph2 && ph1) r |= FA_PH;
    if ((ph2 && qh1) || (qh2 && pf1)) r |= FA_QH;
    if (gs2 || gs1 || (qs2 && gf1)) r |= FA_GS;
    if (qs1 || (qs2 && pf1)) r |= FA_QS;
    return r;
}
// state bits: 1 = f, 2 = h, 4 = s
ZB_HD uint32_t fa_apply(uint32_t t, uint32_t st) {
    const bool f = st & 1u, h = st & 2u, s = st & 4u;
    uint32_t r = 0;
    if ((t & FA_GF) || ((t & FA_PF) && f)) r |= 1u;
    if ((t & FA_GH) || ((t & FA_PH) && h) || ((t & FA_QH) && f)) r |= 2u;
    if ((t & FA_GS) || s || ((t & FA_QS) && f)) r |= 4u;
    return r;
}

struct FaMasks {
    uint32_t W, N, G;  // 16-bit masks: whitespace, newline, '>'
};
ZB_HD FaMasks fa_masks(const uint4& v) {
    FaMasks m{0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t ch = byte_of(v, b);
        const bool ws = (ch == ' ') || (ch >= 9 && ch <= 13);  // \t \n \v \f \r
        m.W |= (ws ? 1u : 0u) << b;
        m.N |= ((ch == '\n') ? 1u : 0u) << b;
        m.G |= ((ch == '>') ? 1u : 0u) << b;
    }
    return m;
}
// carry chain over 16 bits: X(i) = P(i) & (G(i) | X(i-1)), G subset of P, carry-in cin at bit 0
ZB_HD uint32_t chain16(uint32_t P, uint32_t G, bool cin) {
    const uint32_t g = G | ((cin ? 1u : 0u) & P);
    return ((P & ~(P + g)) | g) & 0xffffu;
}
ZB_HD uint32_t rev16(uint32_t x) { return brev32(x) >> 16; }

// per-16-byte analysis given the incoming state; produces per-byte masks
struct FaPiece {
    uint32_t F, B, HS, HIN;
};
ZB_HD FaPiece fa_piece(const FaMasks& m, bool f_in, bool h_in, bool b_in) {
    FaPiece p;
    p.F = chain16(m.W, m.N, f_in);
    p.B = rev16(chain16(rev16(m.W), rev16(m.N), b_in));
    const uint32_t fprev = ((p.F << 1) | (f_in ? 1u : 0u)) & 0xffffu;
    p.HS = m.G & fprev;
    const uint32_t notnl = (~m.N) & 0xffffu;
    p.HIN = chain16(notnl, p.HS, h_in);
    // a '>' inside an already running header line is header text, not a new header
    const uint32_t hprev = ((p.HIN << 1) | (h_in ? 1u : 0u)) & 0xffffu;
    p.HS &= ~hprev;
    return p;
}
ZB_HD uint32_t fa_summary(const FaMasks& m) {
    uint32_t t = 0;
    const bool allws = (m.W == 0xffffu);
    const bool nonl = (m.N == 0);
    if (allws && nonl) t |= FA_PF;
    if (nonl) t |= FA_PH;
    const FaPiece p0 = fa_piece(m, false, false, false);
    if (p0.F & 0x8000u) t |= FA_GF;
    if (!nonl && (p0.HIN & 0x8000u)) t |= FA_GH;
    if (p0.HS) t |= FA_GS;
    // first non-blank byte is '>' and no newline precedes it
    const uint32_t nonws = (~m.W) & 0xffffu;
    if (nonws) {
        const uint32_t first = nonws & (0u - nonws);
        if ((m.G & first) && ((m.N & (first - 1u)) == 0)) {
            t |= FA_QS;
            if (nonl) t |= FA_QH;
        }
    }
    return t;
}


}  // namespace zb
