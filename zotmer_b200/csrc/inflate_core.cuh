// inflate_core.cuh -- RFC 1951 (DEFLATE) decoder for ONE gzip member, written for a warp.
//
// Replaces the `gunzip -c` child + pipe of zotmer/library/file.py:93-97 for block-compressed files (BGZF, what
// `bgzip` writes: independent gzip members of <= 64 KiB of text, each carrying its compressed size): every member
// is inflated by one warp, thousands of members at a time, straight into the text buffer the parser reads.
//
// How a warp inflates: the Huffman decode is inherently serial, so ALL W lanes run it redundantly on identical
// state (bit window, position) -- uniform control flow, shared-memory table reads are broadcasts, nothing is ever
// exchanged between lanes.  What IS parallel is done by the lanes together: building the decode tables (every
// lane fills its share of the entries by decoding the entry's own index), the byte copies of LZ77 matches (byte j
// of a match at distance d is byte j mod d of the d bytes in front of it), stored blocks, and the literal runs
// (lane p mod W keeps the literal of position p in a register; a run is written with one predicated store).
//
// The serial part is kept short (profiles/r02_inflate.md: DNA text is coded mostly as short matches, 12.7 warp
// instructions per byte in the first version):
//   * the bit reader is a funnel shift over two cached words (a third one is already on its way): a symbol costs one
//     shift to look at 32 bits and one add to consume them -- a code and its extra bits are taken from ONE window;
//   * table entries carry the length / distance BASE and the number of extra bits, so a match needs no arithmetic
//     on symbol numbers and no range checks (symbols 286, 287, 30, 31 are marked in their entries);
//   * lanes synchronise only when a match reads bytes written since the last synchronisation (`vis`).
//
// The same source compiles for the host with W = 1 (tests/host/inflate_host.cpp checks it against zlib without a
// GPU); ZI_SYNC(smask) is __syncwarp() on the device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ZI_HD __host__ __device__ __forceinline__
#else
#define ZI_HD inline
#endif
#ifdef __CUDA_ARCH__
#define ZI_SYNC(m) __syncwarp(m)
#elif defined(ZI_HOST_SYNC)
#define ZI_SYNC(m) do { (void)(m); ZI_HOST_SYNC(); } while (0)   // host test with several "lanes" on threads: a real barrier
#else
#define ZI_SYNC(m) do { (void)(m); } while (0)
#endif

namespace zinf {

static constexpr int LBITS = 10;   // literal/length codes up to 10 bits decode with one table read
static constexpr int DBITS = 8;    // distance codes up to 8 bits
static constexpr int CBITS = 7;    // the code-length alphabet's codes are at most 7 bits long

enum {
    ZI_OK = 0,
    ZI_E_BTYPE = 1,      // reserved block type
    ZI_E_STORED = 2,     // LEN / NLEN mismatch
    ZI_E_HEADER = 3,     // too many length / distance codes, repeat without a previous length, lengths overrun
    ZI_E_CODE = 4,       // over-subscribed code, a bit pattern no code is assigned to, or a reserved symbol
    ZI_E_DIST = 5,       // distance beyond the start of the member
    ZI_E_OUT = 6,        // more (or less) output than the member's ISIZE
    ZI_E_IN = 7,         // ran past the member's compressed bytes
};

// Table entries (u32).  bits 0-3: code length (0 = not in the table: a longer code, or no code at all);
// bits 4-7: extra bits; bits 8-9: kind; bits 16-31: the literal, the length base or the distance base.
// A literal entry may hold TWO literals (K_TWO set: bits 0-3 = both code lengths together, second literal in bits
// 24-31) when both codes fit the table index -- text is mostly literals of 3 - 5 bits between short matches
// (3.7 literals per match in FASTQ at level 6), and the second one then costs a few instructions instead of a
// whole trip through the loop.
enum { K_LIT = 0, K_LEN = 1, K_END = 2, K_BAD = 3 };
static constexpr uint32_t K_TWO = 1u << 10;

ZI_HD uint32_t lit_entry(int sym, int len) {
    if (sym < 256) return (uint32_t)len | (K_LIT << 8) | ((uint32_t)sym << 16);
    if (sym == 256) return (uint32_t)len | (K_END << 8);
    if (sym > 285) return (uint32_t)len | (K_BAD << 8);
    uint32_t base, extra;
    if (sym < 265) { base = (uint32_t)sym - 254u; extra = 0; }
    else if (sym == 285) { base = 258; extra = 0; }
    else { extra = (uint32_t)(sym - 261) >> 2; base = ((4u + (uint32_t)((sym - 265) & 3)) << extra) + 3u; }
    return (uint32_t)len | (extra << 4) | (K_LEN << 8) | (base << 16);
}
ZI_HD uint32_t dist_entry(int sym, int len) {
    if (sym > 29) return (uint32_t)len | (K_BAD << 8);
    uint32_t base, extra;
    if (sym < 4) { base = (uint32_t)sym + 1u; extra = 0; }
    else { extra = (uint32_t)(sym >> 1) - 1u; base = ((2u + (uint32_t)(sym & 1)) << extra) + 1u; }
    return (uint32_t)len | (extra << 4) | (base << 16);
}

// per-warp working set (shared memory on the device): 6.5 KB
struct Scratch {
    uint32_t ltab[1 << LBITS];
    uint32_t dtab[1 << DBITS];
    uint16_t ctab[1 << CBITS];   // code-length alphabet: (code length << 9) | symbol, 0 = none
    uint16_t lsym[288];          // symbols ordered by (code length, symbol): the canonical decode of long codes
    uint16_t dsym[32];
    uint16_t csym[20];
    uint16_t lcnt[16], dcnt[16], ccnt[16];   // codes per length
    uint8_t lens[320];           // code lengths: literal/length codes, then distance codes
    uint8_t clens[20];
    int flag;
};

// LSB-first bit reader: a 64-bit window (w0, w1) over the stream's 32-bit words, w2 already loaded.  The member may
// start at any byte.  Beyond `lim` nothing is loaded (zeros are fed: a corrupt stream then runs into an output or
// header check, and the bit count at the end tells).
struct Bits {
    const uint32_t* p0;    // address of the word in w0 (advances even when nothing is loaded any more)
    const uint32_t* lim;
    uint32_t w0, w1, w2;
    uint32_t sh;           // bits of w0 already consumed, 0..31
    ZI_HD void init(const uint8_t* p, const uint8_t* end) {
        const unsigned mis = (unsigned)((uintptr_t)p & 3u);
        p0 = reinterpret_cast<const uint32_t*>(p - mis);
        lim = reinterpret_cast<const uint32_t*>(end - ((uintptr_t)end & 3u));
        w0 = p0[0];
        w1 = p0[1];
        w2 = p0[2];
        sh = 8 * mis;
    }
    ZI_HD uint32_t window() const {   // the next 32 bits
#ifdef __CUDA_ARCH__
        return __funnelshift_r(w0, w1, sh);
#else
        return (uint32_t)((((uint64_t)w1 << 32) | w0) >> sh);
#endif
    }
    ZI_HD void drop(uint32_t n) {     // n <= 32
        sh += n;
        if (sh >= 32) {
            sh -= 32;
            p0++;
            w0 = w1;
            w1 = w2;
            w2 = (p0 + 2 < lim) ? p0[2] : 0u;
        }
    }
    ZI_HD uint32_t take(uint32_t n) {  // n <= 16
        const uint32_t v = window() & ((1u << n) - 1u);
        drop(n);
        return v;
    }
    ZI_HD uint64_t bit_pos(const uint8_t* src) const { return (uint64_t)((const uint8_t*)p0 - src) * 8u + sh; }
    ZI_HD const uint8_t* byte_ptr() const { return reinterpret_cast<const uint8_t*>(p0) + (sh >> 3); }   // sh a multiple of 8
};

// canonical Huffman decode of the code that starts at bit 0 of `bits` (first code bit lowest), at most maxlen bits
ZI_HD int slow_decode(uint32_t bits, int maxlen, const uint16_t* cnt, const uint16_t* sym, int* len_out) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= maxlen; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int count = (int)cnt[len];
        if (code - count < first) {
            *len_out = len;
            return (int)sym[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// tables of one code from its n code lengths.  Lane 0 counts and orders the symbols (a few hundred steps), every lane
// then fills its share of the 2^tbits table entries.  KIND 0: literal/length entries, 1: distance entries, 2: the
// code-length alphabet (u16 entries).  false: the lengths over-subscribe the code space.
template <int W, int KIND>
ZI_HD bool build(int lane, uint32_t smask, const uint8_t* lens, int n, uint16_t* cnt, uint16_t* sym, void* tab, int tbits, int* flag) {
    ZI_SYNC(smask);   // lens[] written by lane 0
    if (lane == 0) {
        int c[16], offs[16];
        for (int i = 0; i < 16; i++) c[i] = 0;
        for (int s = 0; s < n; s++) c[lens[s]]++;
        int left = 1, ok = 1;
        for (int l = 1; l <= 15; l++) {
            left <<= 1;
            left -= c[l];
            if (left < 0) ok = 0;
        }
        offs[0] = 0;
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + c[l];
        for (int s = 0; s < n; s++)
            if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
        c[0] = 0;
        for (int i = 0; i < 16; i++) cnt[i] = (uint16_t)c[i];
        *flag = ok;
    }
    ZI_SYNC(smask);
    if (!*flag) return false;
    for (int e = lane; e < (1 << tbits); e += W) {
        int l = 0;
        const int s = slow_decode((uint32_t)e, tbits, cnt, sym, &l);
        if (KIND == 0) {
            uint32_t ent = (s < 0) ? 0u : lit_entry(s, l);
            if (W >= 2 && s >= 0 && s < 256 && l < tbits) {   // a second literal in the bits behind the first?
                int l2 = 0;
                const int s2 = slow_decode((uint32_t)e >> l, tbits - l, cnt, sym, &l2);
                if (s2 >= 0 && s2 < 256) ent = (uint32_t)(l + l2) | (K_LIT << 8) | K_TWO | ((uint32_t)s << 16) | ((uint32_t)s2 << 24);
            }
            reinterpret_cast<uint32_t*>(tab)[e] = ent;
        } else if (KIND == 1) reinterpret_cast<uint32_t*>(tab)[e] = (s < 0) ? 0u : dist_entry(s, l);
        else reinterpret_cast<uint16_t*>(tab)[e] = (s < 0) ? (uint16_t)0 : (uint16_t)((l << 9) | s);
    }
    ZI_SYNC(smask);
    return true;
}

// order in which the code-length code lengths are stored (RFC 1951 3.2.7), 5 bits each
ZI_HD int clen_order(int i) {
    // 16,17,18,0,8,7,9,6,10,5,11,4 | 12,3,13,2,14,1,15
    const uint64_t a = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) |
                       (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t c = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    return (i < 12) ? (int)((a >> (5 * i)) & 31u) : (int)((c >> (5 * (i - 12))) & 31u);
}

// literals wait in registers: lane p % W holds the byte of position p for the positions [lit0, pos)
template <int W>
ZI_HD void flush_literals(int lane, uint8_t* out, uint32_t& lit0, uint32_t pos, uint32_t mine) {
    if (lit0 < pos) {
        // my position in [lit0, pos): the one congruent to my lane (a run is at most W long)
        const uint32_t p = lit0 + (((uint32_t)lane - lit0) & (uint32_t)(W - 1));
        if (p < pos) out[p] = (uint8_t)mine;
        lit0 = pos;
    }
}

// Inflate the raw deflate stream src[0, clen) into out[0, isize).  W lanes (a power of two: a warp, or an aligned part
// of one -- `smask` names the lanes of the group for its synchronisations) call this together with identical arguments
// and their own `lane` (0 .. W - 1); the result code is the same in every lane.  Reads up to 16 bytes past
// src + clen (never uses them) and may write up to W + 1 bytes past out + isize when the stream is corrupt (the caller
// keeps that much slack behind the last member and rejects the whole group on any error).
template <int W>
ZI_HD int inflate_member(int lane, uint32_t smask, const uint8_t* src, uint32_t clen, uint8_t* out, uint32_t isize, Scratch* S) {
    Bits b;
    const uint8_t* const in_end = src + clen + 16;
    b.init(src, in_end);
    uint32_t pos = 0, lit0 = 0, mine = 0;
    uint32_t vis = 0;   // bytes below `vis` are in memory for every lane (stored before the last synchronisation)
    int last;
    do {
        last = (int)b.take(1);
        const int type = (int)b.take(2);
        if (type == 3) return ZI_E_BTYPE;
        if (type == 0) {
            flush_literals<W>(lane, out, lit0, pos, mine);
            b.drop((32u - b.sh) & 7u);
            const uint32_t len = b.take(16);
            const uint32_t nlen = b.take(16);
            if ((len ^ nlen) != 0xffffu) return ZI_E_STORED;
            const uint8_t* p = b.byte_ptr();
            if (p + len > src + clen) return ZI_E_IN;
            if (pos + len > isize) return ZI_E_OUT;
            for (uint32_t j = (uint32_t)lane; j < len; j += W) out[pos + j] = p[j];
            pos += len;
            lit0 = pos;
            b.init(p + len, in_end);
            continue;
        }
        int nlit, ndist;
        if (type == 1) {
            nlit = 288;
            ndist = 30;
            ZI_SYNC(smask);   // nobody still reads the previous block's lengths
            for (int s = lane; s < 288; s += W) S->lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
            for (int s = lane; s < 30; s += W) S->lens[288 + s] = 5;
        } else {
            nlit = (int)b.take(5) + 257;
            ndist = (int)b.take(5) + 1;
            const int ncode = (int)b.take(4) + 4;
            if (nlit > 286 || ndist > 30) return ZI_E_HEADER;
            ZI_SYNC(smask);
            if (lane == 0)
                for (int i = 0; i < 19; i++) S->clens[i] = 0;
            ZI_SYNC(smask);
            for (int i = 0; i < ncode; i++) {
                const uint32_t v = b.take(3);
                if (lane == 0) S->clens[clen_order(i)] = (uint8_t)v;
            }
            if (!build<W, 2>(lane, smask, S->clens, 19, S->ccnt, S->csym, S->ctab, CBITS, &S->flag)) return ZI_E_CODE;
            int i = 0, prev = -1;
            const int total = nlit + ndist;
            while (i < total) {
                const uint32_t win = b.window();
                uint32_t e = S->ctab[win & ((1u << CBITS) - 1u)];
                if (!e) return ZI_E_CODE;            // codes of this alphabet are at most CBITS long
                const uint32_t l = e >> 9;
                const int s = (int)(e & 511u);
                int rep = 1, val = s;
                uint32_t used = l;
                if (s == 16) {
                    if (prev < 0) return ZI_E_HEADER;
                    val = prev;
                    rep = 3 + (int)((win >> l) & 3u);
                    used += 2;
                } else if (s == 17) {
                    val = 0;
                    rep = 3 + (int)((win >> l) & 7u);
                    used += 3;
                } else if (s == 18) {
                    val = 0;
                    rep = 11 + (int)((win >> l) & 127u);
                    used += 7;
                }
                b.drop(used);
                if (i + rep > total) return ZI_E_HEADER;
                if (lane == 0)
                    for (int r = 0; r < rep; r++) S->lens[i + r] = (uint8_t)val;
                i += rep;
                prev = val;
            }
            ZI_SYNC(smask);
            if (S->lens[256] == 0) return ZI_E_HEADER;   // no end-of-block code
        }
        if (!build<W, 0>(lane, smask, S->lens, nlit, S->lcnt, S->lsym, S->ltab, LBITS, &S->flag)) return ZI_E_CODE;
        if (!build<W, 1>(lane, smask, S->lens + nlit, ndist, S->dcnt, S->dsym, S->dtab, DBITS, &S->flag)) return ZI_E_CODE;

        for (;;) {
            uint32_t win = b.window();
            uint32_t e = S->ltab[win & ((1u << LBITS) - 1u)];
            if ((e & 15u) == 0) {     // a code longer than LBITS bits (rare symbols), or no code
                int l = 0;
                const int s = slow_decode(win, 15, S->lcnt, S->lsym, &l);
                if (s < 0) return ZI_E_CODE;
                e = lit_entry(s, l);
            }
            const uint32_t kind = (e >> 8) & 3u;
            if (kind == K_LIT) {
                b.drop(e & 15u);
                const uint32_t two = (e >> 10) & 1u;
                if (pos - lit0 + two >= (uint32_t)W) {   // the run in the registers is full
                    flush_literals<W>(lane, out, lit0, pos, mine);
                    if (pos > isize) return ZI_E_OUT;
                }
                if (((pos ^ (uint32_t)lane) & (uint32_t)(W - 1)) == 0) mine = e >> 16;
                if (two && (((pos + 1u) ^ (uint32_t)lane) & (uint32_t)(W - 1)) == 0) mine = e >> 24;
                pos += 1u + two;
                continue;
            }
            if (kind != K_LEN) {
                if (kind == K_BAD) return ZI_E_CODE;
                b.drop(e & 15u);
                break;
            }
            const uint32_t cl = e & 15u, xl = (e >> 4) & 15u;
            const uint32_t len = (e >> 16) + ((win >> cl) & ((1u << xl) - 1u));
            b.drop(cl + xl);
            win = b.window();
            uint32_t d = S->dtab[win & ((1u << DBITS) - 1u)];
            if ((d & 15u) == 0) {
                int l = 0;
                const int s = slow_decode(win, 15, S->dcnt, S->dsym, &l);
                if (s < 0) return ZI_E_CODE;
                d = dist_entry(s, l);
            }
            if (d & (K_BAD << 8)) return ZI_E_CODE;
            const uint32_t dl = d & 15u, dx = (d >> 4) & 15u;
            const uint32_t dist = (d >> 16) + ((win >> dl) & ((1u << dx) - 1u));
            b.drop(dl + dx);
            if (dist > pos) return ZI_E_DIST;
            if (pos + len > isize) return ZI_E_OUT;
            flush_literals<W>(lane, out, lit0, pos, mine);
            const uint32_t from0 = pos - dist;
            if (from0 + (dist < len ? dist : len) > vis) {   // the source holds bytes stored since the last synchronisation
                ZI_SYNC(smask);
                vis = pos;
            }
            {
                const uint8_t* from = out + from0;
                uint8_t* to = out + pos;
                if (len <= (uint32_t)W) {      // the common case: every lane at most one byte
                    if ((uint32_t)lane < len) to[lane] = from[dist >= len ? (uint32_t)lane : (dist == 1 ? 0u : (uint32_t)lane % dist)];
                } else if (dist >= len) {
                    for (uint32_t j = (uint32_t)lane; j < len; j += W) to[j] = from[j];
                } else if (dist == 1) {
                    const uint8_t v = from[0];
                    for (uint32_t j = (uint32_t)lane; j < len; j += W) to[j] = v;
                } else {
                    for (uint32_t j = (uint32_t)lane; j < len; j += W) to[j] = from[j % dist];
                }
            }
            pos += len;
            lit0 = pos;
        }
    } while (!last);
    flush_literals<W>(lane, out, lit0, pos, mine);
    ZI_SYNC(smask);
    if (pos != isize) return ZI_E_OUT;
    if (b.bit_pos(src) > (uint64_t)clen * 8u) return ZI_E_IN;
    return ZI_OK;
}

}  // namespace zinf
