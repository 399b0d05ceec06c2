// inflate_core.cuh -- RFC 1951 (DEFLATE) decoder for ONE gzip member, written for a warp.
//
// Replaces the `gunzip -c` child + pipe of zotmer/library/file.py:93-97 for block-compressed files (BGZF, what
// `bgzip` writes: independent gzip members of <= 64 KiB of text, each carrying its compressed size): every member
// is inflated by one warp, thousands of members at a time, straight into the text buffer the parser reads.
//
// How a warp inflates: the Huffman decode is inherently serial, so ALL W lanes run it redundantly on identical
// state (bit buffer, position) -- uniform control flow, shared-memory table reads are broadcasts, nothing is ever
// exchanged between lanes.  What IS parallel is done by the lanes together: building the decode tables (every
// lane fills its share of the entries by decoding the entry's own index), the byte copies of LZ77 matches (byte j
// of a match at distance d is byte j mod d of the d bytes in front of it), stored blocks, and the literal runs
// (lane p mod W keeps the literal of position p in a register; a run is written with one predicated store).
//
// The same source compiles for the host with W = 1 (tests/host/inflate_host.cpp checks it against zlib without a
// GPU); ZI_SYNC() is __syncwarp() on the device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ZI_HD __host__ __device__ __forceinline__
#else
#define ZI_HD inline
#endif
#ifdef __CUDA_ARCH__
#define ZI_SYNC() __syncwarp()
#else
#define ZI_SYNC() do { } while (0)
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
    ZI_E_CODE = 4,       // over-subscribed code or a bit pattern no code is assigned to
    ZI_E_DIST = 5,       // distance beyond the start of the member
    ZI_E_OUT = 6,        // more (or less) output than the member's ISIZE
    ZI_E_IN = 7,         // ran past the member's compressed bytes
};

// per-warp working set (shared memory on the device): 3.9 KB
struct Scratch {
    uint16_t ltab[1 << LBITS];   // (code length << 9) | symbol; 0 = longer than LBITS bits (or unassigned)
    uint16_t dtab[1 << DBITS];
    uint16_t ctab[1 << CBITS];
    uint16_t lsym[288];          // symbols ordered by (code length, symbol): the canonical decode of long codes
    uint16_t dsym[32];
    uint16_t csym[20];
    uint16_t lcnt[16], dcnt[16], ccnt[16];   // codes per length
    uint8_t lens[320];           // code lengths: literal/length codes, then distance codes
    uint8_t clens[20];
    int flag;
};

// LSB-first bit reader over 32-bit words (the member may start at any byte)
struct Bits {
    const uint32_t* wp;    // next word to load
    uint64_t buf;
    int cnt;               // valid bits in buf
    ZI_HD void init(const uint8_t* p) {
        const unsigned mis = (unsigned)((uintptr_t)p & 3u);
        wp = reinterpret_cast<const uint32_t*>(p - mis);
        buf = (uint64_t)(*wp++) >> (8 * mis);
        cnt = 32 - 8 * (int)mis;
    }
    ZI_HD void refill() {   // afterwards cnt >= 33
        if (cnt <= 32) {
            buf |= (uint64_t)(*wp++) << cnt;
            cnt += 32;
        }
    }
    ZI_HD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    ZI_HD void drop(int n) { buf >>= n; cnt -= n; }
    ZI_HD uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
    ZI_HD const uint8_t* byte_ptr() const { return reinterpret_cast<const uint8_t*>(wp) - (cnt >> 3); }   // cnt a multiple of 8
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
// then fills its share of the 2^tbits table entries.  false: the lengths over-subscribe the code space.
template <int W>
ZI_HD bool build(int lane, const uint8_t* lens, int n, uint16_t* cnt, uint16_t* sym, uint16_t* tab, int tbits, int* flag) {
    ZI_SYNC();   // lens[] written by lane 0
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
    ZI_SYNC();
    if (!*flag) return false;
    for (int e = lane; e < (1 << tbits); e += W) {
        int l = 0;
        const int s = slow_decode((uint32_t)e, tbits, cnt, sym, &l);
        tab[e] = (s < 0) ? (uint16_t)0 : (uint16_t)((l << 9) | s);
    }
    ZI_SYNC();
    return true;
}

ZI_HD int decode(Bits& b, const uint16_t* tab, int tbits, const uint16_t* cnt, const uint16_t* sym) {
    const uint32_t e = tab[b.peek(tbits)];
    if (e) {
        b.drop((int)(e >> 9));
        return (int)(e & 511u);
    }
    int l = 0;
    const int s = slow_decode((uint32_t)b.buf, 15, cnt, sym, &l);
    if (s >= 0) b.drop(l);
    return s;
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

// Inflate the raw deflate stream src[0, clen) into out[0, isize).  W lanes (a power of two) call this together with
// identical arguments and their own `lane`; the result code is the same in every lane.  Reads up to 8 bytes past
// src + clen (never uses them).
template <int W>
ZI_HD int inflate_member(int lane, const uint8_t* src, uint32_t clen, uint8_t* out, uint32_t isize, Scratch* S) {
    Bits b;
    b.init(src);
    const uint32_t* const w_end = reinterpret_cast<const uint32_t*>(src + clen + 11);   // loads beyond: the stream is corrupt
    uint32_t pos = 0, lit0 = 0, mine = 0;
    int last;
    do {
        b.refill();
        last = (int)b.take(1);
        const int type = (int)b.take(2);
        if (type == 3) return ZI_E_BTYPE;
        if (type == 0) {
            flush_literals<W>(lane, out, lit0, pos, mine);
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.take(16);
            b.refill();
            const uint32_t nlen = b.take(16);
            if ((len ^ nlen) != 0xffffu) return ZI_E_STORED;
            const uint8_t* p = b.byte_ptr();
            if (p + len > src + clen) return ZI_E_IN;
            if (pos + len > isize) return ZI_E_OUT;
            for (uint32_t j = (uint32_t)lane; j < len; j += W) out[pos + j] = p[j];
            pos += len;
            lit0 = pos;
            b.init(p + len);
            ZI_SYNC();
            continue;
        }
        int nlit, ndist;
        if (type == 1) {
            nlit = 288;
            ndist = 30;
            ZI_SYNC();   // nobody still reads the previous block's lengths
            for (int s = lane; s < 288; s += W) S->lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
            for (int s = lane; s < 30; s += W) S->lens[288 + s] = 5;
        } else {
            nlit = (int)b.take(5) + 257;
            ndist = (int)b.take(5) + 1;
            const int ncode = (int)b.take(4) + 4;
            if (nlit > 286 || ndist > 30) return ZI_E_HEADER;
            ZI_SYNC();
            if (lane == 0)
                for (int i = 0; i < 19; i++) S->clens[i] = 0;
            ZI_SYNC();
            for (int i = 0; i < ncode; i++) {
                b.refill();
                const uint32_t v = b.take(3);
                if (lane == 0) S->clens[clen_order(i)] = (uint8_t)v;
            }
            if (!build<W>(lane, S->clens, 19, S->ccnt, S->csym, S->ctab, CBITS, &S->flag)) return ZI_E_CODE;
            int i = 0, prev = -1;
            const int total = nlit + ndist;
            while (i < total) {
                b.refill();
                if (b.wp > w_end) return ZI_E_IN;
                const int s = decode(b, S->ctab, CBITS, S->ccnt, S->csym);
                if (s < 0) return ZI_E_CODE;
                int rep = 1, val = s;
                if (s == 16) {
                    if (prev < 0) return ZI_E_HEADER;
                    val = prev;
                    rep = 3 + (int)b.take(2);
                } else if (s == 17) {
                    val = 0;
                    rep = 3 + (int)b.take(3);
                } else if (s == 18) {
                    val = 0;
                    rep = 11 + (int)b.take(7);
                }
                if (i + rep > total) return ZI_E_HEADER;
                if (lane == 0)
                    for (int r = 0; r < rep; r++) S->lens[i + r] = (uint8_t)val;
                i += rep;
                prev = val;
            }
            ZI_SYNC();
            if (S->lens[256] == 0) return ZI_E_HEADER;   // no end-of-block code
        }
        if (!build<W>(lane, S->lens, nlit, S->lcnt, S->lsym, S->ltab, LBITS, &S->flag)) return ZI_E_CODE;
        if (!build<W>(lane, S->lens + nlit, ndist, S->dcnt, S->dsym, S->dtab, DBITS, &S->flag)) return ZI_E_CODE;

        for (;;) {
            b.refill();
            if (b.wp > w_end) return ZI_E_IN;
            int s = decode(b, S->ltab, LBITS, S->lcnt, S->lsym);
            if (s < 0) return ZI_E_CODE;
            if (s < 256) {
                if (pos >= isize) return ZI_E_OUT;
                if (pos - lit0 >= (uint32_t)W) flush_literals<W>(lane, out, lit0, pos, mine);
                if (((pos ^ (uint32_t)lane) & (uint32_t)(W - 1)) == 0) mine = (uint32_t)s;
                pos++;
                continue;
            }
            if (s == 256) break;
            if (s > 285) return ZI_E_CODE;
            uint32_t len;
            if (s < 265) {
                len = (uint32_t)s - 254u;
            } else if (s == 285) {
                len = 258;
            } else {
                const int e = (s - 261) >> 2;
                len = ((4u + (uint32_t)((s - 265) & 3)) << e) + 3u + b.take(e);
            }
            b.refill();
            const int ds = decode(b, S->dtab, DBITS, S->dcnt, S->dsym);
            if (ds < 0 || ds > 29) return ZI_E_CODE;
            uint32_t dist;
            if (ds < 4) {
                dist = (uint32_t)ds + 1u;
            } else {
                const int e = (ds >> 1) - 1;
                dist = ((2u + (uint32_t)(ds & 1)) << e) + 1u + b.take(e);
            }
            if (dist > pos) return ZI_E_DIST;
            if (pos + len > isize) return ZI_E_OUT;
            flush_literals<W>(lane, out, lit0, pos, mine);
            ZI_SYNC();   // the bytes in front of `pos` are in memory for every lane
            {
                const uint8_t* from = out + pos - dist;
                uint8_t* to = out + pos;
                if (dist >= len) {
                    for (uint32_t j = (uint32_t)lane; j < len; j += W) to[j] = from[j];
                } else {
                    for (uint32_t j = (uint32_t)lane; j < len; j += W) to[j] = from[j % dist];
                }
            }
            ZI_SYNC();
            pos += len;
            lit0 = pos;
        }
    } while (!last);
    flush_literals<W>(lane, out, lit0, pos, mine);
    ZI_SYNC();
    if (pos != isize) return ZI_E_OUT;
    // bits used: everything loaded minus what is still buffered
    const uint64_t used_bits = (uint64_t)((const uint8_t*)b.wp - src) * 8u - (uint64_t)b.cnt;
    if (used_bits > (uint64_t)clen * 8u) return ZI_E_IN;
    return ZI_OK;
}

}  // namespace zinf
