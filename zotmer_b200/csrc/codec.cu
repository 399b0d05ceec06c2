// codec.cu -- codec64 (+ delta) stream codec.  Replaces zotmer/library/codec64.py:82-150 and the
// delta/undelta wrappers of zotmer/library/files.py:85-110.
//
// Word format: low 4 bits = number of values g (1..6), then g fields of 60/g bits, first value in
// the lowest field.  The encoder is greedy: a word takes the longest prefix of the pending values
// (at most 6) whose widest member fits 60/g bits.
//
// Round 1: the codec runs on the HOST in C++ (SURVEY.md 8b allows a host codec first; the GPU codec
// is the first "next" row of 8f).  The set-level entry points move the arrays over PCIe.
#include <vector>

#include "kernels.h"

namespace zb {

static const int kWidth[7] = {0, 60, 30, 20, 15, 12, 10};

static inline int bit_length(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

// returns number of words, or (size_t)-1 when a payload needs more than 60 bits
template <typename T>
static size_t encode_host(const T* vals, size_t n, bool delta, uint64_t* words) {
    size_t w = 0, i = 0;
    uint64_t prev = 0;
    uint64_t grp[6];
    while (i < n) {
        int g = 0, mw = 0;
        uint64_t p = prev;
        while (g < 6 && i + g < n) {
            const uint64_t x = (uint64_t)vals[i + g];
            const uint64_t v = delta ? x - p : x;
            const int nm = std::max(mw, bit_length(v));
            if (nm > kWidth[g + 1]) break;
            grp[g++] = v;
            mw = nm;
            p = x;
        }
        if (g == 0) return (size_t)-1;
        uint64_t word = 0;
        for (int m = g - 1; m >= 0; m--) word = (word << kWidth[g]) | grp[m];
        words[w++] = (word << 4) | (uint64_t)g;
        i += g;
        prev = p;
    }
    return w;
}

static size_t decode_count_host(const uint64_t* words, size_t nw) {
    size_t n = 0;
    for (size_t i = 0; i < nw; i++) {
        const unsigned g = (unsigned)(words[i] & 15);
        if (g < 1 || g > 6) return (size_t)-1;
        n += g;
    }
    return n;
}

template <typename T>
static void decode_host(const uint64_t* words, size_t nw, bool delta, T* out) {
    size_t m = 0;
    uint64_t acc = 0;
    for (size_t i = 0; i < nw; i++) {
        uint64_t w = words[i];
        const int g = (int)(w & 15);
        w >>= 4;
        const uint64_t msk = (1ull << kWidth[g]) - 1;
        for (int q = 0; q < g; q++) {
            const uint64_t v = w & msk;
            w >>= kWidth[g];
            if (delta) { acc += v; out[m++] = (T)acc; } else out[m++] = (T)v;
        }
    }
}

}  // namespace zb

using namespace zb;

struct zb_set_view {  // mirrors the head of zb_set in api.cu
    Ctx* c;
    DBuf<uint64_t> k;
    DBuf<uint32_t> cnt;
    size_t n;
};

extern "C" {

int zb_encode_u64_stream(int device, const uint64_t* vals, size_t n, int delta, uint64_t* words, size_t* n_words) {
    (void)device;
    if ((n && (!vals || !words)) || !n_words) { set_error("null argument"); return ZB_E_ARG; }
    const size_t w = encode_host(vals, n, delta != 0, words);
    if (w == (size_t)-1) {
        set_error("codec64: value or k-mer gap needs more than 60 bits (reference: IndexError, codec64.py:93-99)");
        return ZB_E_RANGE;
    }
    *n_words = w;
    return ZB_OK;
}

int zb_decode_u64_stream(int device, const uint64_t* words, size_t n_words, int delta, uint64_t* out, size_t* n) {
    (void)device;
    if ((n_words && !words) || !n) { set_error("null argument"); return ZB_E_ARG; }
    const size_t cnt = decode_count_host(words, n_words);
    if (cnt == (size_t)-1) { set_error("codec64: corrupt stream (tag outside 1..6)"); return ZB_E_FORMAT; }
    *n = cnt;
    if (out) decode_host(words, n_words, delta != 0, out);
    return ZB_OK;
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
    size_t n = 0;
    int rc = zb_set_size(s, &n);
    if (rc) return rc;
    try {
        std::vector<uint64_t> k(n);
        std::vector<uint32_t> c(n);
        rc = zb_set_fetch(s, k.data(), c.data());
        if (rc) return rc;
        const size_t wk = encode_host(k.data(), n, true, kmer_words);
        const size_t wc = encode_host(c.data(), n, false, count_words);
        if (wk == (size_t)-1 || wc == (size_t)-1) {
            set_error("codec64: value or k-mer gap needs more than 60 bits (reference: IndexError, codec64.py:93-99)");
            return ZB_E_RANGE;
        }
        *n_kmer_words = wk;
        *n_count_words = wc;
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return ZB_E_NOMEM;
    }
    return ZB_OK;
}

int zb_set_from_streams(int device, const uint64_t* kmer_words, size_t n_kmer_words, const uint64_t* count_words,
                        size_t n_count_words, zb_set** out) {
    const size_t nk = decode_count_host(kmer_words, n_kmer_words);
    if (nk == (size_t)-1) { set_error("codec64: corrupt k-mer stream"); return ZB_E_FORMAT; }
    try {
        std::vector<uint64_t> k(nk);
        decode_host(kmer_words, n_kmer_words, true, k.data());
        if (count_words == nullptr) return zb_set_from_host(device, k.data(), nullptr, nk, out);
        const size_t nc = decode_count_host(count_words, n_count_words);
        if (nc == (size_t)-1 || nc != nk) {
            set_error("k-mer and count streams differ in length (%zu vs %zu)", nk, nc);  // files.py:182 assert
            return ZB_E_FORMAT;
        }
        std::vector<uint64_t> c64(nc);
        decode_host(count_words, n_count_words, false, c64.data());
        std::vector<uint32_t> c(nc);
        for (size_t i = 0; i < nc; i++) {
            if (c64[i] > 0xffffffffull) { set_error("count exceeds 2^32-1"); return ZB_E_RANGE; }
            c[i] = (uint32_t)c64[i];
        }
        return zb_set_from_host(device, k.data(), c.data(), nk, out);
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return ZB_E_NOMEM;
    }
}

}  // extern "C"
