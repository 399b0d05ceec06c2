/*
 * zot_oracle.c -- plain-C restatement of the reference's k-mer hot path (CPU, single thread).
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/libzot_oracle.so by oracle/Makefile and loaded
 * through oracle/c_oracle.py by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product library (zotmer_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every entry point against the
 * fixtures in tests/golden/data (outputs of the reference's own code) and against the Python
 * restatement oracle/zot_oracle.py.
 *
 * Same algorithms as the reference (paths relative to /root/reference), stated once more in C so
 * that multi-megabase inputs finish in seconds:
 *   zo_extract      zotmer/library/file.py:19-52 (readFasta/readFastq) +
 *                   zotmer/library/basics.py:303-347 (kmersList, both strands) +
 *                   zotmer/commands/kmerize.py:492-493 (acgt tally)
 *   zo_sort_count   zotmer/library/misc.py:400-424 (radix_sort) + kmerize.py:41-132 (merge/RLE)
 *   zo_merge        zotmer/commands/merge.py:26-86,127-163
 *   zo_split        zotmer/library/dist.py:241-265
 *   zo_trim         zotmer/commands/trim.py:54-62
 *   zo_hist         kmerize.py:544-545 / merge.py:158 (first-occurrence order)
 *   zo_encode/decode zotmer/library/codec64.py:82-150, files.py:85-110 (delta)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int nuc(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return -1;
    }
}

static int is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0b || c == 0x0c; }

typedef struct {
    uint64_t *v;
    size_t n, cap;
    int count_only;
} keybuf;

static int kb_push(keybuf *b, uint64_t x) {
    if (b->count_only) { b->n++; return 0; }
    if (b->n == b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 1024;
        uint64_t *nv = (uint64_t *)realloc(b->v, nc * sizeof(uint64_t));
        if (!nv) return -1;
        b->v = nv; b->cap = nc;
    }
    b->v[b->n++] = x;
    return 0;
}

/* basics.py:303-347: rolling forward word x and reverse-complement word xb; any byte outside
 * AaCcGgTtUu restarts the window.  Emits x then xb for every full window. */
static int extract_seq(int k, const uint8_t *s, size_t z, keybuf *out, uint64_t acgt[4]) {
    const uint64_t msk = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int sh = 2 * (k - 1);
    uint64_t x = 0, xb = 0;
    int j = 0;
    for (size_t i = 0; i < z; i++) {
        int b = nuc(s[i]);
        if (b < 0) { j = 0; x = 0; xb = 0; continue; }
        x = ((x << 2) | (uint64_t)b) & msk;
        xb = (xb >> 2) | ((uint64_t)(3 - b) << sh);
        if (j < k) j++;
        if (j == k) {
            acgt[x & 3]++; acgt[xb & 3]++;
            if (kb_push(out, x) || kb_push(out, xb)) return -1;
        }
    }
    return 0;
}

/* Parse FASTA (is_fasta != 0) or 4-line FASTQ bytes, append both-strand k-mers of every record.
 * If keys == NULL only counts.  Returns 0, or -1 on allocation failure / -2 if cap too small. */
int zo_extract(int k, const uint8_t *data, size_t n, int is_fasta,
               uint64_t *keys, size_t cap, size_t *n_keys, uint64_t acgt[4], uint64_t *n_records) {
    keybuf kb = {NULL, 0, 0, keys == NULL};
    uint8_t *seq = NULL; size_t sn = 0, scap = 0;
    int have_rec = 0; uint64_t nrec = 0; int rc = 0;
    size_t pos = 0; unsigned line_in_grp = 0;
    /* FASTQ: a record only counts once its 4th line has been seen (file.py:45-52) */
    const uint8_t *fq_seq = NULL; size_t fq_len = 0;
    while (pos < n) {
        size_t e = pos;
        while (e < n && data[e] != '\n') e++;
        size_t a = pos, b = e;                       /* strip(): file.py:27 / :47 */
        while (a < b && is_space(data[a])) a++;
        while (b > a && is_space(data[b - 1])) b--;
        if (is_fasta) {
            if (b > a && data[a] == '>') {
                if (have_rec) { if (extract_seq(k, seq, sn, &kb, acgt)) { rc = -1; goto done; } nrec++; }
                have_rec = 1; sn = 0;
            } else if (have_rec) {
                if (sn + (b - a) > scap) {
                    scap = (sn + (b - a)) * 2 + 64;
                    uint8_t *ns = (uint8_t *)realloc(seq, scap);
                    if (!ns) { rc = -1; goto done; }
                    seq = ns;
                }
                memcpy(seq + sn, data + a, b - a); sn += b - a;
            }
        } else {
            if (line_in_grp == 1) { fq_seq = data + a; fq_len = b - a; }
            if (++line_in_grp == 4) {
                if (extract_seq(k, fq_seq, fq_len, &kb, acgt)) { rc = -1; goto done; }
                nrec++; line_in_grp = 0;
            }
        }
        pos = e + 1;
    }
    if (is_fasta && have_rec) { if (extract_seq(k, seq, sn, &kb, acgt)) { rc = -1; goto done; } nrec++; }
    if (keys) {
        if (kb.n > cap) { rc = -2; goto done; }
    }
done:
    if (keys && rc == 0) memcpy(keys, kb.v, kb.n * sizeof(uint64_t));
    *n_keys = kb.n; *n_records = nrec;
    free(kb.v); free(seq);
    return rc;
}

/* LSD byte radix sort (result == ascending sort == misc.py:400-424), then run-length count
 * (kmerize.py:41-132 with an empty left operand).  keys is clobbered. */
int zo_sort_count(uint64_t *keys, size_t n, uint64_t *kmers, uint32_t *counts, size_t *n_distinct) {
    uint64_t *tmp = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    if (!tmp) return -1;
    uint64_t *src = keys, *dst = tmp;
    for (int pass = 0; pass < 8; pass++) {
        size_t h[256]; memset(h, 0, sizeof h);
        int sh = pass * 8;
        for (size_t i = 0; i < n; i++) h[(src[i] >> sh) & 255]++;
        if (n && h[(src[0] >> sh) & 255] == n) continue;   /* digit constant: nothing to do */
        size_t s = 0;
        for (int d = 0; d < 256; d++) { size_t c = h[d]; h[d] = s; s += c; }
        for (size_t i = 0; i < n; i++) dst[h[(src[i] >> sh) & 255]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    size_t m = 0;
    for (size_t i = 0; i < n;) {
        uint64_t y = src[i]; uint32_t c = 0;
        while (i < n && src[i] == y) { c++; i++; }
        kmers[m] = y; counts[m] = c; m++;
    }
    *n_distinct = m;
    free(tmp);
    return 0;
}

/* merge.py:26-86 + :127-163: union of nsets sorted duplicate-free (kmer,count) lists, counts
 * summed.  Pairwise tree of two-pointer merges; out arrays must hold sum(lens) entries. */
static size_t merge2(const uint64_t *xk, const uint64_t *xc, size_t xn,
                     const uint64_t *yk, const uint64_t *yc, size_t yn, uint64_t *ok, uint64_t *oc) {
    size_t i = 0, j = 0, m = 0;
    while (i < xn && j < yn) {
        if (xk[i] < yk[j]) { ok[m] = xk[i]; oc[m] = xc[i]; i++; }
        else if (xk[i] > yk[j]) { ok[m] = yk[j]; oc[m] = yc[j]; j++; }
        else { ok[m] = xk[i]; oc[m] = xc[i] + yc[j]; i++; j++; }
        m++;
    }
    for (; i < xn; i++, m++) { ok[m] = xk[i]; oc[m] = xc[i]; }
    for (; j < yn; j++, m++) { ok[m] = yk[j]; oc[m] = yc[j]; }
    return m;
}

int zo_merge(int nsets, const uint64_t *const *kmers, const uint64_t *const *counts, const size_t *lens,
             uint64_t *out_k, uint64_t *out_c, size_t *n_out) {
    size_t tot = 0;
    for (int s = 0; s < nsets; s++) tot += lens[s];
    uint64_t *ak = (uint64_t *)malloc((tot ? tot : 1) * 8), *ac = (uint64_t *)malloc((tot ? tot : 1) * 8);
    uint64_t *bk = (uint64_t *)malloc((tot ? tot : 1) * 8), *bc = (uint64_t *)malloc((tot ? tot : 1) * 8);
    if (!ak || !ac || !bk || !bc) { free(ak); free(ac); free(bk); free(bc); return -1; }
    size_t an = 0;
    for (int s = 0; s < nsets; s++) {
        size_t m = merge2(ak, ac, an, kmers[s], counts[s], lens[s], bk, bc);
        uint64_t *t;
        t = ak; ak = bk; bk = t; t = ac; ac = bc; bc = t; an = m;
    }
    memcpy(out_k, ak, an * 8); memcpy(out_c, ac, an * 8);
    *n_out = an;
    free(ak); free(ac); free(bk); free(bc);
    return 0;
}

/* library/dist.py:241-265 */
void zo_split(const uint64_t *xs, size_t xz, const uint64_t *ys, size_t yz, uint64_t abc[3]) {
    size_t i = 0, j = 0; uint64_t b = 0, dx = 0, dy = 0;
    while (i < xz && j < yz) {
        if (xs[i] < ys[j]) { dx++; i++; }
        else if (xs[i] > ys[j]) { dy++; j++; }
        else { b++; i++; j++; }
    }
    abc[0] = b; abc[1] = dx + (xz - i); abc[2] = dy + (yz - j);
}

/* commands/dist.py:36-49: y = x >> shift, drop adjacent duplicates.  Returns new length. */
size_t zo_project(const uint64_t *xs, size_t n, int shift, uint64_t *out) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t y = xs[i] >> shift;
        if (m == 0 || out[m - 1] != y) out[m++] = y;
    }
    return m;
}

/* commands/trim.py:54-62: keep c <= f (and f <= C when C > 0) */
size_t zo_trim(const uint64_t *xs, const uint64_t *cs, size_t n, uint64_t c, uint64_t C,
               uint64_t *ox, uint64_t *oc) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (cs[i] >= c && (C == 0 || cs[i] <= C)) { ox[m] = xs[i]; oc[m] = cs[i]; m++; }
    return m;
}

/* kmerize.py:544-545: histogram of counts, distinct values reported in order of first
 * occurrence (== Python dict insertion order == JSON key order).  vals/freqs hold <= n entries. */
size_t zo_hist(const uint64_t *cs, size_t n, uint64_t *vals, uint64_t *freqs) {
    /* open-addressing table: value -> slot in vals/freqs */
    size_t cap = 64; size_t m = 0;
    int64_t *tab = (int64_t *)malloc(cap * sizeof(int64_t));
    for (size_t i = 0; i < cap; i++) tab[i] = -1;
    for (size_t i = 0; i < n; i++) {
        uint64_t c = cs[i];
        size_t h = (size_t)((c * 0x9E3779B97F4A7C15ULL) >> 7) & (cap - 1);
        while (tab[h] >= 0 && vals[tab[h]] != c) h = (h + 1) & (cap - 1);
        if (tab[h] < 0) {
            vals[m] = c; freqs[m] = 0; tab[h] = (int64_t)m; m++;
            if (m * 2 > cap) {                                   /* grow + rehash */
                size_t nc = cap * 2; int64_t *nt = (int64_t *)malloc(nc * sizeof(int64_t));
                for (size_t q = 0; q < nc; q++) nt[q] = -1;
                for (size_t q = 0; q < m; q++) {
                    size_t g = (size_t)((vals[q] * 0x9E3779B97F4A7C15ULL) >> 7) & (nc - 1);
                    while (nt[g] >= 0) g = (g + 1) & (nc - 1);
                    nt[g] = (int64_t)q;
                }
                free(tab); tab = nt; cap = nc;
                h = (size_t)((c * 0x9E3779B97F4A7C15ULL) >> 7) & (cap - 1);
                while (vals[tab[h]] != c) h = (h + 1) & (cap - 1);
            }
        }
        freqs[tab[h]]++;
    }
    free(tab);
    return m;
}

static int bitlen(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

/* codec64.py:82-120 (+ files.py:85-98 when delta != 0).  Greedy: a word takes the longest prefix
 * of n <= 6 pending values whose widest member fits 60/n bits.  Returns 0, or -3 when a value
 * (or k-mer gap) needs more than 60 bits (the reference raises IndexError / struct.error).
 * words must hold n entries. */
int zo_encode(const uint64_t *vals, size_t n, int delta, uint64_t *words, size_t *n_words) {
    static const int width[7] = {0, 60, 30, 20, 15, 12, 10};
    size_t w = 0, i = 0; uint64_t prev = 0;
    uint64_t grp[6];
    while (i < n) {
        int g = 0, mw = 0;
        uint64_t p = prev;
        while (g < 6 && i + g < n) {
            uint64_t v = delta ? vals[i + g] - p : vals[i + g];
            int bl = bitlen(v);
            int nm = bl > mw ? bl : mw;
            if (nm > width[g + 1]) break;
            grp[g] = v; mw = nm; p = vals[i + g]; g++;
        }
        if (g == 0) return -3;
        uint64_t word = 0;
        for (int m = g - 1; m >= 0; m--) word = (word << width[g]) | grp[m];
        words[w++] = (word << 4) | (uint64_t)g;
        i += g; prev = p;
    }
    *n_words = w;
    return 0;
}

/* codec64.py:122-150 (+ files.py:100-110).  out must hold 6*n_words entries. Returns count or
 * (size_t)-1 for a tag outside 1..6 (never produced by the encoder). */
size_t zo_decode(const uint64_t *words, size_t n_words, int delta, uint64_t *out) {
    static const int width[7] = {0, 60, 30, 20, 15, 12, 10};
    size_t m = 0; uint64_t acc = 0;
    for (size_t i = 0; i < n_words; i++) {
        uint64_t w = words[i]; int g = (int)(w & 15); w >>= 4;
        if (g < 1 || g > 6) return (size_t)-1;
        uint64_t msk = (1ULL << width[g]) - 1;
        for (int q = 0; q < g; q++) {
            uint64_t v = w & msk; w >>= width[g];
            if (delta) { acc += v; out[m++] = acc; } else out[m++] = v;
        }
    }
    return m;
}
