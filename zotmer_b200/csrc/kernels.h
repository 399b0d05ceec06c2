// kernels.h -- internal host-side entry points of the CUDA translation units (not part of the C ABI).
#pragma once
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

// the handle behind `zb_set*` (include/zotmer_b200.h): a device-resident counted k-mer set
struct zb_set {
    zb::Ctx* c;
    zb::DBuf<uint64_t> k;
    zb::DBuf<uint32_t> cnt;
    size_t n;
    // A merge whose sums pass 2^32-1 (merge.py:145-146 adds Python ints) gives a WIDE set: `wide` holds every count as
    // u64 (what the encoder writes), cnt holds min(count, 2^32-1), and the few entries with cnt == 2^32-1 are listed with
    // their true counts (ascending position).  Empty for every other set.
    zb::DBuf<uint64_t> wide;
    std::vector<uint64_t> exc_idx, exc_key, exc_val;
};

// the handle behind `zb_words*`: the two packed codec64 streams of a set ('kmers' delta-coded, 'counts'), in HBM
struct zb_words {
    zb::Ctx* c;
    zb::DBuf<uint64_t> kw, cw;
    size_t nk = 0, nc = 0;
};

// api.cu: r->k and r->wide are filled -> the saturated u32 counts and the list of entries >= 2^32-1 (an ordinary set if none)
void zb_set_finish_wide(zb_set* r);

// body of every extern "C" entry point: no exception crosses the C boundary
#define ZB_TRY try {
#define ZB_CATCH                                   \
    }                                              \
    catch (const zb::Fail& f) { return f.code; }   \
    catch (const std::bad_alloc&) {                \
        zb::set_error("out of host memory");       \
        return ZB_E_NOMEM;                         \
    }                                              \
    return ZB_OK;

namespace zb {

// ---- api.cu: bulk copies and the per-device engine locks ---------------------------------------
// Large host <-> device copies in 64 MiB pieces, at most three queued (see api.cu)
void copy_chunked(Ctx* c, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
enum { ENG_H2D = 0, ENG_SM = 1, ENG_D2H = 2 };
extern std::mutex g_engine_mu[64][3];
int engine_lock_mask();
struct EngineLock {
    std::unique_lock<std::mutex> lk;
    EngineLock(const Ctx* c, int kind) : lk(g_engine_mu[c->device & 63][kind], std::defer_lock) {
        if ((engine_lock_mask() >> kind) & 1) lk.lock();
    }
};

// ---- sort.cu ---------------------------------------------------------------------------------
// LSD radix sort ("onesweep": one upfront multi-digit histogram, then one chained-scan scatter
// pass per digit) of n < 2^30 u64 keys whose significant bits are [0, key_bits).  Buffers are
// ping-ponged; returns 0 or 1 = which of (k0,v0)/(k1,v1) holds the sorted result.  v0/v1 may be
// null (keys only).  Replaces zotmer/library/misc.py:400-424 (radix_sort).
int radix_sort(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits);
// Same, but only the bits [lo_bit, lo_bit + nbits) take part (stable): the first stage of sort_count.
// stable_first = false: equal keys may come out in any order even when a payload is attached (keys-only sorts
// always take that liberty: it is unobservable) -- the first pass then ranks with one atomic per key.
// Digit histograms gathered while the keys were PRODUCED (extract.cu tallies them on the way out), so that the sort
// does not have to read the keys once more for them (sort_hist_kernel: 1 GB and 0.24 ms of a bench step).  The plan is
// chosen before the number of keys is known; the sort uses the histograms only when its own plan turns out the same.
struct SortPre {
    int passes = 0;
    int shift[4] = {0, 0, 0, 0};
    int bits[4] = {0, 0, 0, 0};
    uint32_t* d_hist = nullptr;   // [passes][256] counts; consumed (scanned in place) by the sort that uses them
};
// digits of an LSD sort of the bits [lo_bit, lo_bit + nbits) with at most maxbits per pass, lowest first
int sort_digit_plan(int lo_bit, int nbits, int maxbits, int* shift, int* bits, int max_passes);
int radix_sort_range(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int lo_bit, int nbits,
                     bool stable_first = true, const SortPre* pre = nullptr);
extern int g_sort_max_bits;
extern int g_sort_cfg;   // onesweep CTA shape, see sort.cu  // digit width cap (8..11), tunable from bench via ZB_SORT_BITS

// ---- segsort.cu ------------------------------------------------------------------------------
// Sort + run-length count in one go (the hot path of kmerize: KmerAccumulator2.flush,
// kmerize.py:412-424 = radix_sort + merge/RLE).  LSD passes run over the TOP bits of the key only, until
// keys that still share those bits form short segments; one kernel then orders every segment in shared
// memory, sums duplicates (weights v0, or 1 each when v0 == null) and writes distinct keys + counts.
// k0/k1 (v0/v1) are ping-pong scratch and are destroyed; out_k/out_c (n entries) must not alias them.
// Returns the number of distinct keys (synchronises).  Result is identical to radix_sort + reduce_by_key.
// distinct = true: the caller guarantees that no key occurs twice (v0 is then a payload that is carried along);
// a violation is detected and reported as ZB_E_ARG.
size_t sort_count(Ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, size_t n, int key_bits,
                  uint64_t* out_k, uint32_t* out_c, bool distinct = false, const SortPre* pre = nullptr);
// the digit plan sort_count's bucket route will use for n keys of key_bits bits (false: none -- tiny inputs, wide digits)
bool sort_count_plan(size_t n, int key_bits, SortPre* plan);
// The both-strand set of kmerize: the sorted canonical counted set (ck, cc, n) united with its nm mirrored pairs
// (mk, mc: distinct, disjoint from ck, unordered; destroyed, mk2 / mc2 are scratch) into out_k / out_c (n + nm entries).
// Returns false when the key space is too skewed for its buckets; M is then still complete in mk/mc (which_out = 0) or
// mk2/mc2 (1) and the caller sorts + merges instead.
bool merge_mirrored(Ctx* c, const uint64_t* ck, const uint32_t* cc, size_t n, uint64_t* mk, uint64_t* mk2, uint32_t* mc,
                    uint32_t* mc2, size_t nm, int key_bits, uint64_t* out_k, uint32_t* out_c, int* which_out);
extern int g_mm_cfg;           // mirror_merge_kernel shape (ZB_MM_CFG): 0 = 256 threads / 2048 entries, 1 = 512 / 4096
extern int g_sort_count_mode;  // 0 = auto, 1 = always the classic full LSD sort + reduce_by_key (ZB_SORT_COUNT)

// ---- nwaymerge.cu ----------------------------------------------------------------------------
// N-way union of sorted counted sets with the counts summed, one pass over the inputs (key-range buckets merged in
// shared memory).  merge.py:26-86, :94-163.  Allocates out_k / out_c (n_out entries).  Returns false when the
// inputs do not suit it (> 1024 of them, key space too skewed for the buckets): the caller falls back to sort_count.
// key_base / key_bits: every key lies in [key_base, key_base + 2^key_bits) -- a slab of the key space is bucketed by
// (key - key_base), so that its buckets are all inside the slab.
bool merge_nway(Ctx* c, const std::vector<const uint64_t*>& ks, const std::vector<const uint32_t*>& cs,
                const std::vector<size_t>& ns, int key_bits, DBuf<uint64_t>* out_k, DBuf<uint32_t>* out_c, size_t* n_out,
                uint64_t key_base = 0);

// The same slab by slab of the key space: at most ~slab_entries input entries are merged (and staged) at a time.
bool merge_nway_slabs(Ctx* c, const std::vector<const uint64_t*>& ks, const std::vector<const uint32_t*>& cs,
                      const std::vector<size_t>& ns, int key_bits, size_t slab_entries, DBuf<uint64_t>* out_k,
                      DBuf<uint32_t>* out_c, size_t* n_out);

// ---- setops.cu -------------------------------------------------------------------------------
// Run-length count of a sorted key array (optionally weighted by `w`): distinct keys + counts.
// Replaces kmerize.py:41-132 (merge with an empty left run) / KmerAccumulator2.flush :412-424.
// Returns the number of distinct keys (synchronises the stream).  Fails with ZB_E_RANGE if a count
// exceeds 2^32-1.
size_t reduce_by_key(Ctx* c, const uint64_t* keys, const uint32_t* w, size_t n, uint64_t* out_k, uint32_t* out_c);
// Merge two sorted (key,count) lists keeping duplicates adjacent (merge-path), n = na + nb outputs.
void merge_pairs(Ctx* c, const uint64_t* ak, const uint32_t* ac, size_t na, const uint64_t* bk, const uint32_t* bc,
                 size_t nb, uint64_t* ok, uint32_t* oc);
// Both-strand expansion of a canonical counted set (kmerize.py both=True semantics):
// writes rc(k, x) for every non-palindromic x (+ its count) and doubles palindrome counts in place.
// Returns number of mirrored keys written.
size_t mirror_keys(Ctx* c, int k, const uint64_t* ck, uint32_t* cc, size_t n, uint64_t* rk, uint32_t* rcnt);
// Compaction cmin <= count (<= cmax when cmax > 0); returns kept count.  trim.py:54-62.
size_t trim_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, uint64_t cmin, uint64_t cmax,
                  uint64_t* ok, uint32_t* oc);
// Hash sub-sampling (mode 0: `zot sample`, commands/sample.py:27-34; mode 1: `zot kmerize -D`, basics.py:251-259)
// and restriction to the k-mers of a reference set (`zot project`, commands/project.py:18-40); both keep order.
size_t sample_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, int mode, uint64_t seed, double p,
                    uint64_t* ok, uint32_t* oc);
size_t restrict_pairs(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, const uint64_t* ref, size_t nref,
                      uint64_t* ok, uint32_t* oc);
// y = x >> shift with adjacent duplicates dropped.  commands/dist.py:36-49.
size_t project_keys(Ctx* c, const uint64_t* k, size_t n, int shift, uint64_t* ok);
// Histogram of counts in first-occurrence order + acgt tallies.  kmerize.py:544-545, merge.py:158-159.
void set_stats(Ctx* c, const uint64_t* k, const uint32_t* cnt, size_t n, uint64_t acgt_w[4], uint64_t acgt_p[4],
               uint64_t* total, std::vector<std::pair<uint64_t, uint64_t>>* hist_first_order,
               std::vector<uint64_t>* first_idx = nullptr /* position of the first entry with each count value */);
void fill_u32(Ctx* c, uint32_t* p, size_t n, uint32_t v);
void lower_bound(Ctx* c, const uint64_t* keys, size_t n, const uint64_t* h_probe, size_t m, uint64_t* h_idx);
// Intersection / difference cardinalities for a batch of pairs.  library/dist.py:241-265.
struct SetRef {
    const uint64_t* k;
    uint64_t n;
};
void pairs_abc(Ctx* c, const SetRef* d_sets, const uint32_t* d_I, const uint32_t* d_J, size_t npairs, uint64_t* d_abc,
               uint64_t total_work);

void pairs_abc_host(Ctx* c, const std::vector<SetRef>& refs, const uint32_t* I, const uint32_t* J, size_t npairs, uint64_t* abc);

// ---- allpairs.cu -----------------------------------------------------------------------------
// All pairs i < j of `refs`: abc[3p .. 3p+2] with p = i (2n - i - 1) / 2 + (j - i - 1).  Sets in blocks of 32; a tile is a
// pair of blocks (row-major over the upper triangle of blocks incl. the diagonal; up to 64 sets: ONE tile), a work unit is
// (tile, one of 8 key-range shards).  Only the units tile_begin, tile_begin + unit_stride, ... < tile_end are computed
// (tile_end 0 = all); abc then holds the cardinalities restricted to those units' key shards and zero elsewhere, so the
// parts of several GPUs add up to the full matrix.  library/dist.py:241-265, jaccard.py:31-54.
uint64_t allpairs_tiles(int nsets);
// With unit_stride > 1 only the units tile_begin, tile_begin + stride, ... are computed: rank r of W takes (r, 0, W) and
// so holds some key-range shards of EVERY tile, which balances ranks whatever the tiles cost.
void allpairs_abc(Ctx* c, const std::vector<SetRef>& refs, uint64_t tile_begin, uint64_t tile_end, uint64_t unit_stride,
                  uint64_t* abc_host);

// ---- inflate.cu ------------------------------------------------------------------------------
// One gzip member of a BGZF file: its raw deflate stream comp[src, src + clen) inflates to out[dst, dst + isize).
struct BgzfMember {
    uint64_t src, dst;
    uint32_t clen, isize;
};
// one warp per member; d_err[0] counts the members that failed, [1] = 1 + the first of them, [2] its code
void bgzf_inflate(Ctx* c, const uint8_t* d_comp, const BgzfMember* d_tab, uint32_t members, uint8_t* d_out, unsigned int* d_err);
// record-aligned cut of FASTA / FASTQ text on the device (see inflate.cu); 0 = none.  Synchronises.
uint64_t text_cut(Ctx* c, const uint8_t* d_text, uint64_t n, bool is_fasta);

// ---- parse.cu --------------------------------------------------------------------------------
// Raw FASTA/FASTQ bytes (device) -> dense base codes (0..3, 4 = break).  `codes` must have room for
// n + 64 bytes.  Returns number of codes written and the number of records (synchronises).
// Replaces file.py:19-52 (readFasta/readFastq) as used by reads.py:86-125.
// mark_records: the code between two records is 5 instead of 4 (both are "no base" to the extractor), so that
// capture_records can tell where a record ends.
void parse_fastq(Ctx* c, const uint8_t* raw, size_t n, uint8_t* codes, size_t* n_codes, uint64_t* n_records,
                 bool mark_records = false);
void parse_fasta(Ctx* c, const uint8_t* raw, size_t n, uint8_t* codes, size_t* n_codes, uint64_t* n_records,
                 bool mark_records = false);

// ---- extract.cu ------------------------------------------------------------------------------
// codes: device pointer to a buffer laid out as [32 bytes of 4][n codes][padding of 4 up to a
// multiple of EXTRACT_TILE + 16]; pointer passed is to the first real code.  Appends the canonical
// k-mer (min(x, rc x)) of every valid window to out[*d_count ...] (unordered within the batch).
// basics.py:303-347.
static const int EXTRACT_TILE = 4096;
// nranks > 1 with d_owner_counts: also adds to d_owner_counts[o] the number of emitted keys whose owner is o.
// pre: also adds the keys' digits to pre->d_hist (see SortPre).
void extract_canonical(Ctx* c, int k, const uint8_t* codes, size_t n, uint64_t* out, unsigned long long* d_count, int nranks = 0,
                       unsigned long long* d_owner_counts = nullptr, const SortPre* pre = nullptr);
// `zot kmerize -C` (kmerize.py:478-483, :507-517): a record is kept, whole, iff one of its k-mers (either strand) is in
// the bait set; every other record of the code stream (records delimited by code 5, see parse_*) is blanked out.
// baits: sorted k-mers closed under reverse complement (a both-strand kmerize of the bait FASTA).
void capture_records(Ctx* c, int k, uint8_t* codes, size_t n, const uint64_t* baits, size_t nbaits);
// multi-GPU: same, but also tallies owner-bucket sizes (owner = mix64(key) range partition).
void bucket_count(Ctx* c, const uint64_t* keys, size_t n, int nranks, unsigned long long* d_bucket_counts);
void bucket_scatter(Ctx* c, const uint64_t* keys, size_t n, int nranks, unsigned long long* d_bucket_cursor,
                    uint64_t* out);
// multi-GPU, fused: route every key into the owner's receive buffer (own memory or a peer's, CUDA IPC)
extern int g_route_per;   // ZB_ROUTE_PER, see extract.cu
struct PeerPtrs {
    uint64_t* p[64];
};
// rcur != nullptr: reserve mode -- rcur->p[r] is rank r's cursor word (peer memory), a CTA reserves its run there with a
// system-scope atomic; cap = keys a receive buffer holds, *d_err is raised when a reservation passes it.
void route_p2p(Ctx* c, const uint64_t* keys, size_t n, int nranks, const PeerPtrs& dst, unsigned long long* d_cursor,
               const PeerPtrs* rcur = nullptr, unsigned long long cap = 0, unsigned int* d_err = nullptr,
               cudaStream_t stream = nullptr /* default: the context's stream */);

}  // namespace zb
