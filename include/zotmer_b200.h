/*
 * zotmer_b200.h -- C ABI of libzot_b200.so, the B200 (sm_100a) implementation of zotmer's k-mer hot path.
 *
 * The reference (drtconway/zotmer) is pure Python and has no FFI; the seam this library fills is
 * "zotmer/commands + zotmer/library  ->  per-k-mer work", i.e. every place where the reference runs
 * an interpreter loop over k-mers.  Each entry point below names the reference call site(s) it
 * replaces (paths relative to the reference root).  The Python binding a maintainer would add is
 * zotmer_b200/_native.py (ctypes); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 (ZB_OK) or a negative ZB_E_* code; zb_last_error() gives the message
 *     of the last failure on the calling thread.  No exception crosses this boundary.
 *   - the caller owns every host buffer; handles (zb_set*, zb_kmerizer*) own device memory until
 *     the matching *_free / *_close.  Sizes are obtained first (two-call pattern), then fetched.
 *   - plain pointers and sizes only.  `*_dev` variants take DEVICE pointers on the handle's device
 *     (used by bench.py for the HBM-resident measurement and by the multi-GPU exchange).
 *   - one host thread drives one device context; calls on one handle must not be concurrent.
 *   - k-mer encoding is the reference's: A=0 C=1 G=2 T/U=3, first base most significant, k <= 32
 *     (zotmer/library/basics.py:42-59).  A "counted set" is a strictly ascending u64 k-mer array
 *     with a parallel u32 count array (kmerize.py:373-374 `array('L')`/`array('I')`).
 */
#ifndef ZOTMER_B200_H
#define ZOTMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_OK 0
#define ZB_E_CUDA (-1)     /* a CUDA runtime call failed */
#define ZB_E_ARG (-2)      /* bad argument (k out of 1..32, null pointer, ...) */
#define ZB_E_RANGE (-3)    /* value does not fit: codec64 payload > 60 bits (codec64.py:93-99 IndexError),
                              count > 2^32-1 (kmerize.py:374 array('I') OverflowError) */
#define ZB_E_FORMAT (-4)   /* corrupt codec64 stream (files.py:58 / codec64.py:128 KeyError) */
#define ZB_E_NOMEM (-5)
#define ZB_E_NOGPU (-6)    /* no CUDA device: this library has no CPU path */

typedef struct zb_set zb_set;             /* device-resident counted k-mer set */
typedef struct zb_kmerizer zb_kmerizer;   /* streaming kmerize+count state */
typedef struct zb_words zb_words;         /* the two packed codec64 streams of a set ('kmers', 'counts'), device-resident */
typedef struct zb_encplan zb_encplan;     /* first phase of a range-partitioned encode (several GPUs, one file) */
typedef struct zb_staged zb_staged;       /* input text on its way to the device (asynchronous, pinned staging ring) */

const char* zb_last_error(void);
int zb_version(void);
int zb_device_count(int* n);
/* number of kernels this library has launched on `device` from the calling thread (bench "gpu_launches") */
int zb_launch_count(int device, uint64_t* n);
int zb_device_sync(int device);
/* a few bytes (<= 64 KiB, multiple of 4) between host memory and device memory, moved by a kernel through mapped pinned
 * memory instead of a copy engine: an engine serves the transfers of all host threads in the order they were queued, so
 * a 16-byte copy can wait milliseconds behind another thread's 315 MB input (the multi-GPU exchange's count matrix and
 * flags go this way).  Both return after the transfer has completed. */
int zb_dev_read_small(int device, const void* d_src, size_t bytes, void* host_dst);
int zb_dev_write_small(int device, void* d_dst, const void* host_src, size_t bytes);
/* give the device memory cached by the library (freed sets, scratch) back to the driver */
int zb_release_cache(int device);

/* ------------------------------------------------------------------------------------------
 * kmerize + count          replaces zotmer/commands/kmerize.py:463-545
 *   reads()/readFasta/readFastq (library/reads.py:35-125, file.py:19-52), kmersList(K,seq,True)
 *   (basics.py:303-347), the acgt tally (kmerize.py:492-493), KmerAccumulator2.addList/flush
 *   (kmerize.py:401-424), radix_sort (misc.py:400-424) and merge() (kmerize.py:41-132).
 *
 * feed(): `raw` is the text of ONE input file, or a record-aligned piece of it (a piece must start
 * at the first byte of a record line group / '>' header line and is parsed exactly as a whole file
 * would be: FASTA per file.py:19-36, 4-line FASTQ per file.py:38-52 incl. the dropped trailing
 * partial record).  n < 2^31.  Every window of k consecutive AaCcGgTtUu bytes inside one record
 * contributes the forward k-mer AND its reverse complement (both=True).
 * finish(): sorts, counts and returns the both-strand counted set; the kmerizer can be closed after.
 * ------------------------------------------------------------------------------------------ */
int zb_kmerize_open(int k, int device, zb_kmerizer** out);
int zb_kmerize_feed(zb_kmerizer* h, const uint8_t* raw, size_t n, int is_fasta);
int zb_kmerize_feed_dev(zb_kmerizer* h, const uint8_t* d_raw, size_t n, int is_fasta);
/* capture mode -- `zot kmerize -C BAITS`, kmerize.py:478-483 (the bait k-mers) and :507-517 (`if found:
 * buf.addList(xs)`): from now on a fed record contributes ALL its k-mers iff one of them is in `baits` (a both-strand
 * k-mer set of the bait sequences, e.g. the result of kmerizing the bait FASTA); other records only count as records.
 * `baits` must stay alive while the kmerizer is fed; NULL switches the mode off. */
int zb_kmerize_set_baits(zb_kmerizer* h, const zb_set* baits);
/* pre-parsed input: one code per base (0..3 = ACGT, 4 = break) -- used by the synthetic read generator */
int zb_kmerize_feed_codes_dev(zb_kmerizer* h, const uint8_t* d_codes, size_t n, uint64_t n_records);
int zb_kmerize_finish(zb_kmerizer* h, zb_set** result, uint64_t* n_records);
int zb_kmerize_close(zb_kmerizer* h);
/* multi-GPU (hash-range all-to-all): extract canonical k-mers of everything fed so far WITHOUT
 * counting, bucketed by owner = (mix64(kmer) >> 32) * nranks >> 32.  bucket_counts[nranks] (host)
 * receives the sizes; d_keys (device, capacity >= zb_kmerize_pending) the keys grouped by owner. */
/* the keys this kmerizer extracts from now on will be exchanged among `nranks` GPUs: extraction then also tallies the keys
 * per owner, and zb_kmerize_bucket_counts(h, nranks) returns the tallies instead of making a pass over the keys (0 = off) */
int zb_kmerize_set_owners(zb_kmerizer* h, int nranks);
int zb_kmerize_pending(zb_kmerizer* h, uint64_t* n_keys);
int zb_kmerize_take_bucketed_dev(zb_kmerizer* h, int nranks, uint64_t* d_keys, uint64_t* bucket_counts);
/* The same exchange fused into one kernel over NVLink peer memory: zb_kmerize_bucket_counts tells how many of the
 * pending keys each owner gets (so that the ranks can agree on write offsets), zb_kmerize_route_p2p then writes
 * every pending key to d_dst[owner][...] -- d_dst[r] points into rank r's receive buffer (this rank's own memory
 * for r == self, a buffer opened with zb_ipc_open otherwise), already advanced to this rank's slot in it.
 * Returns after all stores have completed; the pending list is empty afterwards. */
int zb_kmerize_bucket_counts(zb_kmerizer* h, int nranks, uint64_t* bucket_counts);
int zb_kmerize_route_p2p(zb_kmerizer* h, int nranks, uint64_t* const* d_dst);
/* zb_kmerize_route_p2p in two halves: _begin launches the routing kernel on a second stream and returns at once, _end
 * waits for it; in between the caller may count what the PREVIOUS exchange delivered (zb_kmerize_flush), so that the sort
 * of batch i - 1 runs while batch i's keys travel over NVLink.  Nothing may be fed between the two calls. */
int zb_kmerize_route_p2p_begin(zb_kmerizer* h, int nranks, uint64_t* const* d_dst);
int zb_kmerize_route_p2p_end(zb_kmerizer* h);
/* The same without zb_kmerize_bucket_counts and without any agreement between the ranks beforehand: d_dst[r] is the
 * START of rank r's receive buffer (capacity_keys keys) and d_cursor[r] a u64 word in rank r's memory, zero before the
 * step; every thread block reserves its run with one system-scope atomic add on the owner's word (over NVLink for a
 * remote owner) and stores it there.  After all ranks have returned (a barrier), *d_cursor[self] is the number of keys
 * received; their order depends on timing.  sent_counts[nranks] (host, may be NULL) receives what this rank sent to
 * whom.  ZB_E_RANGE when a reservation passed capacity_keys (nothing was stored out of bounds). */
int zb_kmerize_route_p2p_reserve(zb_kmerizer* h, int nranks, uint64_t* const* d_dst, uint64_t* const* d_cursor,
                                 uint64_t capacity_keys, uint64_t* sent_counts);
/* one process driving several GPUs (the `zot` commands with ZB_GPUS=N): kernels on `device` may then load from and
 * store to memory of `peer` (NVLink peer mapping, cudaDeviceEnablePeerAccess); buffers from zb_ipc_alloc and the arrays
 * of sets on `peer` become valid arguments of the *_dev entry points on `device` */
int zb_peer_enable(int device, int peer);
/* device buffers that other processes on the node can map (cudaIpc*): alloc/free on the owner, open/close on peers */
int zb_ipc_alloc(int device, size_t bytes, void** d_ptr, uint8_t handle[64]);
int zb_ipc_open(int device, const uint8_t handle[64], void** d_ptr);
int zb_ipc_close(int device, void* d_ptr);
int zb_ipc_free(int device, void* d_ptr);
/* count canonical keys received from peers (device array, any order) into the kmerizer */
int zb_kmerize_add_canonical_dev(zb_kmerizer* h, const uint64_t* d_keys, size_t n);
/* the same without the copy: the caller's array itself is sorted in place (its contents are destroyed) when the
 * kmerizer next counts -- at the latest in zb_kmerize_finish; it must stay valid and untouched until then */
int zb_kmerize_adopt_canonical_dev(zb_kmerizer* h, uint64_t* d_keys, size_t n);
/* count NOW whatever is pending (extracted but uncounted keys, an adopted array) and fold it into the running counted
 * set -- KmerAccumulator2.flush, kmerize.py:412-424; afterwards an adopted array may be reused by its owner */
int zb_kmerize_flush(zb_kmerizer* h);

/* ------------------------------------------------------------------------------------------
 * counted sets
 * ------------------------------------------------------------------------------------------ */
/* upload a sorted, duplicate-free set; counts may be NULL (all 1) -- readKmersAndCounts result, files.py:219-227 */
int zb_set_from_host(int device, const uint64_t* kmers, const uint32_t* counts, size_t n, zb_set** out);
/* the same from DEVICE arrays on `device` (copied) -- e.g. a share received from another GPU */
int zb_set_from_device(int device, const uint64_t* d_kmers, const uint32_t* d_counts, size_t n, zb_set** out);
/* key-range sharding (multi-GPU merge): idx[i] = first position whose k-mer is >= probes[i];
 * zb_set_slice copies the entries [begin, end) into a new set */
int zb_set_lower_bound(const zb_set* s, const uint64_t* probes, size_t m, uint64_t* idx);
int zb_set_slice(const zb_set* s, size_t begin, size_t end, zb_set** out);
int zb_set_size(const zb_set* s, size_t* n);
int zb_set_fetch(const zb_set* s, uint64_t* kmers, uint32_t* counts); /* either may be NULL */
int zb_set_dev_ptrs(const zb_set* s, const uint64_t** d_kmers, const uint32_t** d_counts);
int zb_set_free(zb_set* s);

/* hist = {count: number of k-mers} with distinct counts in order of FIRST OCCURRENCE along the
 * sorted set (the reference's dict insertion order -> JSON key order; kmerize.py:544-545,
 * merge.py:158); acgt_weighted[x&3] += count (kmerize.py:492-493 == merge.py:159);
 * acgt_plain[x&3] += 1 (merge.py:88-92, the 2-input path).  Fills min(hist_cap, *n_hist) entries; retry when *n_hist > hist_cap. */
int zb_set_stats(const zb_set* s, uint64_t acgt_weighted[4], uint64_t acgt_plain[4], uint64_t* total_count,
                 uint64_t* hist_vals, uint64_t* hist_freqs, size_t hist_cap, size_t* n_hist);

/* N-way union with counts summed -- merge.py:26-86 (merge) + :127-163 (mergeNinto).
 * The reference sums Python ints and writes them with codec64 (up to 60 bits): when a sum passes 2^32-1 the result is
 * a WIDE set -- zb_set_is_wide; its counts come back as u64 through zb_set_fetch_counts64 (zb_set_fetch saturates them
 * at 2^32-1), zb_set_stats and zb_set_encode* work with the true counts, zb_set_from_streams* gives a wide set for
 * streams that hold such counts, zb_merge takes wide inputs, and trim / sample / restrict / slice of one keep the true
 * counts of what they keep (trim thresholds beyond 2^32-1 on a wide set: ZB_E_RANGE). */
int zb_merge(int nsets, zb_set* const* sets, zb_set** out);
int zb_set_is_wide(const zb_set* s, int* wide);
int zb_set_fetch_counts64(const zb_set* s, uint64_t* counts);   /* any set: its counts as u64 */
/* keep cmin <= count (and count <= cmax when cmax > 0) -- trim.py:54-62 */
int zb_trim(const zb_set* s, uint64_t cmin, uint64_t cmax, zb_set** out);
/* deterministic sub-sampling by k-mer hash, order kept.  mode 0: keep x iff
 * float(murmer(x, seed) & 0xFFFFFFFFFF) / float(0xFFFFFFFFFF) < p  -- commands/sample.py:27-34 (sampleD; the only path
 * `zot sample` ever takes under docopt);  mode 1: float(murmer(x, seed)) / float(0x1FFFFFFFFFFFFFFF) < p --
 * library/basics.py:251-259 sub(), the filter of `zot kmerize -D` (kmerize.py:494-506).  Same IEEE double expression. */
int zb_sample(const zb_set* s, int mode, uint64_t seed, double p, zb_set** out);
/* keep the entries of s whose k-mer occurs in ref -- commands/project.py:18-40 (project1 / project2) */
int zb_restrict(const zb_set* s, const zb_set* ref, zb_set** out);
/* y = x >> shift_bits, adjacent duplicates dropped, counts discarded -- commands/dist.py:36-49 (Measure.prep) */
int zb_project(const zb_set* s, int shift_bits, zb_set** out);
/* (|X n Y|, |X \ Y|, |Y \ X|) for each listed pair -- library/dist.py:241-265 (split), jaccard.py:31-54.
 * abc holds 3*npairs u64.  All sets must live on the same device. */
int zb_pairs_abc(int nsets, zb_set* const* sets, const uint32_t* I, const uint32_t* J, size_t npairs, uint64_t* abc);
/* The same for ALL pairs i < j -- the loop nests of commands/dist.py:145-168 and jaccard.py:148-166 (-a).
 * abc holds 3 * nsets (nsets - 1) / 2 u64, pair (i, j) at p = i (2 nsets - i - 1) / 2 + (j - i - 1).
 * The work is cut into zb_allpairs_tiles(nsets) independent units (a pair of blocks of 32 sets x one of 8
 * key-range shards of the k-mer space); only units [tile_begin, tile_end) are computed (tile_end 0 = all): abc then
 * holds the cardinalities restricted to the k-mers of those shards and 0 for pairs outside those units, so that the
 * parts computed on several GPUs simply add up (the "final gather" of the distance matrix). */
int zb_allpairs_tiles(int nsets, uint64_t* n_tiles);
int zb_allpairs_abc(int nsets, zb_set* const* sets, uint64_t tile_begin, uint64_t tile_end, uint64_t* abc);
/* The same over the units unit_begin, unit_begin + unit_stride, ... < unit_end (unit_end 0 = all).  Rank r of W GPUs
 * takes (r, 0, W): it then holds some key-range shards of every tile, so the ranks stay balanced however unevenly the
 * tiles cost (sets of different sizes or divergence). */
int zb_allpairs_abc_strided(int nsets, zb_set* const* sets, uint64_t unit_begin, uint64_t unit_end, uint64_t unit_stride,
                            uint64_t* abc);

/* ------------------------------------------------------------------------------------------
 * stream codec             replaces zotmer/library/codec64.py:82-150 + files.py:85-110 (delta)
 * host arrays in, host arrays out; the work runs on `device`.
 * encode: words must hold n entries (worst case one value per word).  ZB_E_RANGE when a value
 * (or, with delta, a gap) needs more than 60 bits.
 * decode: two-call -- out == NULL returns the value count in *n.
 * ------------------------------------------------------------------------------------------ */
int zb_encode_u64_stream(int device, const uint64_t* vals, size_t n, int delta, uint64_t* words, size_t* n_words);
int zb_decode_u64_stream(int device, const uint64_t* words, size_t n_words, int delta, uint64_t* out, size_t* n);
/* device-resident forms: encode a set straight to its two file streams / build a set from them
 * (writeKmersAndCounts2 files.py:209-217, readKmersAndCounts files.py:219-227) */
int zb_set_encode(const zb_set* s, uint64_t* kmer_words, size_t* n_kmer_words, uint64_t* count_words, size_t* n_count_words);
int zb_set_encode_sizes(const zb_set* s, size_t* n_kmer_words, size_t* n_count_words);
int zb_set_from_streams(int device, const uint64_t* kmer_words, size_t n_kmer_words,
                        const uint64_t* count_words, size_t n_count_words, zb_set** out);
/* the same from words that are already in `device` memory */
int zb_set_from_streams_dev(int device, const uint64_t* d_kmer_words, size_t n_kmer_words,
                            const uint64_t* d_count_words, size_t n_count_words, zb_set** out);
/* The same encode with the words LEFT IN HBM (codec64 kernels only): the result is fetched into host memory
 * (zb_words_fetch; pinned destinations copy at full PCIe rate) or written to a file by the library's I/O threads
 * (zb_words_write_fd: device -> pinned ring -> pwrite at the given file offsets, the two streams in parallel) --
 * the body of writeKmersAndCounts2 (files.py:209-217) + writeWords (files.py:65-83). */
int zb_set_encode_dev(const zb_set* s, zb_words** out);
int zb_words_sizes(const zb_words* w, size_t* n_kmer_words, size_t* n_count_words);
int zb_words_dev_ptrs(const zb_words* w, const uint64_t** d_kmer_words, const uint64_t** d_count_words);
int zb_words_fetch(const zb_words* w, uint64_t* kmer_words, uint64_t* count_words); /* either may be NULL */
int zb_words_write_fd(const zb_words* w, int fd, uint64_t kmers_offset, uint64_t counts_offset);
int zb_words_free(zb_words* w);
/* Range-partitioned encode: `s` is ONE RANGE of a longer sorted set whose ranges live on several GPUs and go to one
 * file.  codec64's greedy word boundaries (codec64.py:82-120) depend on everything before them, but only through
 * "how many values of this range the last word of the previous range has already taken" (0..5).  plan: prev_kmer =
 * last k-mer of the previous range (0 for the first; the delta base, files.py:85-98), next_* = the first n_next <= 5
 * entries behind this range (a word that starts here may run into them; fewer than 5 only at the end of the whole
 * set); *_map[r] = the state this range hands to the next one when entered in state r, *_map[6 + r] = the words it
 * then writes.  The caller chains the maps of the ranges in order (state 0 in front of the first) and calls emit with
 * each range's entry states; the concatenation of the emitted streams is the single-GPU stream, word for word.
 * emit consumes the plan. */
int zb_set_encode_plan(const zb_set* s, uint64_t prev_kmer, const uint64_t* next_kmers, const uint32_t* next_counts, int n_next,
                       zb_encplan** out, uint64_t kmer_map[12], uint64_t count_map[12]);
int zb_set_encode_emit(zb_encplan* p, int kmer_entry_state, int count_entry_state, zb_words** out);
int zb_encplan_free(zb_encplan* p);

/* ------------------------------------------------------------------------------------------
 * host I/O runtime         replaces zotmer/library/file.py:79-123 (openFile: the `gunzip -c` pipe / plain read feeding
 * the parser) and files.py:65-83 (writeWords) at the device boundary.
 * zb_stage_input: starts copying `raw` (any host memory: a mapping of the input file, a decompressed buffer, pinned
 * memory) to `device` and returns at once -- the library's I/O threads move it through a ring of pinned chunks
 * (memcpy + asynchronous H2D per chunk, several chunks in flight), so the copy of piece i + 1 overlaps the parsing /
 * extraction of piece i.  `raw` must stay valid until the piece has been fed or freed.
 * zb_kmerize_feed_staged: zb_kmerize_feed of a staged piece (waits on the device for its copy); consumes it.
 * zb_host_count_byte: occurrences of `byte` in host memory, on the I/O threads (cutting FASTQ text at record
 * boundaries -- every 4th newline -- for several GPUs or > 1 GiB pieces, library/reads.py:pieces).
 * ------------------------------------------------------------------------------------------ */
int zb_stage_input(int device, const uint8_t* raw, size_t n, zb_staged** out);
/* the same straight from a file: the I/O threads pread() [offset, offset + n) of `fd` into the pinned chunks (no
 * mapping, no page-table fill) */
int zb_stage_fd(int device, int fd, uint64_t offset, size_t n, zb_staged** out);
int zb_kmerize_feed_staged(zb_kmerizer* h, zb_staged* st, int is_fasta);
int zb_staged_free(zb_staged* st);
/* a k-mer set from its two word streams staged with zb_stage_fd / zb_stage_input (count_words may be NULL: all counts
 * 1) -- readKmersAndCounts, files.py:219-227, without a host copy of the streams; consumes both.  Staging the streams
 * of file i + 1 before this call for file i overlaps reading and copying with decoding (`zot merge`, `zot dist`). */
int zb_set_from_staged(zb_staged* kmer_words, zb_staged* count_words, zb_set** out);
int zb_host_count_byte(const uint8_t* p, size_t n, int byte, uint64_t* count);
/* ------------------------------------------------------------------------------------------
 * block-compressed input   replaces zotmer/library/file.py:93-97 (`gunzip -c` child + pipe) for files written by
 * bgzip (BGZF: independent gzip members of <= 64 KiB of text, each with its compressed size in a 'BC' extra field).
 * The compressed bytes are staged to the device and every member is inflated there by one warp.
 * zb_bgzf_probe:  ZB_OK iff raw[0, n) is a sequence of BGZF members; their number and the size of the text.
 * zb_stage_bgzf:  inflates whole members from the front of raw -- as many as fit max_out bytes of text together with
 *                 the carry (0 = the largest piece the parser takes) -- into a new staged piece on `device`;
 *                 *n_in_used = compressed bytes consumed.  carry_from / carry_off: the bytes [carry_off, len) of an
 *                 earlier piece (same device, not yet fed or freed) are copied in front of the inflated text -- the
 *                 incomplete last record of the previous group.  ZB_E_FORMAT: not BGZF, or a member does not inflate
 *                 to its recorded size (corrupt file).
 * zb_staged_cut:  where the text of a staged piece can be cut so that [0, cut) parses like a whole file (FASTQ: after
 *                 the last newline whose number is a multiple of 4; FASTA: in front of the last '>' that follows a
 *                 newline); 0 = nowhere.  library/reads.py:pieces on the device.
 * zb_staged_set_len shortens a piece (to its cut); zb_staged_fetch copies its first n bytes to the host (tests).
 * ------------------------------------------------------------------------------------------ */
int zb_bgzf_probe(const uint8_t* raw, size_t n, uint64_t* members, uint64_t* text_bytes);
int zb_stage_bgzf(int device, const uint8_t* raw, size_t n, uint64_t max_out, zb_staged* carry_from, uint64_t carry_off,
                  zb_staged** out, uint64_t* n_in_used);
int zb_staged_cut(zb_staged* st, int is_fasta, uint64_t* cut);
int zb_staged_len(const zb_staged* st, uint64_t* n);
int zb_staged_set_len(zb_staged* st, uint64_t n);
int zb_staged_fetch(zb_staged* st, uint8_t* host, size_t n);
/* Several devices inflate one BGZF file (library/devices.py): zb_bgzf_groups cuts the file into runs of whole members of
 * at most max_text bytes of text (starts[i] = compressed offset of group i; *n_groups may exceed cap: call again);
 * every device inflates its group with zb_stage_bgzf, then, in file order, puts the incomplete record the previous
 * group left behind in front of its text (zb_stage_concat: host prefix + a staged piece -> a new piece, consumes the
 * piece), finds its own cut (zb_staged_cut) and hands the bytes behind it on (zb_staged_fetch_range). */
int zb_bgzf_groups(const uint8_t* raw, size_t n, uint64_t max_text, uint64_t* starts, size_t cap, size_t* n_groups);
int zb_stage_concat(int device, const uint8_t* prefix, size_t prefix_len, zb_staged* body, zb_staged** out);
int zb_staged_fetch_range(zb_staged* st, uint64_t off, uint8_t* host, size_t n);

/* pinned host memory from the library's arena (cudaHostAlloc, cached): destinations of zb_words_fetch / zb_set_fetch */
int zb_host_alloc(size_t bytes, void** p);
/* the same, write-combined (cudaHostAllocWriteCombined): for buffers the CPU only WRITES and the device reads (input
 * text); PCIe reads of it are not snooped through the CPU caches.  Reading it from the CPU is very slow. */
int zb_host_alloc_wc(size_t bytes, void** p);
int zb_host_free(void* p);

/* ------------------------------------------------------------------------------------------
 * diagnostics: stage-level entry points for tests/ and bench.py (kernel isolation and timing).
 * Not part of the drop-in surface; host arrays in/out.
 * ------------------------------------------------------------------------------------------ */
/* sort keys (and optional u32 payload) in place; max_bits 8..11 = digit width cap; times `iters` sorts */
int zb_dbg_sort_u64(int device, uint64_t* keys, uint32_t* vals, size_t n, int key_bits, int max_bits, int iters,
                    float* ms_per_sort);
/* sort + run-length count (weights NULL = 1 each): distinct keys ascending + summed counts; out arrays hold n
 * entries; mode 0 = segmented path when profitable, 1 = full LSD sort + reduce-by-key, 2 = segmented path with the
 * promise that the keys are distinct (weights = payload; ZB_E_ARG when the promise is broken) */
int zb_dbg_sort_count(int device, const uint64_t* keys, const uint32_t* weights, size_t n, int key_bits, int mode,
                      int iters, uint64_t* out_k, uint32_t* out_c, size_t* n_out, float* ms_per_call);
/* text -> dense base codes (codes must hold n + 64 bytes) */
int zb_dbg_parse(int device, const uint8_t* raw, size_t n, int is_fasta, uint8_t* codes, size_t* n_codes,
                 uint64_t* n_records);
/* dense base codes -> canonical k-mers in unspecified order (keys must hold n entries) */
int zb_dbg_extract(int device, int k, const uint8_t* codes, size_t n, uint64_t* keys, size_t* n_keys);
/* per-stage device timing (CUDA events on the library stream): turn on/off; `report` receives
 * "stage total_ms calls" lines for everything run since it was switched on */
int zb_dbg_profile(int device, int on, char* report, size_t cap);
/* CUDA-event timer on the library stream: op 0 records the start, op 1 records the stop and returns ms */
int zb_dbg_timer(int device, int op, float* ms);
/* guard bands (ZB_GUARD=1 in the environment when the library is first used): every device block of the library's
 * allocator has 256 pattern bytes in front and behind; this scans the bands of all live and cached blocks of the
 * calling thread's context and returns in *n_bad how many blocks have a damaged band (an out-of-bounds store of some
 * kernel); *n_blocks = blocks scanned.  compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer.md). */
int zb_dbg_guard_check(int device, uint64_t* n_blocks, uint64_t* n_bad);

#ifdef __cplusplus
}
#endif
#endif
