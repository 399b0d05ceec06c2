// hostio.cu -- the host side of the device boundary: a small pool of I/O threads with a ring of pinned chunks.
//
// The reference feeds its parser from `open(fn)` / a `gunzip -c` pipe (zotmer/library/file.py:79-123) and writes its
// streams word by word (files.py:65-83).  Here the inputs are hundreds of megabytes that have to cross PCIe, and a
// plain cudaMemcpy from pageable memory (a mapping of the input file, a decompressed buffer) runs at ~11 GB/s -- one
// driver thread staging through one bounce buffer (profiles/r01_cli_end_to_end.md: 28 ms for 315 MB, 44 ms for the
// 203 MB result).  The pool does the same thing on several threads: every worker owns two pinned chunks and one
// stream per device;
//   H2D job  : memcpy (or pread) a chunk of the input into a pinned chunk, cudaMemcpyAsync it to its place in the
//              device buffer, record the chunk's event (the chunk is reused once that event has completed);
//   D2H job  : cudaMemcpyAsync a chunk of a device array into a pinned chunk, wait, pwrite it at its file offset;
//   count job: occurrences of a byte in a slice of host memory (record boundaries of FASTQ text).
// zb_stage_input returns at once; zb_kmerize_feed_staged makes the kmerizer's stream wait for the workers' streams.
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <thread>

#include "kernels.h"

namespace zb {

struct IoWorker {
    int id = 0;
    uint8_t* slot[2] = {nullptr, nullptr};
    cudaEvent_t ev[2][64];          // [slot][device], created on first use
    int pending_dev[2] = {-1, -1};  // device whose event guards the slot, -1 = free
    cudaStream_t stream[64];        // one per device, created by IoPool::ensure_device
    int turn = 0;
};

class IoPool {
  public:
    static IoPool& get() {
        static IoPool* p = new IoPool();   // never destroyed: worker threads may outlive static destruction order
        return *p;
    }
    int nthreads() const { return (int)workers_.size(); }
    size_t chunk() const { return chunk_; }

    // streams of every worker on `device` exist afterwards
    void ensure_device(int device) {
        std::lock_guard<std::mutex> lk(dev_mu_);
        if (dev_ready_[device & 63]) return;
        ZB_CUDA(cudaSetDevice(device));
        for (auto& w : workers_) ZB_CUDA(cudaStreamCreateWithFlags(&w->stream[device & 63], cudaStreamNonBlocking));
        dev_ready_[device & 63] = true;
    }
    cudaStream_t stream_of(int worker, int device) { return workers_[worker]->stream[device & 63]; }

    void submit(std::function<void(IoWorker&)> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(std::move(fn));
        }
        cv_.notify_one();
    }

  private:
    IoPool() {
        int nt = 8;
        if (const char* e = getenv("ZB_IO_THREADS")) nt = atoi(e);
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw && (unsigned)nt > hw) nt = (int)hw;
        if (nt < 1) nt = 1;
        if (nt > 32) nt = 32;
        size_t mb = 4;
        if (const char* e = getenv("ZB_IO_CHUNK_MB")) mb = (size_t)atoi(e);
        if (mb < 1) mb = 1;
        chunk_ = mb << 20;
        for (int i = 0; i < 64; i++) dev_ready_[i] = false;
        for (int i = 0; i < nt; i++) {
            IoWorker* w = new IoWorker();
            w->id = i;
            memset(w->ev, 0, sizeof w->ev);
            memset(w->stream, 0, sizeof w->stream);
            workers_.push_back(w);
        }
        for (auto* w : workers_) threads_.emplace_back([this, w] { run(w); });
        for (auto& t : threads_) t.detach();
    }

    void run(IoWorker* w) {
        while (true) {
            std::function<void(IoWorker&)> fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                fn = std::move(q_.front());
                q_.pop_front();
            }
            fn(*w);
        }
    }

    std::vector<IoWorker*> workers_;
    std::vector<std::thread> threads_;
    std::deque<std::function<void(IoWorker&)>> q_;
    std::mutex mu_, dev_mu_;
    std::condition_variable cv_;
    bool dev_ready_[64];
    size_t chunk_;
};

// a pinned chunk of the worker that is free to be overwritten (waits for the copy that last read it)
static uint8_t* take_slot(IoWorker& w, size_t chunk, int* which) {
    const int s = w.turn & 1;
    w.turn++;
    if (!w.slot[s]) {
        if (cudaHostAlloc((void**)&w.slot[s], chunk, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    }
    if (w.pending_dev[s] >= 0) {
        cudaSetDevice(w.pending_dev[s]);
        cudaEventSynchronize(w.ev[s][w.pending_dev[s] & 63]);
        w.pending_dev[s] = -1;
    }
    *which = s;
    return w.slot[s];
}

static bool guard_slot(IoWorker& w, int s, int device, cudaStream_t st) {
    cudaEvent_t& e = w.ev[s][device & 63];
    if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventRecord(e, st) != cudaSuccess) return false;
    w.pending_dev[s] = device;
    return true;
}

// completion of a group of jobs
struct IoGroup {
    std::mutex mu;
    std::condition_variable cv;
    size_t remaining = 0;
    uint32_t touched = 0;      // workers that enqueued device work for this group
    int error = 0;             // errno-like, 0 = ok
    void done(int worker, int err) {
        std::lock_guard<std::mutex> lk(mu);
        if (worker >= 0) touched |= 1u << worker;
        if (err && !error) error = err;
        if (--remaining == 0) cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return remaining == 0; });
    }
};

}  // namespace zb

using namespace zb;

struct zb_staged {
    Ctx* c;
    DBuf<uint8_t> d;
    size_t n = 0;
    IoGroup g;
    cudaEvent_t ev[32];
    cudaEvent_t alloc_ev = nullptr;   // the device buffer may be a recycled block still in use on the context's stream
};

// defined in api.cu: the body of zb_kmerize_feed_dev once the text is (or will be, in stream order) on the device
int zb_kmerize_feed_dev_ordered(zb_kmerizer* h, const uint8_t* d_raw, size_t n, int is_fasta);
zb::Ctx* zb_kmerizer_ctx(zb_kmerizer* h);

static void stage_jobs(zb_staged* st, const uint8_t* raw, int fd, uint64_t file_off, size_t n) {
    IoPool& pool = IoPool::get();
    const size_t chunk = pool.chunk();
    const size_t nchunks = div_up(n, chunk);
    const int device = st->c->device;
    pool.ensure_device(device);
    uint8_t* dst = st->d.get();
    ZB_CUDA(cudaEventCreateWithFlags(&st->alloc_ev, cudaEventDisableTiming));
    ZB_CUDA(cudaEventRecord(st->alloc_ev, st->c->stream));
    cudaEvent_t alloc_ev = st->alloc_ev;
    // a source that is pinned already (zb_host_alloc, cudaHostAlloc, cudaHostRegister) is copied from where it lies
    bool pinned = false;
    if (raw) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, raw) == cudaSuccess) pinned = pa.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    st->g.remaining = nchunks;   // nothing below throws: every counted chunk is submitted
    for (size_t i = 0; i < nchunks; i++) {
        const size_t off = i * chunk, len = std::min(chunk, n - off);
        if (pinned) {
            pool.submit([st, raw, dst, off, len, device, alloc_ev](IoWorker& w) {
                cudaSetDevice(device);
                cudaStream_t ws = w.stream[device & 63];
                const bool ok = cudaStreamWaitEvent(ws, alloc_ev, 0) == cudaSuccess &&
                                cudaMemcpyAsync(dst + off, raw + off, len, cudaMemcpyHostToDevice, ws) == cudaSuccess;
                st->g.done(w.id, ok ? 0 : EIO);
            });
            continue;
        }
        pool.submit([st, raw, fd, file_off, dst, off, len, device, chunk, alloc_ev](IoWorker& w) {
            int err = 0, s = 0;
            uint8_t* slot = take_slot(w, chunk, &s);
            if (!slot) { st->g.done(-1, ENOMEM); return; }
            if (raw) {
                memcpy(slot, raw + off, len);
            } else {
                size_t got = 0;
                while (got < len) {
                    const ssize_t r = pread(fd, slot + got, len - got, (off_t)(file_off + off + got));
                    if (r <= 0) { err = r < 0 ? errno : EIO; break; }
                    got += (size_t)r;
                }
            }
            if (!err) {
                cudaSetDevice(device);
                cudaStream_t ws = w.stream[device & 63];
                if (cudaStreamWaitEvent(ws, alloc_ev, 0) != cudaSuccess ||
                    cudaMemcpyAsync(dst + off, slot, len, cudaMemcpyHostToDevice, ws) != cudaSuccess || !guard_slot(w, s, device, ws)) err = EIO;
            }
            st->g.done(w.id, err);
        });
    }
}

extern "C" {

int zb_stage_input(int device, const uint8_t* raw, size_t n, zb_staged** out) {
    ZB_TRY
    if (!out || (n && !raw)) ZB_FAIL(ZB_E_ARG, "null argument");
    if (n >= ((size_t)1 << 31)) ZB_FAIL(ZB_E_ARG, "stage: piece of %zu bytes; split the input at record boundaries below 2 GiB", n);
    Ctx* c = ctx_for(device);
    zb_staged* st = new zb_staged();
    st->c = c;
    st->n = n;
    memset(st->ev, 0, sizeof st->ev);
    try {
        st->d.alloc(c, n + 16);
        if (n) stage_jobs(st, raw, -1, 0, n);
    } catch (...) {
        st->g.wait();
        delete st;
        throw;
    }
    *out = st;
    ZB_CATCH
}

int zb_stage_fd(int device, int fd, uint64_t offset, size_t n, zb_staged** out) {
    ZB_TRY
    if (!out || fd < 0) ZB_FAIL(ZB_E_ARG, "bad argument");
    if (n >= ((size_t)1 << 31)) ZB_FAIL(ZB_E_ARG, "stage: piece of %zu bytes; split the input at record boundaries below 2 GiB", n);
    Ctx* c = ctx_for(device);
    zb_staged* st = new zb_staged();
    st->c = c;
    st->n = n;
    memset(st->ev, 0, sizeof st->ev);
    try {
        st->d.alloc(c, n + 16);
        if (n) stage_jobs(st, nullptr, fd, offset, n);
    } catch (...) {
        st->g.wait();
        delete st;
        throw;
    }
    *out = st;
    ZB_CATCH
}

static void staged_release(zb_staged* st) {
    st->g.wait();
    cudaSetDevice(st->c->device);
    // the workers' copies may still be running: the buffer goes back to the allocator (and may be handed out again on
    // the context's stream) only after they have finished
    IoPool& pool = IoPool::get();
    for (int w = 0; w < pool.nthreads(); w++)
        if ((st->g.touched >> w) & 1u) cudaStreamSynchronize(pool.stream_of(w, st->c->device));
    for (int w = 0; w < 32; w++)
        if (st->ev[w]) cudaEventDestroy(st->ev[w]);
    if (st->alloc_ev) cudaEventDestroy(st->alloc_ev);
    delete st;
}

int zb_staged_free(zb_staged* st) {
    ZB_TRY
    if (st) staged_release(st);
    ZB_CATCH
}

int zb_kmerize_feed_staged(zb_kmerizer* h, zb_staged* st, int is_fasta) {
    int rc = ZB_OK;
    try {
        if (!h || !st) ZB_FAIL(ZB_E_ARG, "null argument");
        Ctx* c = zb_kmerizer_ctx(h);
        if (st->c != c) ZB_FAIL(ZB_E_ARG, "the staged piece belongs to another device context (stage and feed from one thread)");
        ZB_CUDA(cudaSetDevice(c->device));
        {
            Stage stg(c, "h2d_staged");
            st->g.wait();   // every chunk has been handed to a worker stream
            if (st->g.error) ZB_FAIL(ZB_E_CUDA, "staging the input failed (%s)", strerror(st->g.error));
            IoPool& pool = IoPool::get();
            for (int w = 0; w < pool.nthreads(); w++) {
                if (!((st->g.touched >> w) & 1u)) continue;
                if (!st->ev[w]) ZB_CUDA(cudaEventCreateWithFlags(&st->ev[w], cudaEventDisableTiming));
                ZB_CUDA(cudaEventRecord(st->ev[w], pool.stream_of(w, c->device)));
                ZB_CUDA(cudaStreamWaitEvent(c->stream, st->ev[w], 0));
            }
        }
        if (st->n) rc = zb_kmerize_feed_dev_ordered(h, st->d.get(), st->n, is_fasta);
    } catch (const zb::Fail& f) {
        rc = f.code;
    } catch (const std::bad_alloc&) {
        zb::set_error("out of host memory");
        rc = ZB_E_NOMEM;
    }
    if (st) {
        const std::string keep = zb_last_error();
        staged_release(st);
        if (rc != ZB_OK) zb::set_error("%s", keep.c_str());
    }
    return rc;
}

// the kmerizer-independent half of zb_kmerize_feed_staged: the context's stream waits for the piece's copies
static void staged_wait(zb_staged* st) {
    Ctx* c = st->c;
    ZB_CUDA(cudaSetDevice(c->device));
    Stage stg(c, "h2d_staged");
    st->g.wait();   // every chunk has been handed to a worker stream
    if (st->g.error) ZB_FAIL(ZB_E_CUDA, "staging the input failed (%s)", strerror(st->g.error));
    IoPool& pool = IoPool::get();
    for (int w = 0; w < pool.nthreads(); w++) {
        if (!((st->g.touched >> w) & 1u)) continue;
        if (!st->ev[w]) ZB_CUDA(cudaEventCreateWithFlags(&st->ev[w], cudaEventDisableTiming));
        ZB_CUDA(cudaEventRecord(st->ev[w], pool.stream_of(w, c->device)));
        ZB_CUDA(cudaStreamWaitEvent(c->stream, st->ev[w], 0));
    }
}

int zb_set_from_staged(zb_staged* kmer_words, zb_staged* count_words, zb_set** out) {
    int rc = ZB_OK;
    try {
        if (!kmer_words || !out) ZB_FAIL(ZB_E_ARG, "null argument");
        if (count_words && count_words->c != kmer_words->c) ZB_FAIL(ZB_E_ARG, "the two streams were staged on different contexts");
        if ((kmer_words->n & 7) || (count_words && (count_words->n & 7)))
            ZB_FAIL(ZB_E_FORMAT, "a word stream's length is not a multiple of 8 bytes");     // files.py:58 assert
        staged_wait(kmer_words);
        if (count_words) staged_wait(count_words);
        rc = zb_set_from_streams_dev(kmer_words->c->device, reinterpret_cast<const uint64_t*>(kmer_words->d.get()), kmer_words->n / 8,
                                     count_words ? reinterpret_cast<const uint64_t*>(count_words->d.get()) : nullptr,
                                     count_words ? count_words->n / 8 : 0, out);
    } catch (const zb::Fail& f) {
        rc = f.code;
    } catch (const std::bad_alloc&) {
        zb::set_error("out of host memory");
        rc = ZB_E_NOMEM;
    }
    const std::string keep = zb_last_error();
    if (kmer_words) staged_release(kmer_words);
    if (count_words) staged_release(count_words);
    if (rc != ZB_OK) zb::set_error("%s", keep.c_str());
    return rc;
}

// ---- block-compressed input (BGZF): members are found on the host, inflated on the device (inflate.cu) -------------
// Appends the members of raw[0, n) to `tab` until the text would exceed max_out (at least one member is taken).
// false: raw is not a sequence of BGZF members.
static bool bgzf_walk(const uint8_t* raw, size_t n, uint64_t max_out, std::vector<BgzfMember>* tab, size_t* used, uint64_t* text) {
    size_t p = 0;
    uint64_t out = 0;
    size_t taken = 0;
    while (p < n) {
        if (n - p < 28 || raw[p] != 0x1f || raw[p + 1] != 0x8b || raw[p + 2] != 8 || !(raw[p + 3] & 4)) return false;
        if (raw[p + 3] & ~4) return false;   // bgzip sets FEXTRA only (no name, comment or header CRC)
        const size_t xlen = (size_t)raw[p + 10] | ((size_t)raw[p + 11] << 8);
        if (p + 12 + xlen > n) return false;
        size_t q = p + 12, bsize = 0;
        const size_t xend = q + xlen;
        while (q + 4 <= xend) {
            const size_t slen = (size_t)raw[q + 2] | ((size_t)raw[q + 3] << 8);
            if (raw[q] == 'B' && raw[q + 1] == 'C' && slen == 2 && q + 6 <= xend) bsize = ((size_t)raw[q + 4] | ((size_t)raw[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < xlen + 20 + 2 || p + bsize > n) return false;
        const uint8_t* t = raw + p + bsize - 4;
        const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        if (isize > (1u << 16)) return false;
        if (taken && out + isize > max_out) break;
        if (tab) {
            BgzfMember m;
            m.src = p + 12 + xlen;
            m.dst = out;
            m.clen = (uint32_t)(bsize - xlen - 20);
            m.isize = isize;
            tab->push_back(m);
        }
        out += isize;
        taken++;
        p += bsize;
    }
    *used = p;
    *text = out;
    return true;
}

int zb_bgzf_probe(const uint8_t* raw, size_t n, uint64_t* members, uint64_t* text_bytes) {
    ZB_TRY
    if (!raw && n) ZB_FAIL(ZB_E_ARG, "null argument");
    std::vector<BgzfMember> tab;
    size_t used = 0;
    uint64_t text = 0;
    if (n == 0 || !bgzf_walk(raw, n, ~0ull, &tab, &used, &text)) ZB_FAIL(ZB_E_FORMAT, "not a BGZF file");
    if (members) *members = tab.size();
    if (text_bytes) *text_bytes = text;
    ZB_CATCH
}

int zb_stage_bgzf(int device, const uint8_t* raw, size_t n, uint64_t max_out, zb_staged* carry_from, uint64_t carry_off,
                  zb_staged** out, uint64_t* n_in_used) {
    zb_staged* comp = nullptr;
    zb_staged* res = nullptr;
    int rc = ZB_OK;
    try {
        if (!out || !raw || !n || !n_in_used) ZB_FAIL(ZB_E_ARG, "null argument");
        Ctx* c = ctx_for(device);
        if (carry_from && (carry_from->c != c || carry_off > carry_from->n)) ZB_FAIL(ZB_E_ARG, "bad carry");
        const uint64_t carry_len = carry_from ? carry_from->n - carry_off : 0;
        const uint64_t limit = ((uint64_t)1 << 31) - 4096;
        if (max_out == 0 || max_out > limit) max_out = limit;
        if (carry_len + (1u << 16) > max_out) ZB_FAIL(ZB_E_ARG, "a record of %llu bytes does not fit a piece", (unsigned long long)carry_len);
        std::vector<BgzfMember> tab;
        size_t used = 0;
        uint64_t text = 0;
        if (!bgzf_walk(raw, n, max_out - carry_len, &tab, &used, &text)) ZB_FAIL(ZB_E_FORMAT, "not a BGZF file");
        for (auto& m : tab) m.dst += carry_len;
        // the compressed bytes go through the pinned ring like any input
        comp = new zb_staged();
        comp->c = c;
        comp->n = used;
        memset(comp->ev, 0, sizeof comp->ev);
        comp->d.alloc(c, used + 64);
        stage_jobs(comp, raw, -1, 0, used);
        res = new zb_staged();
        res->c = c;
        res->n = carry_len + text;
        memset(res->ev, 0, sizeof res->ev);
        res->d.alloc(c, res->n + 64);
        DBuf<BgzfMember> d_tab(c, tab.size());
        DBuf<unsigned int> d_err(c, 4);
        ZB_CUDA(dev_memset(c, d_err.get(), 0, 16));
        ZB_CUDA(cudaMemcpyAsync(d_tab.get(), tab.data(), tab.size() * sizeof(BgzfMember), cudaMemcpyHostToDevice, c->stream));
        if (carry_len) ZB_CUDA(dev_copy(c, res->d.get(), carry_from->d.get() + carry_off, carry_len));
        staged_wait(comp);
        bgzf_inflate(c, comp->d.get(), d_tab.get(), (uint32_t)tab.size(), res->d.get(), d_err.get());
        ZB_CUDA(read_back(c, d_err.get(), 16));
        ZB_CUDA(cudaStreamSynchronize(c->stream));   // also: `tab` may go out of scope
        const uint32_t* e = reinterpret_cast<const uint32_t*>(c->h_scalars);
        if (e[0]) ZB_FAIL(ZB_E_FORMAT, "BGZF member %u of this group does not inflate (code %u; %u members failed)", e[1] - 1, e[2], e[0]);
        *n_in_used = used;
    } catch (const zb::Fail& f) {
        rc = f.code;
    } catch (const std::bad_alloc&) {
        zb::set_error("out of host memory");
        rc = ZB_E_NOMEM;
    }
    const std::string keep = rc != ZB_OK ? zb_last_error() : "";
    if (comp) staged_release(comp);
    if (rc != ZB_OK) {
        if (res) staged_release(res);
        zb::set_error("%s", keep.c_str());
        return rc;
    }
    *out = res;
    return ZB_OK;
}

int zb_bgzf_groups(const uint8_t* raw, size_t n, uint64_t max_text, uint64_t* starts, size_t cap, size_t* n_groups) {
    ZB_TRY
    if (!raw || !n || !n_groups || (cap && !starts)) ZB_FAIL(ZB_E_ARG, "null argument");
    if (max_text < (1u << 16)) max_text = 1u << 16;
    size_t off = 0, g = 0;
    while (off < n) {
        size_t used = 0;
        uint64_t text = 0;
        if (!bgzf_walk(raw + off, n - off, max_text, nullptr, &used, &text) || used == 0) ZB_FAIL(ZB_E_FORMAT, "not a BGZF file");
        if (g < cap) starts[g] = off;
        g++;
        off += used;
    }
    *n_groups = g;
    ZB_CATCH
}

// prefix (host bytes) + the text of `body` (same device; consumed) as one new piece: the incomplete record that the
// previous group of members left behind, in front of this group's text
int zb_stage_concat(int device, const uint8_t* prefix, size_t prefix_len, zb_staged* body, zb_staged** out) {
    int rc = ZB_OK;
    zb_staged* res = nullptr;
    try {
        if (!body || !out || (prefix_len && !prefix)) ZB_FAIL(ZB_E_ARG, "null argument");
        Ctx* c = ctx_for(device);
        if (body->c != c) ZB_FAIL(ZB_E_ARG, "the piece belongs to another device context");
        if (prefix_len + body->n >= ((size_t)1 << 31)) ZB_FAIL(ZB_E_ARG, "a piece of %zu bytes", prefix_len + body->n);
        staged_wait(body);
        res = new zb_staged();
        res->c = c;
        res->n = prefix_len + body->n;
        memset(res->ev, 0, sizeof res->ev);
        res->d.alloc(c, res->n + 64);
        if (prefix_len) ZB_CUDA(cudaMemcpyAsync(res->d.get(), prefix, prefix_len, cudaMemcpyHostToDevice, c->stream));
        if (body->n) ZB_CUDA(dev_copy(c, res->d.get() + prefix_len, body->d.get(), body->n));
        ZB_CUDA(cudaStreamSynchronize(c->stream));   // `prefix` may go out of scope
    } catch (const zb::Fail& f) {
        rc = f.code;
    } catch (const std::bad_alloc&) {
        zb::set_error("out of host memory");
        rc = ZB_E_NOMEM;
    }
    const std::string keep = rc != ZB_OK ? zb_last_error() : "";
    if (body) staged_release(body);
    if (rc != ZB_OK) {
        if (res) staged_release(res);
        zb::set_error("%s", keep.c_str());
        return rc;
    }
    *out = res;
    return ZB_OK;
}

int zb_staged_fetch_range(zb_staged* st, uint64_t off, uint8_t* host, size_t n) {
    ZB_TRY
    if (!st || (!host && n) || off > st->n || n > st->n - off) ZB_FAIL(ZB_E_ARG, "bad argument");
    staged_wait(st);
    if (n) ZB_CUDA(cudaMemcpyAsync(host, st->d.get() + off, n, cudaMemcpyDeviceToHost, st->c->stream));
    ZB_CUDA(cudaStreamSynchronize(st->c->stream));
    ZB_CATCH
}

int zb_staged_cut(zb_staged* st, int is_fasta, uint64_t* cut) {
    ZB_TRY
    if (!st || !cut) ZB_FAIL(ZB_E_ARG, "null argument");
    staged_wait(st);
    *cut = text_cut(st->c, st->d.get(), st->n, is_fasta != 0);
    ZB_CATCH
}

int zb_staged_len(const zb_staged* st, uint64_t* n) {
    ZB_TRY
    if (!st || !n) ZB_FAIL(ZB_E_ARG, "null argument");
    *n = st->n;
    ZB_CATCH
}

int zb_staged_set_len(zb_staged* st, uint64_t n) {
    ZB_TRY
    if (!st || n > st->n) ZB_FAIL(ZB_E_ARG, "bad length");
    st->n = (size_t)n;
    ZB_CATCH
}

int zb_staged_fetch(zb_staged* st, uint8_t* host, size_t n) {
    ZB_TRY
    if (!st || (!host && n) || n > st->n) ZB_FAIL(ZB_E_ARG, "bad argument");
    staged_wait(st);
    if (n) ZB_CUDA(cudaMemcpyAsync(host, st->d.get(), n, cudaMemcpyDeviceToHost, st->c->stream));
    ZB_CUDA(cudaStreamSynchronize(st->c->stream));
    ZB_CATCH
}

int zb_host_count_byte(const uint8_t* p, size_t n, int byte, uint64_t* count) {
    ZB_TRY
    if (!count || (n && !p)) ZB_FAIL(ZB_E_ARG, "null argument");
    *count = 0;
    if (n == 0) return ZB_OK;
    IoPool& pool = IoPool::get();
    const size_t parts = std::min<size_t>((size_t)pool.nthreads(), div_up(n, (size_t)1 << 20));
    std::vector<uint64_t> partial(parts, 0);
    IoGroup g;
    g.remaining = parts;
    const size_t per = div_up(n, parts);
    const uint8_t b = (uint8_t)byte;
    for (size_t i = 0; i < parts; i++) {
        const size_t lo = std::min(n, i * per), hi = std::min(n, lo + per);
        uint64_t* dst = &partial[i];
        pool.submit([p, lo, hi, b, dst, &g](IoWorker&) {
            uint64_t cnt = 0;
            // blocks of 255 x 16 bytes keep the per-byte-lane sums in 8 bits (the compiler vectorises the inner loop)
            for (size_t i0 = lo; i0 < hi; i0 += 4096) {
                const size_t i1 = std::min(hi, i0 + 4096);
                uint32_t c32 = 0;
                for (size_t j = i0; j < i1; j++) c32 += (p[j] == b);
                cnt += c32;
            }
            *dst = cnt;
            g.done(-1, 0);
        });
    }
    g.wait();
    uint64_t tot = 0;
    for (uint64_t v : partial) tot += v;
    *count = tot;
    ZB_CATCH
}

int zb_words_write_fd(const zb_words* w, int fd, uint64_t kmers_offset, uint64_t counts_offset) {
    ZB_TRY
    if (!w || fd < 0) ZB_FAIL(ZB_E_ARG, "bad argument");
    Ctx* c = w->c;
    ZB_CUDA(cudaSetDevice(c->device));
    ZB_CUDA(cudaStreamSynchronize(c->stream));   // the words are complete
    IoPool& pool = IoPool::get();
    const int device = c->device;
    pool.ensure_device(device);
    const size_t chunk = pool.chunk();
    struct Part { const uint8_t* src; size_t bytes; uint64_t off; };
    const Part parts[2] = {{reinterpret_cast<const uint8_t*>(w->kw.get()), w->nk * 8, kmers_offset},
                           {reinterpret_cast<const uint8_t*>(w->cw.get()), w->nc * 8, counts_offset}};
    IoGroup g;
    size_t njobs = 0;
    for (const Part& p : parts) njobs += div_up(p.bytes, chunk);
    if (njobs == 0) return ZB_OK;
    g.remaining = njobs;
    for (const Part& p : parts) {
        for (size_t o = 0; o < p.bytes; o += chunk) {
            const size_t len = std::min(chunk, p.bytes - o);
            const uint8_t* src = p.src + o;
            const uint64_t foff = p.off + o;
            pool.submit([src, len, foff, fd, device, chunk, &g](IoWorker& wk) {
                int err = 0, s = 0;
                uint8_t* slot = take_slot(wk, chunk, &s);
                if (!slot) { g.done(-1, ENOMEM); return; }
                cudaSetDevice(device);
                cudaStream_t ws = wk.stream[device & 63];
                if (cudaMemcpyAsync(slot, src, len, cudaMemcpyDeviceToHost, ws) != cudaSuccess ||
                    cudaStreamSynchronize(ws) != cudaSuccess) err = EIO;
                size_t put = 0;
                while (!err && put < len) {
                    const ssize_t r = pwrite(fd, slot + put, len - put, (off_t)(foff + put));
                    if (r <= 0) { err = r < 0 ? errno : EIO; break; }
                    put += (size_t)r;
                }
                g.done(-1, err);
            });
        }
    }
    g.wait();
    if (g.error) ZB_FAIL(ZB_E_CUDA, "writing the word streams failed (%s)", strerror(g.error));
    ZB_CATCH
}

// ---- pinned host memory (destinations of zb_words_fetch / zb_set_fetch): cudaHostAlloc is slow (~0.3 ms per MB), so
// released blocks are kept and handed out again
static std::mutex g_host_mu;
static std::multimap<size_t, void*> g_host_free[2];     // [plain, write-combined]
static std::map<void*, std::pair<size_t, int>> g_host_live;

static int host_alloc(size_t bytes, int wc, void** p) {
    ZB_TRY
    if (!p) ZB_FAIL(ZB_E_ARG, "null argument");
    ctx_for(0);   // fails with ZB_E_NOGPU on a machine without a device
    const size_t want = std::max<size_t>(4096, (bytes + 4095) & ~(size_t)4095);
    const unsigned flags = cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0u);
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto& fl = g_host_free[wc ? 1 : 0];
    auto it = fl.lower_bound(want);
    if (it != fl.end() && it->first <= 2 * want) {
        *p = it->second;
        g_host_live[*p] = {it->first, wc ? 1 : 0};
        fl.erase(it);
        return ZB_OK;
    }
    void* q = nullptr;
    if (cudaHostAlloc(&q, want, flags) != cudaSuccess) {
        cudaGetLastError();
        for (auto& f : g_host_free) {
            for (auto& kv : f) cudaFreeHost(kv.second);
            f.clear();
        }
        if (cudaHostAlloc(&q, want, flags) != cudaSuccess) {
            cudaGetLastError();
            ZB_FAIL(ZB_E_NOMEM, "out of pinned host memory: %zu bytes requested", want);
        }
    }
    g_host_live[q] = {want, wc ? 1 : 0};
    *p = q;
    ZB_CATCH
}

int zb_host_alloc(size_t bytes, void** p) { return host_alloc(bytes, 0, p); }
int zb_host_alloc_wc(size_t bytes, void** p) { return host_alloc(bytes, 1, p); }

int zb_host_free(void* p) {
    ZB_TRY
    if (!p) return ZB_OK;
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_live.find(p);
    if (it == g_host_live.end()) ZB_FAIL(ZB_E_ARG, "not a pointer of zb_host_alloc");
    g_host_free[it->second.second].insert({it->second.first, p});
    g_host_live.erase(it);
    ZB_CATCH
}

}  // extern "C"
