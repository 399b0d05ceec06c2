"""
ctypes binding of libzot_b200.so (include/zotmer_b200.h) -- the only door between the Python host
layer (zotmer_b200/commands, zotmer_b200/library) and the sm_100a kernels.

There is NO CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, the call raises.  The library is built in-tree by
`python -c "import __graft_entry__ as g; g.build()"` (or `make -C zotmer_b200/csrc`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzot_b200.so")

ZB_OK = 0
ZB_E_CUDA, ZB_E_ARG, ZB_E_RANGE, ZB_E_FORMAT, ZB_E_NOMEM, ZB_E_NOGPU = -1, -2, -3, -4, -5, -6

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/zotmer_b200.h
SIGNATURES = {
    "zb_last_error": (C.c_char_p, []),
    "zb_version": (C.c_int, []),
    "zb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "zb_launch_count": (C.c_int, [C.c_int, u64p]),
    "zb_device_sync": (C.c_int, [C.c_int]),
    "zb_release_cache": (C.c_int, [C.c_int]),
    "zb_dev_read_small": (C.c_int, [C.c_int, vp, C.c_size_t, vp]),
    "zb_dev_write_small": (C.c_int, [C.c_int, vp, vp, C.c_size_t]),
    "zb_kmerize_open": (C.c_int, [C.c_int, C.c_int, C.POINTER(vp)]),
    "zb_kmerize_feed": (C.c_int, [vp, vp, C.c_size_t, C.c_int]),
    "zb_kmerize_feed_dev": (C.c_int, [vp, vp, C.c_size_t, C.c_int]),
    "zb_kmerize_set_baits": (C.c_int, [vp, vp]),
    "zb_kmerize_feed_codes_dev": (C.c_int, [vp, vp, C.c_size_t, C.c_uint64]),
    "zb_kmerize_finish": (C.c_int, [vp, C.POINTER(vp), u64p]),
    "zb_kmerize_close": (C.c_int, [vp]),
    "zb_kmerize_set_owners": (C.c_int, [vp, C.c_int]),
    "zb_kmerize_pending": (C.c_int, [vp, u64p]),
    "zb_kmerize_take_bucketed_dev": (C.c_int, [vp, C.c_int, vp, u64p]),
    "zb_kmerize_bucket_counts": (C.c_int, [vp, C.c_int, u64p]),
    "zb_kmerize_route_p2p": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "zb_kmerize_route_p2p_begin": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "zb_kmerize_route_p2p_end": (C.c_int, [vp]),
    "zb_kmerize_route_p2p_reserve": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_uint64, C.POINTER(C.c_uint64)]),
    "zb_peer_enable": (C.c_int, [C.c_int, C.c_int]),
    "zb_ipc_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(vp), C.c_char_p]),
    "zb_ipc_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(vp)]),
    "zb_ipc_close": (C.c_int, [C.c_int, vp]),
    "zb_ipc_free": (C.c_int, [C.c_int, vp]),
    "zb_kmerize_adopt_canonical_dev": (C.c_int, [vp, vp, C.c_size_t]),
    "zb_kmerize_add_canonical_dev": (C.c_int, [vp, vp, C.c_size_t]),
    "zb_kmerize_flush": (C.c_int, [vp]),
    "zb_set_from_host": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.POINTER(vp)]),
    "zb_set_from_device": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.POINTER(vp)]),
    "zb_set_lower_bound": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "zb_set_slice": (C.c_int, [vp, C.c_size_t, C.c_size_t, C.POINTER(vp)]),
    "zb_set_size": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "zb_set_fetch": (C.c_int, [vp, vp, vp]),
    "zb_set_dev_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "zb_set_free": (C.c_int, [vp]),
    "zb_set_stats": (C.c_int, [vp, u64p, u64p, u64p, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "zb_merge": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "zb_set_is_wide": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "zb_set_fetch_counts64": (C.c_int, [vp, vp]),
    "zb_trim": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "zb_sample": (C.c_int, [vp, C.c_int, C.c_uint64, C.c_double, C.POINTER(vp)]),
    "zb_restrict": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "zb_project": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "zb_pairs_abc": (C.c_int, [C.c_int, C.POINTER(vp), vp, vp, C.c_size_t, vp]),
    "zb_allpairs_tiles": (C.c_int, [C.c_int, u64p]),
    "zb_allpairs_abc": (C.c_int, [C.c_int, C.POINTER(vp), C.c_uint64, C.c_uint64, vp]),
    "zb_allpairs_abc_strided": (C.c_int, [C.c_int, C.POINTER(vp), C.c_uint64, C.c_uint64, C.c_uint64, vp]),
    "zb_encode_u64_stream": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_int, vp, C.POINTER(C.c_size_t)]),
    "zb_decode_u64_stream": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_int, vp, C.POINTER(C.c_size_t)]),
    "zb_set_encode": (C.c_int, [vp, vp, C.POINTER(C.c_size_t), vp, C.POINTER(C.c_size_t)]),
    "zb_set_encode_sizes": (C.c_int, [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "zb_set_from_streams": (C.c_int, [C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]),
    "zb_set_encode_dev": (C.c_int, [vp, C.POINTER(vp)]),
    "zb_words_sizes": (C.c_int, [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "zb_words_dev_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "zb_words_fetch": (C.c_int, [vp, vp, vp]),
    "zb_words_write_fd": (C.c_int, [vp, C.c_int, C.c_uint64, C.c_uint64]),
    "zb_words_free": (C.c_int, [vp]),
    "zb_set_encode_plan": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, C.POINTER(vp), u64p, u64p]),
    "zb_set_encode_emit": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "zb_encplan_free": (C.c_int, [vp]),
    "zb_stage_input": (C.c_int, [C.c_int, vp, C.c_size_t, C.POINTER(vp)]),
    "zb_stage_fd": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_size_t, C.POINTER(vp)]),
    "zb_kmerize_feed_staged": (C.c_int, [vp, vp, C.c_int]),
    "zb_staged_free": (C.c_int, [vp]),
    "zb_bgzf_probe": (C.c_int, [vp, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zb_stage_bgzf": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_uint64, vp, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "zb_staged_cut": (C.c_int, [vp, C.c_int, C.POINTER(C.c_uint64)]),
    "zb_staged_len": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "zb_staged_set_len": (C.c_int, [vp, C.c_uint64]),
    "zb_staged_fetch": (C.c_int, [vp, vp, C.c_size_t]),
    "zb_bgzf_groups": (C.c_int, [vp, C.c_size_t, C.c_uint64, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "zb_stage_concat": (C.c_int, [C.c_int, vp, C.c_size_t, vp, C.POINTER(vp)]),
    "zb_staged_fetch_range": (C.c_int, [vp, C.c_uint64, vp, C.c_size_t]),
    "zb_set_from_staged": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "zb_set_from_streams_dev": (C.c_int, [C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]),
    "zb_host_count_byte": (C.c_int, [vp, C.c_size_t, C.c_int, u64p]),
    "zb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "zb_host_alloc_wc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "zb_host_free": (C.c_int, [vp]),
    "zb_dbg_guard_check": (C.c_int, [C.c_int, u64p, u64p]),
    "zb_dbg_sort_u64": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "zb_dbg_sort_count": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_float)]),
    "zb_dbg_parse": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_int, vp, C.POINTER(C.c_size_t), u64p]),
    "zb_dbg_extract": (C.c_int, [C.c_int, C.c_int, vp, C.c_size_t, vp, C.POINTER(C.c_size_t)]),
    "zb_dbg_profile": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "zb_dbg_timer": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libzot_b200: %s (code %d)" % (msg, code))
        self.code = code


def lib():
    """Load libzot_b200.so; fails loudly when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `make -C zotmer_b200/csrc` "
                              "(zotmer_b200 has no CPU path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for nm, (res, args) in SIGNATURES.items():
            fn = getattr(L, nm)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != ZB_OK:
        msg = lib().zb_last_error().decode("utf-8", "replace")
        if rc == ZB_E_RANGE:
            # the reference raises IndexError (codec64.py:93-99) / OverflowError here
            raise IndexError(msg)
        if rc == ZB_E_FORMAT:
            raise AssertionError(msg)  # files.py:58,182 asserts
        raise NativeError(rc, msg)


def device_count():
    n = C.c_int(0)
    _check(lib().zb_device_count(C.byref(n)))
    return n.value


def launch_count(device=0):
    n = C.c_uint64(0)
    _check(lib().zb_launch_count(device, C.byref(n)))
    return n.value


def dev_read_small(dptr, n, dtype=np.int64, device=0):
    """n elements of device memory -> numpy array, by a kernel through mapped pinned memory (no copy engine)"""
    out = np.empty(n, dtype)
    _check(lib().zb_dev_read_small(device, vp(int(dptr)), out.nbytes, _ptr(out)))
    return out


def dev_write_small(dptr, arr, device=0):
    a = np.ascontiguousarray(arr)
    _check(lib().zb_dev_write_small(device, vp(int(dptr)), _ptr(a), a.nbytes))


def release_cache(device=0):
    """give the device memory cached by the library's allocator back to the driver"""
    _check(lib().zb_release_cache(device))


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None and len(a) else None


class KmerSet(object):
    """Device-resident counted k-mer set (sorted u64 k-mers + u32 counts)."""

    def __init__(self, handle, device):
        self.h = vp(handle) if not isinstance(handle, vp) else handle
        self.device = device

    @staticmethod
    def from_arrays(kmers, counts=None, device=0):
        k = np.ascontiguousarray(kmers, dtype=np.uint64)
        c = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
        if c is not None and len(c) != len(k):
            raise AssertionError("k-mer and count arrays differ in length")  # files.py:182
        h = vp()
        _check(lib().zb_set_from_host(device, _ptr(k), _ptr(c) if c is not None else None, len(k), C.byref(h)))
        return KmerSet(h, device)

    @staticmethod
    def from_device(d_kmers, d_counts, n, device=0):
        """copy of device arrays (sorted, duplicate-free k-mers + counts) into a new set"""
        h = vp()
        _check(lib().zb_set_from_device(device, vp(d_kmers), vp(d_counts) if d_counts else None, n, C.byref(h)))
        return KmerSet(h, device)

    @staticmethod
    def from_streams(kmer_words, count_words=None, device=0):
        kw = np.ascontiguousarray(kmer_words, dtype=np.uint64)
        cw = None if count_words is None else np.ascontiguousarray(count_words, dtype=np.uint64)
        h = vp()
        _check(lib().zb_set_from_streams(device, _ptr(kw), len(kw), _ptr(cw) if cw is not None else None,
                                         0 if cw is None else len(cw), C.byref(h)))
        return KmerSet(h, device)

    @staticmethod
    def from_staged(kmer_words, count_words=None, device=0):
        """the set whose two word streams were staged with stage_fd / stage_input (consumes them)"""
        h = vp()
        kh, kmer_words.h = kmer_words.h, None
        ch = None
        if count_words is not None:
            ch, count_words.h = count_words.h, None
        try:
            _check(lib().zb_set_from_staged(kh, ch, C.byref(h)))
        finally:
            kmer_words.keep = None
            if count_words is not None:
                count_words.keep = None
        return KmerSet(h, device)

    def __len__(self):
        n = C.c_size_t(0)
        _check(lib().zb_set_size(self.h, C.byref(n)))
        return n.value

    def fetch(self, counts=True, out_k=None, out_c=None):
        """copy to host; out_k/out_c may be preallocated (e.g. pinned) arrays of sufficient size"""
        n = len(self)
        k = np.empty(n, np.uint64) if out_k is None else out_k[:n]
        c = None
        if counts and self.is_wide():
            # a merge whose sums passed 2^32-1 (merge.py:145-146 adds Python ints): the counts come back as u64
            _check(lib().zb_set_fetch(self.h, _ptr(k), None))
            c = np.empty(n, np.uint64)
            _check(lib().zb_set_fetch_counts64(self.h, _ptr(c)))
            return k, c
        if counts:
            c = np.empty(n, np.uint32) if out_c is None else out_c[:n]
        _check(lib().zb_set_fetch(self.h, _ptr(k), _ptr(c) if counts else None))
        return (k, c) if counts else k

    def is_wide(self):
        """True when some count exceeds 2^32-1 (only a merge produces such a set)"""
        w = C.c_int(0)
        _check(lib().zb_set_is_wide(self.h, C.byref(w)))
        return bool(w.value)

    def dev_ptrs(self):
        a, b = vp(), vp()
        _check(lib().zb_set_dev_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stats(self):
        """-> dict(acgt_weighted, acgt_plain, total, hist=[(count, n_kmers)] in first-occurrence order)"""
        aw = (C.c_uint64 * 4)()
        ap = (C.c_uint64 * 4)()
        tot = C.c_uint64(0)
        nh = C.c_size_t(0)
        cap = 4096
        while True:
            hv = np.empty(cap, np.uint64)
            hf = np.empty(cap, np.uint64)
            _check(lib().zb_set_stats(self.h, aw, ap, C.byref(tot), _ptr(hv), _ptr(hf), cap, C.byref(nh)))
            if nh.value <= cap:
                break
            cap = nh.value
        return {"acgt_weighted": [int(x) for x in aw], "acgt_plain": [int(x) for x in ap], "total": int(tot.value),
                "hist": [(int(hv[i]), int(hf[i])) for i in range(nh.value)]}

    def trim(self, cmin, cmax=0):
        h = vp()
        _check(lib().zb_trim(self.h, int(cmin), int(cmax), C.byref(h)))
        return KmerSet(h, self.device)

    def lower_bound(self, probes):
        """first position whose k-mer is >= each probe -> uint64 array"""
        p = np.ascontiguousarray(probes, dtype=np.uint64)
        out = np.zeros(len(p), np.uint64)
        _check(lib().zb_set_lower_bound(self.h, _ptr(p), len(p), _ptr(out)))
        return out

    def slice(self, begin, end):
        """copy of the entries [begin, end)"""
        h = vp()
        _check(lib().zb_set_slice(self.h, int(begin), int(end), C.byref(h)))
        return KmerSet(h, self.device)

    def sample(self, p, seed=0, mode=0):
        """hash sub-sampling: mode 0 = `zot sample`, mode 1 = `zot kmerize -D` (basics.sub)"""
        h = vp()
        _check(lib().zb_sample(self.h, int(mode), int(seed) & 0xFFFFFFFFFFFFFFFF, float(p), C.byref(h)))
        return KmerSet(h, self.device)

    def restrict(self, ref):
        """the entries whose k-mer occurs in `ref` (`zot project`)"""
        h = vp()
        _check(lib().zb_restrict(self.h, ref.h, C.byref(h)))
        return KmerSet(h, self.device)

    def project(self, shift_bits):
        h = vp()
        _check(lib().zb_project(self.h, int(shift_bits), C.byref(h)))
        return KmerSet(h, self.device)

    def encode_dev(self):
        """the two packed file streams of the set, left on the device -> Words"""
        h = vp()
        _check(lib().zb_set_encode_dev(self.h, C.byref(h)))
        return Words(h, self.device)

    def encode_plan(self, prev_kmer=0, next_kmers=(), next_counts=()):
        """first phase of a range-partitioned encode (this set = one range of a longer sorted set):
        -> (EncodePlan, kmer_map, count_map); map = ([exit state per entry state], [words per entry state])"""
        nk = np.ascontiguousarray(next_kmers, dtype=np.uint64)
        nc = np.ascontiguousarray(next_counts, dtype=np.uint32)
        assert len(nk) == len(nc) <= 5
        h = vp()
        km = (C.c_uint64 * 12)()
        cm = (C.c_uint64 * 12)()
        _check(lib().zb_set_encode_plan(self.h, int(prev_kmer), _ptr(nk), _ptr(nc), len(nk), C.byref(h), km, cm))
        return (EncodePlan(h, self), ([int(x) for x in km[:6]], [int(x) for x in km[6:]]),
                ([int(x) for x in cm[:6]], [int(x) for x in cm[6:]]))

    def encode(self):
        """-> (kmer stream words, count stream words) exactly as the reference writes them."""
        nk, nc = C.c_size_t(0), C.c_size_t(0)
        _check(lib().zb_set_encode_sizes(self.h, C.byref(nk), C.byref(nc)))
        kw = np.empty(max(nk.value, 1), np.uint64)
        cw = np.empty(max(nc.value, 1), np.uint64)
        _check(lib().zb_set_encode(self.h, _ptr(kw), C.byref(nk), _ptr(cw), C.byref(nc)))
        return kw[:nk.value], cw[:nc.value]

    def free(self):
        if self.h is not None and self.h.value:
            lib().zb_set_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Words(object):
    """The packed codec64 streams ('kmers' delta-coded, 'counts') of a set, resident on the device."""

    def __init__(self, handle, device):
        self.h = handle
        self.device = device

    def sizes(self):
        nk, nc = C.c_size_t(0), C.c_size_t(0)
        _check(lib().zb_words_sizes(self.h, C.byref(nk), C.byref(nc)))
        return nk.value, nc.value

    def dev_ptrs(self):
        a, b = vp(), vp()
        _check(lib().zb_words_dev_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fetch(self, out_k=None, out_c=None):
        """-> (kmer words, count words) on the host; out_k / out_c: preallocated (e.g. pinned) uint64 arrays"""
        nk, nc = self.sizes()
        kw = np.empty(nk, np.uint64) if out_k is None else out_k[:nk]
        cw = np.empty(nc, np.uint64) if out_c is None else out_c[:nc]
        _check(lib().zb_words_fetch(self.h, _ptr(kw), _ptr(cw)))
        return kw, cw

    def write_fd(self, fd, kmers_offset, counts_offset):
        """device -> pinned ring -> pwrite(fd) at the two offsets, on the library's I/O threads"""
        _check(lib().zb_words_write_fd(self.h, int(fd), int(kmers_offset), int(counts_offset)))

    def free(self):
        if self.h is not None and self.h.value:
            lib().zb_words_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class EncodePlan(object):
    def __init__(self, handle, kset):
        self.h = handle
        self.kset = kset      # the plan reads the set again in emit

    def emit(self, kmer_entry_state, count_entry_state):
        """second phase -> Words; consumes the plan"""
        h = vp()
        ph, self.h = self.h, None
        _check(lib().zb_set_encode_emit(ph, int(kmer_entry_state), int(count_entry_state), C.byref(h)))
        return Words(h, self.kset.device)

    def __del__(self):
        try:
            if self.h is not None and self.h.value:
                lib().zb_encplan_free(self.h)
                self.h = None
        except Exception:
            pass


def chain_ranges(maps):
    """maps: per range (exit[6], words[6]) in order -> (entry state per range, word offset per range, total words)
    (zb_set_encode_plan: state 0 in front of the first range)"""
    st, off = 0, 0
    entry, offs = [], []
    for (ex, wd) in maps:
        entry.append(st)
        offs.append(off)
        off += wd[st]
        st = ex[st]
    return entry, offs, off


class Staged(object):
    """input text on its way to the device (zb_stage_input / zb_stage_fd)"""

    def __init__(self, handle, keep=None):
        self.h = handle
        self.keep = keep      # the host buffer must outlive the copy

    def free(self):
        if self.h is not None and self.h.value:
            lib().zb_staged_free(self.h)
            self.h = None
        self.keep = None

    def __len__(self):
        n = C.c_uint64(0)
        _check(lib().zb_staged_len(self.h, C.byref(n)))
        return n.value

    def cut(self, is_fasta):
        """offset at which the text can be cut at a record boundary (0: nowhere) -- reads.pieces on the device"""
        n = C.c_uint64(0)
        _check(lib().zb_staged_cut(self.h, 1 if is_fasta else 0, C.byref(n)))
        return n.value

    def set_len(self, n):
        _check(lib().zb_staged_set_len(self.h, int(n)))

    def fetch(self, n=None):
        """the first n bytes of the piece as it sits on the device (tests)"""
        n = len(self) if n is None else int(n)
        out = np.empty(n, dtype=np.uint8)
        _check(lib().zb_staged_fetch(self.h, _ptr(out), n))
        return out.tobytes()

    def fetch_range(self, off, n):
        """bytes [off, off + n) of the piece (the incomplete record behind a cut)"""
        out = np.empty(int(n), dtype=np.uint8)
        _check(lib().zb_staged_fetch_range(self.h, int(off), _ptr(out), int(n)))
        return out.tobytes()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def stage_input(data, device=0):
    """start the asynchronous copy of `data` (bytes / memoryview / mmap / uint8 array) to the device -> Staged"""
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    h = vp()
    _check(lib().zb_stage_input(device, _ptr(a), len(a), C.byref(h)))
    return Staged(h, (a, data))


def bgzf_probe(data):
    """(members, text bytes) when `data` is a BGZF file (bgzip output: gzip members that carry their own size), else
    None -- the members are found from their headers, nothing is inflated"""
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    if len(a) < 28:
        return None
    m, t = C.c_uint64(0), C.c_uint64(0)
    if lib().zb_bgzf_probe(_ptr(a), len(a), C.byref(m), C.byref(t)) != 0:
        return None
    return m.value, t.value


def stage_bgzf(data, device=0, max_out=0, carry=None, carry_off=0):
    """inflate whole BGZF members from the front of `data` ON THE DEVICE into a new staged piece (at most max_out bytes of
    text, the carried tail of the piece `carry` in front) -> (Staged, compressed bytes consumed)"""
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    h, used = vp(), C.c_uint64(0)
    _check(lib().zb_stage_bgzf(device, _ptr(a), len(a), int(max_out), carry.h if carry is not None else None,
                               int(carry_off), C.byref(h), C.byref(used)))
    return Staged(h), used.value


def bgzf_groups(data, max_text):
    """compressed offsets at which runs of whole BGZF members of at most max_text bytes of text start"""
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    cap = 1024
    while True:
        starts = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t(0)
        _check(lib().zb_bgzf_groups(_ptr(a), len(a), int(max_text), _ptr(starts), cap, C.byref(n)))
        if n.value <= cap:
            return [int(x) for x in starts[:n.value]]
        cap = n.value


def stage_concat(prefix, body, device=0):
    """host bytes `prefix` + the staged piece `body` (consumed) -> a new Staged on the same device"""
    a = np.frombuffer(prefix, dtype=np.uint8) if len(prefix) else None
    h = vp()
    bh, body.h = body.h, None
    _check(lib().zb_stage_concat(device, _ptr(a), len(prefix), bh, C.byref(h)))
    return Staged(h)


def stage_fd(fd, offset, n, device=0):
    """the same straight from a file descriptor (pread on the I/O threads)"""
    h = vp()
    _check(lib().zb_stage_fd(device, int(fd), int(offset), int(n), C.byref(h)))
    return Staged(h)


def host_count_byte(data, byte):
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    n = C.c_uint64(0)
    _check(lib().zb_host_count_byte(_ptr(a), len(a), int(byte), C.byref(n)))
    return n.value


class PinnedArray(object):
    """a numpy array over pinned host memory of the library's arena (zb_host_alloc); .a is the array"""

    def __init__(self, n, dtype, write_combined=False):
        dt = np.dtype(dtype)
        self.p = vp()
        self.nbytes = max(int(n), 1) * dt.itemsize
        fn = lib().zb_host_alloc_wc if write_combined else lib().zb_host_alloc
        _check(fn(self.nbytes, C.byref(self.p)))
        buf = (C.c_uint8 * self.nbytes).from_address(self.p.value)
        self.a = np.frombuffer(buf, dtype=dt, count=max(int(n), 1))

    def free(self):
        if self.p is not None and self.p.value:
            self.a = None
            lib().zb_host_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def guard_check(device=0):
    """(blocks scanned, blocks with a damaged guard band); needs ZB_GUARD=1 in the environment"""
    nb, bad = C.c_uint64(0), C.c_uint64(0)
    _check(lib().zb_dbg_guard_check(device, C.byref(nb), C.byref(bad)))
    return nb.value, bad.value


def merge(sets):
    arr = (vp * len(sets))(*[s.h for s in sets])
    h = vp()
    _check(lib().zb_merge(len(sets), arr, C.byref(h)))
    return KmerSet(h, sets[0].device)


def pairs_abc(sets, I, J):
    """(|X n Y|, |X \\ Y|, |Y \\ X|) for each pair (I[p], J[p]) -> uint64 array [npairs, 3]"""
    I = np.ascontiguousarray(I, dtype=np.uint32)
    J = np.ascontiguousarray(J, dtype=np.uint32)
    assert len(I) == len(J)
    arr = (vp * len(sets))(*[s.h for s in sets])
    out = np.zeros((len(I), 3), np.uint64)
    _check(lib().zb_pairs_abc(len(sets), arr, _ptr(I), _ptr(J), len(I), _ptr(out.reshape(-1)) if len(I) else None))
    return out


def peer_enable(device, peer):
    """kernels on `device` may access memory of `peer` afterwards (one process, several GPUs)"""
    _check(lib().zb_peer_enable(device, peer))


def ipc_alloc(nbytes, device=0):
    """-> (device pointer, 64-byte handle that another process on the node can open)"""
    p = vp()
    h = C.create_string_buffer(64)
    _check(lib().zb_ipc_alloc(device, nbytes, C.byref(p), h))
    return p.value, h.raw


def ipc_open(handle, device=0):
    p = vp()
    _check(lib().zb_ipc_open(device, handle, C.byref(p)))
    return p.value


def ipc_close(ptr, device=0):
    _check(lib().zb_ipc_close(device, vp(ptr)))


def ipc_free(ptr, device=0):
    _check(lib().zb_ipc_free(device, vp(ptr)))


def allpairs_tiles(nsets):
    n = C.c_uint64(0)
    _check(lib().zb_allpairs_tiles(nsets, C.byref(n)))
    return n.value


def allpairs_abc(sets, tile_begin=0, tile_end=0, stride=1):
    """(|X n Y|, |X \\ Y|, |Y \\ X|) for all pairs i < j in row-major order -> uint64 array [n (n - 1) / 2, 3];
    with a unit range only the pairs of those units are filled (restricted to the units' key-range shards), the others
    are 0 (multi-GPU shards add up); stride > 1 takes every stride-th unit of the range (rank r of W: (r, 0, W))"""
    n = len(sets)
    out = np.zeros((n * (n - 1) // 2, 3), np.uint64)
    if n < 2:
        return out
    arr = (vp * n)(*[s.h for s in sets])
    _check(lib().zb_allpairs_abc_strided(n, arr, tile_begin, tile_end, stride, _ptr(out.reshape(-1))))
    return out


class Kmerizer(object):
    """Streaming kmerize+count (zb_kmerize_*)."""

    def __init__(self, k, device=0):
        self.h = vp()
        self.device = device
        self.k = k
        _check(lib().zb_kmerize_open(int(k), device, C.byref(self.h)))

    def feed(self, data, is_fasta):
        """data: bytes / bytearray / memoryview / uint8 ndarray holding a whole file or a record-aligned piece."""
        a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        if len(a):
            _check(lib().zb_kmerize_feed(self.h, _ptr(a), len(a), 1 if is_fasta else 0))

    def feed_staged(self, staged, is_fasta):
        """feed a piece started with stage_input / stage_fd (consumes it)"""
        h, staged.h = staged.h, None
        try:
            _check(lib().zb_kmerize_feed_staged(self.h, h, 1 if is_fasta else 0))
        finally:
            staged.keep = None

    def feed_dev(self, dptr, n, is_fasta):
        _check(lib().zb_kmerize_feed_dev(self.h, vp(dptr), n, 1 if is_fasta else 0))

    def set_baits(self, baits):
        """capture mode (`zot kmerize -C`): records fed from now on count only if they hold a k-mer of the KmerSet
        `baits` (which must stay alive while this kmerizer is fed); None switches it off"""
        self._baits = baits
        _check(lib().zb_kmerize_set_baits(self.h, baits.h if baits is not None else None))

    def feed_codes_dev(self, dptr, n, n_records):
        _check(lib().zb_kmerize_feed_codes_dev(self.h, vp(dptr), n, n_records))

    def set_owners(self, nranks):
        """the keys extracted from now on are tallied per owner among `nranks` GPUs (bucket_counts then costs no pass)"""
        _check(lib().zb_kmerize_set_owners(self.h, int(nranks)))

    def pending(self):
        n = C.c_uint64(0)
        _check(lib().zb_kmerize_pending(self.h, C.byref(n)))
        return n.value

    def take_bucketed_dev(self, nranks, dptr):
        cnt = (C.c_uint64 * nranks)()
        _check(lib().zb_kmerize_take_bucketed_dev(self.h, nranks, vp(dptr), cnt))
        return [int(x) for x in cnt]

    def bucket_counts(self, nranks):
        counts = (C.c_uint64 * nranks)()
        _check(lib().zb_kmerize_bucket_counts(self.h, nranks, counts))
        return [int(x) for x in counts]

    def route_p2p(self, dst_ptrs):
        """write every pending key to dst_ptrs[owner] (device addresses, own or peer memory)"""
        arr = (vp * len(dst_ptrs))(*[vp(int(p)) for p in dst_ptrs])
        _check(lib().zb_kmerize_route_p2p(self.h, len(dst_ptrs), arr))

    def route_p2p_begin(self, dst_ptrs):
        """route_p2p on the context's second stream; returns at once (route_p2p_end waits)"""
        arr = (vp * len(dst_ptrs))(*[vp(int(p)) for p in dst_ptrs])
        _check(lib().zb_kmerize_route_p2p_begin(self.h, len(dst_ptrs), arr))

    def route_p2p_end(self):
        _check(lib().zb_kmerize_route_p2p_end(self.h))

    def route_p2p_reserve(self, dst_ptrs, cursor_ptrs, capacity_keys):
        """write every pending key to its owner's receive buffer (dst_ptrs[owner] = its START), reserving the place with
        an atomic add on the owner's cursor word (cursor_ptrs[owner]) -> keys sent per owner"""
        n = len(dst_ptrs)
        arr = (vp * n)(*[vp(int(p)) for p in dst_ptrs])
        cur = (vp * n)(*[vp(int(p)) for p in cursor_ptrs])
        sent = (C.c_uint64 * n)()
        _check(lib().zb_kmerize_route_p2p_reserve(self.h, n, arr, cur, int(capacity_keys), sent))
        return [int(x) for x in sent]

    def adopt_canonical_dev(self, dptr, n):
        _check(lib().zb_kmerize_adopt_canonical_dev(self.h, vp(dptr), n))

    def flush(self):
        """count what is pending / adopted now (an adopted array is free for its owner afterwards)"""
        _check(lib().zb_kmerize_flush(self.h))

    def add_canonical_dev(self, dptr, n):
        _check(lib().zb_kmerize_add_canonical_dev(self.h, vp(dptr), n))

    def finish(self):
        """-> (KmerSet with both strands, number of records)"""
        s = vp()
        nr = C.c_uint64(0)
        _check(lib().zb_kmerize_finish(self.h, C.byref(s), C.byref(nr)))
        return KmerSet(s, self.device), nr.value

    def close(self):
        if self.h is not None and self.h.value:
            lib().zb_kmerize_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_stream(vals, delta=False, device=0):
    v = np.ascontiguousarray(vals, dtype=np.uint64)
    w = np.empty(max(len(v), 1), np.uint64)
    nw = C.c_size_t(0)
    _check(lib().zb_encode_u64_stream(device, _ptr(v), len(v), 1 if delta else 0, _ptr(w), C.byref(nw)))
    return w[:nw.value]


def decode_stream(words, delta=False, device=0):
    w = np.ascontiguousarray(words, dtype=np.uint64)
    n = C.c_size_t(0)
    _check(lib().zb_decode_u64_stream(device, _ptr(w), len(w), 1 if delta else 0, None, C.byref(n)))
    out = np.empty(max(n.value, 1), np.uint64)
    _check(lib().zb_decode_u64_stream(device, _ptr(w), len(w), 1 if delta else 0, _ptr(out), C.byref(n)))
    return out[:n.value]


# ---- diagnostics (tests / bench) -------------------------------------------------------------
def dbg_sort(keys, vals=None, key_bits=64, max_bits=8, iters=1, device=0):
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    v = None if vals is None else np.ascontiguousarray(vals, dtype=np.uint32).copy()
    ms = C.c_float(0)
    _check(lib().zb_dbg_sort_u64(device, _ptr(k), _ptr(v) if v is not None else None, len(k), key_bits, max_bits,
                                 iters, C.byref(ms)))
    return k, v, ms.value


def dbg_sort_count(keys, weights=None, key_bits=64, mode=0, iters=1, device=0):
    """sort + run-length count: (distinct keys ascending, summed counts, ms per call)"""
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
    ok = np.empty(max(len(k), 1), np.uint64)
    oc = np.empty(max(len(k), 1), np.uint32)
    n = C.c_size_t(0)
    ms = C.c_float(0)
    _check(lib().zb_dbg_sort_count(device, _ptr(k), _ptr(w) if w is not None else None, len(k), key_bits, mode, iters,
                                   _ptr(ok), _ptr(oc), C.byref(n), C.byref(ms)))
    return ok[:n.value].copy(), oc[:n.value].copy(), ms.value


def dbg_parse(data, is_fasta, device=0):
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    codes = np.empty(len(a) + 64, np.uint8)
    n = C.c_size_t(0)
    nr = C.c_uint64(0)
    _check(lib().zb_dbg_parse(device, _ptr(a), len(a), 1 if is_fasta else 0, _ptr(codes), C.byref(n), C.byref(nr)))
    return codes[:n.value].copy(), nr.value


def dbg_extract(k, codes, device=0):
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    keys = np.empty(max(len(c), 1), np.uint64)
    n = C.c_size_t(0)
    _check(lib().zb_dbg_extract(device, k, _ptr(c), len(c), _ptr(keys), C.byref(n)))
    return keys[:n.value].copy()


def dbg_profile(on, device=0):
    """switch per-stage timing on/off; returns {stage: (total_ms, calls)} collected so far"""
    buf = C.create_string_buffer(8192)
    _check(lib().zb_dbg_profile(device, 1 if on else 0, buf, len(buf)))
    out = {}
    for ln in buf.value.decode().splitlines():
        nm, ms, calls = ln.split()
        out[nm] = (float(ms), int(calls))
    return out


def timer_start(device=0):
    _check(lib().zb_dbg_timer(device, 0, None))


def timer_stop(device=0):
    ms = C.c_float(0)
    _check(lib().zb_dbg_timer(device, 1, C.byref(ms)))
    return ms.value


def device_sync(device=0):
    _check(lib().zb_device_sync(device))
