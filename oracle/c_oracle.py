"""
ctypes binding of oracle/libzot_oracle.so (plain-C restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY -- see the header of zot_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; never by zotmer_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzot_oracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def build():
    src = os.path.join(_HERE, "zot_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libzot_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.zo_extract.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                 C.POINTER(C.c_size_t), u64p, u64p]
        L.zo_extract.restype = C.c_int
        L.zo_sort_count.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.zo_sort_count.restype = C.c_int
        L.zo_merge.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                               C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.zo_merge.restype = C.c_int
        L.zo_split.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, u64p]
        L.zo_split.restype = None
        L.zo_project.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.zo_project.restype = C.c_size_t
        L.zo_trim.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.zo_trim.restype = C.c_size_t
        L.zo_hist.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.zo_hist.restype = C.c_size_t
        L.zo_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
        L.zo_encode.restype = C.c_int
        L.zo_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.zo_decode.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def extract(k, data, is_fasta):
    """-> (both-strand keys u64[], acgt[4], n_records) in emission order."""
    L = lib()
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)
    n = C.c_size_t(0)
    acgt = (C.c_uint64 * 4)()
    nrec = C.c_uint64(0)
    rc = L.zo_extract(k, _p(buf), len(buf), int(is_fasta), None, 0, C.byref(n), acgt, C.byref(nrec))
    assert rc == 0
    keys = np.empty(n.value, np.uint64)
    acgt = (C.c_uint64 * 4)()
    rc = L.zo_extract(k, _p(buf), len(buf), int(is_fasta), _p(keys), len(keys), C.byref(n), acgt, C.byref(nrec))
    assert rc == 0 and n.value == len(keys)
    return keys, [int(v) for v in acgt], int(nrec.value)


def sort_count(keys):
    L = lib()
    keys = np.ascontiguousarray(keys, np.uint64).copy()
    ks = np.empty(len(keys), np.uint64)
    cs = np.empty(len(keys), np.uint32)
    m = C.c_size_t(0)
    assert L.zo_sort_count(_p(keys), len(keys), _p(ks), _p(cs), C.byref(m)) == 0
    return ks[:m.value].copy(), cs[:m.value].copy()


def kmerize(k, inputs):
    """inputs = [(file bytes, is_fasta)] -> (kmers, counts(u32), acgt, n_records)"""
    allk = []
    acgt = [0, 0, 0, 0]
    nr = 0
    for data, is_fa in inputs:
        ks, a, r = extract(k, data, is_fa)
        allk.append(ks)
        acgt = [x + y for x, y in zip(acgt, a)]
        nr += r
    keys = np.concatenate(allk) if allk else np.zeros(0, np.uint64)
    ks, cs = sort_count(keys)
    return ks, cs, acgt, nr


def merge(sets):
    """sets = [(kmers u64, counts u64)] -> (kmers, counts u64)"""
    L = lib()
    ks = [np.ascontiguousarray(s[0], np.uint64) for s in sets]
    cs = [np.ascontiguousarray(s[1], np.uint64) for s in sets]
    n = len(sets)
    kp = (C.c_void_p * n)(*[a.ctypes.data for a in ks])
    cp = (C.c_void_p * n)(*[a.ctypes.data for a in cs])
    ln = (C.c_size_t * n)(*[len(a) for a in ks])
    tot = sum(len(a) for a in ks)
    ok = np.empty(tot, np.uint64)
    oc = np.empty(tot, np.uint64)
    m = C.c_size_t(0)
    assert L.zo_merge(n, kp, cp, ln, _p(ok), _p(oc), C.byref(m)) == 0
    return ok[:m.value].copy(), oc[:m.value].copy()


def split(xs, ys):
    L = lib()
    xs = np.ascontiguousarray(xs, np.uint64)
    ys = np.ascontiguousarray(ys, np.uint64)
    abc = (C.c_uint64 * 3)()
    L.zo_split(_p(xs), len(xs), _p(ys), len(ys), abc)
    return int(abc[0]), int(abc[1]), int(abc[2])


def project(xs, shift):
    L = lib()
    xs = np.ascontiguousarray(xs, np.uint64)
    out = np.empty(len(xs), np.uint64)
    m = L.zo_project(_p(xs), len(xs), shift, _p(out))
    return out[:m].copy()


def trim(xs, cs, c, C_=0):
    L = lib()
    xs = np.ascontiguousarray(xs, np.uint64)
    cs = np.ascontiguousarray(cs, np.uint64)
    ox = np.empty(len(xs), np.uint64)
    oc = np.empty(len(xs), np.uint64)
    m = L.zo_trim(_p(xs), _p(cs), len(xs), c, C_, _p(ox), _p(oc))
    return ox[:m].copy(), oc[:m].copy()


def hist(cs):
    L = lib()
    cs = np.ascontiguousarray(cs, np.uint64)
    v = np.empty(len(cs), np.uint64)
    f = np.empty(len(cs), np.uint64)
    m = L.zo_hist(_p(cs), len(cs), _p(v), _p(f))
    return [(int(a), int(b)) for a, b in zip(v[:m], f[:m])]


def encode(vals, delta=False):
    L = lib()
    vals = np.ascontiguousarray(vals, np.uint64)
    ws = np.empty(len(vals), np.uint64)
    m = C.c_size_t(0)
    rc = L.zo_encode(_p(vals), len(vals), int(delta), _p(ws), C.byref(m))
    if rc != 0:
        raise IndexError("codec64: value needs more than 60 bits")
    return ws[:m.value].copy()


def decode(ws, delta=False):
    L = lib()
    ws = np.ascontiguousarray(ws, np.uint64)
    out = np.empty(6 * len(ws), np.uint64)
    m = L.zo_decode(_p(ws), len(ws), int(delta), _p(out))
    assert m != C.c_size_t(-1).value
    return out[:m].copy()
