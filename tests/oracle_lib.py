"""ctypes binding of oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the mauvealigner_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
MODE_UNIQUE, MODE_SEED_ENUM, MODE_UNIQUE_COUNT, MODE_PAIRWISE, MODE_REPEAT = 0, 1, 2, 3, 4


class _Result(C.Structure):
    _fields_ = [
        ("n_matches", C.c_uint64), ("n_comps", C.c_uint64),
        ("length", C.POINTER(C.c_uint32)), ("comp_off", C.POINTER(C.c_uint64)),
        ("comp_seq", C.POINTER(C.c_uint32)), ("comp_start", C.POINTER(C.c_int64)),
        ("unique_mers", C.c_uint64), ("unique_mers_per_seq", C.POINTER(C.c_uint64)),
        ("nseq", C.c_uint32),
        ("n_seeds", C.c_uint64), ("n_buckets", C.c_uint64), ("n_candidates", C.c_uint64), ("n_contained", C.c_uint64),
        ("t_mers", C.c_double), ("t_sort", C.c_double), ("t_match", C.c_double), ("t_total", C.c_double),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_find.restype = C.c_int
        _lib.orc_find.argtypes = [C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint64, C.c_int,
                                  C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.POINTER(_Result))]
        _lib.orc_find_family.restype = C.c_int
        _lib.orc_find_family.argtypes = [C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32,
                                         C.c_uint64, C.POINTER(C.POINTER(_Result))]
        _lib.orc_result_free.argtypes = [C.POINTER(_Result)]
        _lib.orc_mers.restype = C.c_int64
        _lib.orc_mers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        _lib.orc_sml.restype = C.c_int64
        _lib.orc_sml.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        _lib.orc_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _lib.orc_seed_valid.argtypes = [C.c_uint64]
        _lib.orc_seed_length.argtypes = [C.c_uint64]
        _lib.orc_seed_weight.argtypes = [C.c_uint64]
    return _lib


def _as_u8(s):
    if isinstance(s, str):
        s = s.encode()
    if isinstance(s, (bytes, bytearray)):
        return np.frombuffer(bytes(s), dtype=np.uint8)
    return np.ascontiguousarray(s, dtype=np.uint8)


def find(seqs, pattern, mode, min_multi=2, max_multi=1000, direct_only=False, nway_mask=0):
    """Returns dict with numpy CSR arrays + stats (same field names as mauvealigner_b200 results)."""
    L = lib()
    arrs = [_as_u8(s) for s in seqs]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_uint64 * n)(*[a.size for a in arrs])
    out = C.POINTER(_Result)()
    rc = L.orc_find(n, ptrs, lens, pattern, mode, min_multi, max_multi, int(direct_only), nway_mask, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_find failed: {rc}")
    r = out.contents
    nm, nc = r.n_matches, r.n_comps
    res = dict(
        n_matches=nm, n_comps=nc,
        length=np.ctypeslib.as_array(r.length, shape=(nm + 1,))[:nm].copy(),
        comp_off=np.ctypeslib.as_array(r.comp_off, shape=(nm + 1,)).copy(),
        comp_seq=np.ctypeslib.as_array(r.comp_seq, shape=(nc + 1,))[:nc].copy(),
        comp_start=np.ctypeslib.as_array(r.comp_start, shape=(nc + 1,))[:nc].copy(),
        unique_mers=r.unique_mers,
        unique_mers_per_seq=np.ctypeslib.as_array(r.unique_mers_per_seq, shape=(n,)).copy(),
        n_seeds=r.n_seeds, n_buckets=r.n_buckets, n_candidates=r.n_candidates, n_contained=r.n_contained,
        t_mers=r.t_mers, t_sort=r.t_sort, t_match=r.t_match, t_total=r.t_total,
    )
    L.orc_result_free(out)
    return res


def find_family(seqs, patterns, nway_mask=0):
    """Seed-family search (orc_find_family): MODE_UNIQUE once per pattern, in order, one persistent MemHash table."""
    L = lib()
    arrs = [_as_u8(s) for s in seqs]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_uint64 * n)(*[a.size for a in arrs])
    pats = (C.c_uint64 * len(patterns))(*[int(p) for p in patterns])
    out = C.POINTER(_Result)()
    rc = L.orc_find_family(n, ptrs, lens, pats, len(patterns), nway_mask, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_find_family failed: {rc}")
    r = out.contents
    nm, nc = r.n_matches, r.n_comps
    res = dict(n_matches=nm, n_comps=nc,
               length=np.ctypeslib.as_array(r.length, shape=(nm + 1,))[:nm].copy(),
               comp_off=np.ctypeslib.as_array(r.comp_off, shape=(nm + 1,)).copy(),
               comp_seq=np.ctypeslib.as_array(r.comp_seq, shape=(nc + 1,))[:nc].copy(),
               comp_start=np.ctypeslib.as_array(r.comp_start, shape=(nc + 1,))[:nc].copy(),
               n_candidates=r.n_candidates, n_contained=r.n_contained)
    L.orc_result_free(out)
    return res


def matches_as_list(res):
    """[(length, [(g, start), ...]), ...] in result order."""
    out = []
    off = res["comp_off"]
    for i in range(int(res["n_matches"])):
        a, b = int(off[i]), int(off[i + 1])
        out.append((int(res["length"][i]), [(int(g), int(s)) for g, s in zip(res["comp_seq"][a:b], res["comp_start"][a:b])]))
    return out


def mers(seq, pattern):
    a = _as_u8(seq)
    Ls = lib().orc_seed_length(pattern)
    n = max(0, a.size - Ls + 1)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    got = lib().orc_mers(a.ctypes.data, a.size, pattern, out.ctypes.data)
    assert got == n, (got, n)
    return out[:n]


def sml(seq, pattern):
    a = _as_u8(seq)
    Ls = lib().orc_seed_length(pattern)
    n = max(0, a.size - Ls + 1)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    got = lib().orc_sml(a.ctypes.data, a.size, pattern, out.ctypes.data)
    assert got == n
    return out[:n]


def pack(seq):
    a = _as_u8(seq)
    out = np.zeros((a.size + 31) // 32 or 1, dtype=np.uint64)
    lib().orc_pack(a.ctypes.data, a.size, out.ctypes.data)
    return out[: (a.size + 31) // 32]
