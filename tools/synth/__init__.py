"""Synthetic genomes for BASELINE.json configs C1..C5 (C++ generator in synth.cpp, built into libmbsynth.so).

Bench / test infrastructure: loads only its own host library, never libmauve_b200.so, so bench.py's reference arm
can import it (by path) without mapping the product library."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libmbsynth.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = C.CDLL(path)
        L.mb_synth_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
        L.mb_synth_nseq.argtypes = [C.c_void_p]
        L.mb_synth_nseq.restype = C.c_uint32
        L.mb_synth_len.argtypes = [C.c_void_p, C.c_uint32]
        L.mb_synth_len.restype = C.c_uint64
        L.mb_synth_seq.argtypes = [C.c_void_p, C.c_uint32]
        L.mb_synth_seq.restype = C.c_void_p
        L.mb_synth_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def synth_genomes(config, scale=1):
    """Returns a list of numpy uint8 arrays (ASCII ACGT). Deterministic in (config, scale)."""
    L = lib()
    h = C.c_void_p()
    rc = L.mb_synth_create(int(config), int(scale), C.byref(h))
    if rc != 0:
        raise ValueError(f"mb_synth_create({config}, {scale}) failed: {rc}")
    try:
        out = []
        for i in range(L.mb_synth_nseq(h)):
            n = L.mb_synth_len(h, i)
            p = L.mb_synth_seq(h, i)
            out.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).copy())
        return out
    finally:
        L.mb_synth_free(h)
