"""ctypes loader for oracle/liboracle.so (the C++ CPU restatement).

TEST INFRASTRUCTURE ONLY — see oracle/sigk_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from signature_kmers_b200.capi import KeptTable, PackedProteins, SigkProteins, SigkTable, table_to_numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

IMMEDIATE_MEAN = 0x1
NO_SORT = 0x2

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sigk_oracle.cpp")
    stale = (
        force
        or not os.path.exists(LIB_PATH)
        or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "sigk_oracle.h")))
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.sigk_oracle_build.argtypes = [C.POINTER(SigkProteins), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.sigk_oracle_result.argtypes = [C.c_void_p, C.POINTER(SigkTable)]
        lib.sigk_oracle_seconds.argtypes = [C.c_void_p]
        lib.sigk_oracle_seconds.restype = C.c_double
        lib.sigk_oracle_extract_seconds.argtypes = [C.c_void_p]
        lib.sigk_oracle_extract_seconds.restype = C.c_double
        lib.sigk_oracle_free.argtypes = [C.c_void_p]
        lib.sigk_oracle_free.restype = None
        lib.sigk_oracle_accumulate.argtypes = [
            C.c_void_p, C.c_uint64, C.c_int,
            C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
            C.POINTER(C.c_double), C.POINTER(C.c_double),
        ]
        lib.sigk_oracle_accumulate.restype = None
        lib.sigk_oracle_u16_from_double.argtypes = [C.c_double]
        lib.sigk_oracle_u16_from_double.restype = C.c_uint16
        lib.sigk_oracle_keep.argtypes = [C.c_int, C.c_int]
        lib.sigk_oracle_keep.restype = C.c_int
        lib.sigk_oracle_tbb_hash.argtypes = [C.c_char_p]
        lib.sigk_oracle_tbb_hash.restype = C.c_uint64
        _lib = lib
    return _lib


def oracle_build(p: PackedProteins, n_threads: int = 1, flags: int = 0, want_table: bool = True):
    """Run the CPU oracle.  Returns (KeptTable | None, seconds of extract+process)."""
    lib = load()
    h = C.c_void_p()
    st = p.as_struct()
    rc = lib.sigk_oracle_build(C.byref(st), n_threads, flags, C.byref(h))
    if rc != 0:
        raise RuntimeError(f"sigk_oracle_build failed: {rc}")
    try:
        secs = lib.sigk_oracle_seconds(h)
        if not want_table:
            t = SigkTable()
            lib.sigk_oracle_result(h, C.byref(t))
            return _CountsOnly(int(t.n_kept), int(t.n_occurrences), int(t.n_distinct_kmers)), secs
        t = SigkTable()
        lib.sigk_oracle_result(h, C.byref(t))
        return table_to_numpy(t, copy=True), secs
    finally:
        lib.sigk_oracle_free(h)


class _CountsOnly:
    def __init__(self, n_kept, n_occurrences, n_distinct_kmers):
        self.n_kept = n_kept
        self.n_occurrences = n_occurrences
        self.n_distinct_kmers = n_distinct_kmers


def accumulate(samples, flags: int = 0):
    """mean, median, var (u16) and the f64 median/var of the Boost accumulator set."""
    lib = load()
    a = np.ascontiguousarray(samples, dtype=np.uint32)
    m, md, v = C.c_uint16(), C.c_uint16(), C.c_uint16()
    mdf, vf = C.c_double(), C.c_double()
    lib.sigk_oracle_accumulate(a.ctypes.data, len(a), flags, C.byref(m), C.byref(md), C.byref(v), C.byref(mdf), C.byref(vf))
    return m.value, md.value, v.value, mdf.value, vf.value


def u16_from_double(d: float) -> int:
    return load().sigk_oracle_u16_from_double(d)


def keep(best_count: int, count: int) -> bool:
    return bool(load().sigk_oracle_keep(best_count, count))


def tbb_hash(kmer: str) -> int:
    return load().sigk_oracle_tbb_hash(kmer.encode("latin-1"))
