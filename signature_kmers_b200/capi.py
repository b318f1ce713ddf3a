"""ctypes view of include/sigk.h.

This module is the only place Python touches libsigk.so.  It is a test/bench
driver: the product is the C-ABI library and the C++ host code above it.  There
is no fallback: if the shared library is missing or CUDA is unusable, calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIGK_LIB") or os.path.join(_HERE, "libsigk.so")   # SIGK_LIB: a tuning variant

SIGK_ABI_VERSION = 2
SIGK_K = 8
SIGK_UNDEFINED_FUNCTION = 0xFFFF
SIGK_N_FUNCTION_SLOTS = 65536
SIGK_F_NO_ORDER_STATS = 0x1
SIGK_COMM_ID_BYTES = 128


class SigkConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("k", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("flags", C.c_uint32),
    ]


class SigkProteins(C.Structure):
    _fields_ = [
        ("residues", C.c_void_p),
        ("starts", C.c_void_p),
        ("function_index", C.c_void_p),
        ("seq_id", C.c_void_p),
        ("n_proteins", C.c_uint64),
    ]


class SigkTable(C.Structure):
    _fields_ = [
        ("n_kept", C.c_uint64),
        ("kmer", C.c_void_p),
        ("avg_from_end", C.c_void_p),
        ("function_index", C.c_void_p),
        ("mean", C.c_void_p),
        ("median", C.c_void_p),
        ("var", C.c_void_p),
        ("n_occurrences", C.c_uint64),
        ("n_distinct_kmers", C.c_uint64),
        ("distinct_signatures", C.c_uint64),
        ("num_seqs_with_a_signature", C.c_uint64),
        ("distinct_functions", C.c_void_p),
        ("seqs_with_func", C.c_void_p),
        ("n_upper", C.c_uint64),
    ]


class SigkFastaRecords(C.Structure):
    """struct sigk_fasta_records (include/sigk.h)"""
    _fields_ = [
        ("n_records", C.c_uint64),
        ("n_residues", C.c_uint64),
        ("n_errors", C.c_uint64),
        ("header_pos", C.POINTER(C.c_uint64)),
        ("id_end", C.POINTER(C.c_uint64)),
        ("line_end", C.POINTER(C.c_uint64)),
        ("seq_begin", C.POINTER(C.c_uint64)),
        ("errors", C.POINTER(C.c_uint64)),
        ("error_record", C.POINTER(C.c_uint32)),
        ("h2d_ms", C.c_float),
        ("parse_ms", C.c_float),
        ("d2h_ms", C.c_float),
    ]


SIGK_FASTA_NO_POS = 0xFFFFFFFFFFFFFFFF
SIGK_FASTA_MAX_ERRORS = 65536


class SigkTimings(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float),
        ("encode_ms", C.c_float),
        ("histogram_ms", C.c_float),
        ("sort_ms", C.c_float),
        ("reduce_ms", C.c_float),
        ("order_stats_ms", C.c_float),
        ("squeeze_ms", C.c_float),
        ("exchange_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("device_total_ms", C.c_float),
        ("sort_passes", C.c_uint32),
        ("record_bytes", C.c_uint32),
        ("key_bytes", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("pass_ms", C.c_float * 8),
        ("count_ms", C.c_float),
        ("side_sort_ms", C.c_float),
        ("reduce_comm_ms", C.c_float),
        ("reduce_count_ms", C.c_float),
        ("reduce_emit_ms", C.c_float),
        ("reduce_groups_ms", C.c_float),
        ("records_sorted", C.c_uint64),
        ("exchange_bytes_out", C.c_uint64),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "pass_ms" else v
        return d


@dataclass
class PackedProteins:
    """Host arrays in the layout of struct sigk_proteins (include/sigk.h)."""

    residues: np.ndarray  # uint8 [total]
    starts: np.ndarray  # uint64 [n+1]
    function_index: np.ndarray  # uint16 [n]
    seq_id: np.ndarray  # uint32 [n]

    def __post_init__(self):
        self.residues = np.ascontiguousarray(self.residues, dtype=np.uint8)
        self.starts = np.ascontiguousarray(self.starts, dtype=np.uint64)
        self.function_index = np.ascontiguousarray(self.function_index, dtype=np.uint16)
        self.seq_id = np.ascontiguousarray(self.seq_id, dtype=np.uint32)
        n = len(self.function_index)
        if len(self.starts) != n + 1 or len(self.seq_id) != n:
            raise ValueError("starts must have n+1 entries, seq_id n")
        if n and int(self.starts[-1]) != len(self.residues):
            raise ValueError("starts[-1] must equal len(residues)")

    @property
    def n_proteins(self) -> int:
        return len(self.function_index)

    def as_struct(self) -> SigkProteins:
        return SigkProteins(
            self.residues.ctypes.data,
            self.starts.ctypes.data,
            self.function_index.ctypes.data,
            self.seq_id.ctypes.data,
            self.n_proteins,
        )

    @staticmethod
    def from_sequences(seqs, function_index, seq_id=None) -> "PackedProteins":
        bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
        starts = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            starts[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        residues = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
        if seq_id is None:
            seq_id = np.arange(len(bs), dtype=np.uint32)
        return PackedProteins(residues, starts, np.asarray(function_index, dtype=np.uint16), np.asarray(seq_id, dtype=np.uint32))

    def slice(self, lo: int, hi: int) -> "PackedProteins":
        s0, s1 = int(self.starts[lo]), int(self.starts[hi])
        return PackedProteins(
            self.residues[s0:s1].copy(),
            (self.starts[lo : hi + 1] - np.uint64(s0)).copy(),
            self.function_index[lo:hi].copy(),
            self.seq_id[lo:hi].copy(),
        )


@dataclass
class KeptTable:
    """Copy of struct sigk_table as numpy arrays (rows in the table order of include/sigk.h)."""

    kmer: np.ndarray  # uint8 [n,8]
    avg_from_end: np.ndarray
    function_index: np.ndarray
    mean: np.ndarray
    median: np.ndarray
    var: np.ndarray
    n_occurrences: int
    n_distinct_kmers: int
    distinct_signatures: int
    num_seqs_with_a_signature: int
    distinct_functions: np.ndarray  # uint32 [65536]
    seqs_with_func: np.ndarray  # uint32 [65536]
    n_upper: int = -1           # rows of the first section (k-mers without a lower-case residue); -1 = not recorded

    @property
    def n_kept(self) -> int:
        return len(self.avg_from_end)

    def kmer_strings(self):
        return [bytes(r).decode("latin-1") for r in self.kmer]

    def row(self, kmer: str):
        key = np.frombuffer(kmer.encode("latin-1"), dtype=np.uint8)
        idx = np.nonzero((self.kmer == key).all(axis=1))[0]
        if len(idx) == 0:
            return None
        i = int(idx[0])
        return dict(
            avg_from_end=int(self.avg_from_end[i]),
            function_index=int(self.function_index[i]),
            mean=int(self.mean[i]),
            median=int(self.median[i]),
            var=int(self.var[i]),
        )


def _arr(ptr, n, dtype, copy=True):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=n)
    return a.copy() if copy else a


def table_to_numpy(t: SigkTable, copy=True) -> KeptTable:
    n = int(t.n_kept)
    return KeptTable(
        kmer=_arr(t.kmer, n * 8, np.uint8, copy).reshape(n, 8),
        avg_from_end=_arr(t.avg_from_end, n, np.uint16, copy),
        function_index=_arr(t.function_index, n, np.uint16, copy),
        mean=_arr(t.mean, n, np.uint16, copy),
        median=_arr(t.median, n, np.uint16, copy),
        var=_arr(t.var, n, np.uint16, copy),
        n_occurrences=int(t.n_occurrences),
        n_distinct_kmers=int(t.n_distinct_kmers),
        distinct_signatures=int(t.distinct_signatures),
        num_seqs_with_a_signature=int(t.num_seqs_with_a_signature),
        distinct_functions=_arr(t.distinct_functions, SIGK_N_FUNCTION_SLOTS, np.uint32, copy),
        seqs_with_func=_arr(t.seqs_with_func, SIGK_N_FUNCTION_SLOTS, np.uint32, copy),
        n_upper=int(t.n_upper),
    )


# ---- the table order of include/sigk.h ---------------------------------------------------------
# k-mers without a lower-case residue first, in byte order; then the others by (case-folded bytes,
# case mask with residue j in bit j).

def table_order_key(kmer) -> tuple:
    """Sort key of one k-mer (str or bytes) in table order."""
    b = kmer.encode("latin-1") if isinstance(kmer, str) else bytes(kmer)
    mask = sum(((c >> 5) & 1) << j for j, c in enumerate(b))
    return (mask != 0, bytes(c & 0xDF for c in b), mask)


def kmer_case_masks(kmer: np.ndarray) -> np.ndarray:
    """uint8 [n,8] k-mer bytes -> uint8 [n] case masks (bit j set iff residue j is lower case)."""
    k = np.ascontiguousarray(kmer, dtype=np.uint8).reshape(-1, 8)
    return (((k >> 5) & 1).astype(np.uint16) << np.arange(8, dtype=np.uint16)).sum(axis=1).astype(np.uint8)


def table_order_argsort(kmer: np.ndarray) -> np.ndarray:
    """Permutation that puts uint8 [n,8] k-mer rows into table order."""
    k = np.ascontiguousarray(kmer, dtype=np.uint8).reshape(-1, 8)
    mask = kmer_case_masks(k)
    folded = np.ascontiguousarray(k & 0xDF).view(">u8").ravel().astype(np.uint64)
    return np.lexsort((mask, folded, mask != 0))


# Every symbol include/sigk.h declares; tests/test_capi_symbols.py checks the
# list against the header and against the built library.
EXPORTED_SYMBOLS = [
    "sigk_version",
    "sigk_device_count",
    "sigk_create",
    "sigk_destroy",
    "sigk_last_error",
    "sigk_host_alloc",
    "sigk_host_free",
    "sigk_set_proteins",
    "sigk_build",
    "sigk_upload",
    "sigk_build_device",
    "sigk_download",
    "sigk_result",
    "sigk_get_timings",
    "sigk_synchronize",
    "sigk_event_record",
    "sigk_event_elapsed_ms",
    "sigk_comm_make_id",
    "sigk_comm_join",
    "sigk_dbg_encode",
    "sigk_dbg_sort_pairs",
    "sigk_dbg_ddiv",
    "sigk_fasta_parse",
    "sigk_fasta_commit",
    "sigk_dbg_fasta_stream",
    "sigk_lookup",
    "sigk_set_table",
    "sigk_kmer_encode",
    "sigk_kmer_decode",
]

_lib = None


class SigkError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """Load libsigk.so.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SigkError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(libsigk has no CPU fallback)"
        )
    lib = C.CDLL(p)
    lib.sigk_version.restype = C.c_char_p
    lib.sigk_device_count.restype = C.c_int
    lib.sigk_create.argtypes = [C.POINTER(SigkConfig), C.POINTER(C.c_void_p)]
    lib.sigk_create.restype = C.c_int
    lib.sigk_destroy.argtypes = [C.c_void_p]
    lib.sigk_destroy.restype = None
    lib.sigk_last_error.argtypes = [C.c_void_p]
    lib.sigk_last_error.restype = C.c_char_p
    lib.sigk_host_alloc.argtypes = [C.c_size_t]
    lib.sigk_host_alloc.restype = C.c_void_p
    lib.sigk_host_free.argtypes = [C.c_void_p]
    lib.sigk_host_free.restype = None
    lib.sigk_set_proteins.argtypes = [C.c_void_p, C.POINTER(SigkProteins)]
    for name in ("sigk_build", "sigk_upload", "sigk_build_device", "sigk_download"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_int
    lib.sigk_result.argtypes = [C.c_void_p, C.POINTER(SigkTable)]
    lib.sigk_get_timings.argtypes = [C.c_void_p, C.POINTER(SigkTimings)]
    lib.sigk_synchronize.argtypes = [C.c_void_p]
    lib.sigk_event_record.argtypes = [C.c_void_p, C.c_int]
    lib.sigk_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.sigk_comm_make_id.argtypes = [C.c_char_p]
    lib.sigk_comm_join.argtypes = [C.c_void_p, C.c_char_p]
    lib.sigk_dbg_encode.argtypes = [
        C.c_void_p, C.POINTER(SigkProteins), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
    ]
    lib.sigk_dbg_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    lib.sigk_dbg_ddiv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.sigk_fasta_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SigkFastaRecords)]
    lib.sigk_fasta_commit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sigk_dbg_fasta_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.sigk_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.sigk_set_table.argtypes = [C.c_void_p, C.POINTER(SigkTable)]
    lib.sigk_kmer_encode.argtypes = [C.c_char_p]
    lib.sigk_kmer_encode.restype = C.c_uint64
    lib.sigk_kmer_decode.argtypes = [C.c_uint64, C.c_char_p]
    lib.sigk_kmer_decode.restype = None
    if path is None:
        _lib = lib
    return lib
