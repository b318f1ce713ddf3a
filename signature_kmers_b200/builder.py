"""Python driver over the C ABI (tests and bench only; the product host code is
the C++ above libsigk).  Mirrors the two reference calls it replaces:

    builder.extract_kmers(deleted_fids); builder.process_kmers();
    (reference src/kmers-build-signatures.cc:194-196)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import KeptTable, PackedProteins, SigkConfig, SigkError, SigkTable, SigkTimings


class GpuSignatureBuilder:
    """One libsigk handle on one GPU."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, flags: int = 0):
        self.lib = capi.load_library()
        cfg = SigkConfig(capi.SIGK_ABI_VERSION, capi.SIGK_K, device, rank, world, flags)
        h = C.c_void_p()
        rc = self.lib.sigk_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise SigkError(f"sigk_create failed ({rc}): {self.lib.sigk_last_error(None).decode()}")
        self.h = h
        self._proteins = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sigk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise SigkError(f"{what} failed ({rc}): {self.lib.sigk_last_error(self.h).decode()}")

    def set_proteins(self, p: PackedProteins):
        self._proteins = p  # keep the arrays alive
        st = p.as_struct()
        self._check(self.lib.sigk_set_proteins(self.h, C.byref(st)), "sigk_set_proteins")

    def build(self, fetch: bool = True):
        """extract_kmers + process_kmers: host arrays in, kept table out (in host memory
        owned by the handle; fetch=True also copies it into numpy arrays)."""
        self._check(self.lib.sigk_build(self.h), "sigk_build")
        return self.result() if fetch else None

    def upload(self):
        self._check(self.lib.sigk_upload(self.h), "sigk_upload")

    def build_device(self):
        self._check(self.lib.sigk_build_device(self.h), "sigk_build_device")

    def download(self):
        self._check(self.lib.sigk_download(self.h), "sigk_download")

    def result(self, copy: bool = True) -> KeptTable:
        t = SigkTable()
        self._check(self.lib.sigk_result(self.h, C.byref(t)), "sigk_result")
        return capi.table_to_numpy(t, copy=copy)

    def result_counts(self):
        t = SigkTable()
        self._check(self.lib.sigk_result(self.h, C.byref(t)), "sigk_result")
        return dict(n_kept=int(t.n_kept), n_occurrences=int(t.n_occurrences), n_distinct_kmers=int(t.n_distinct_kmers),
                    num_seqs_with_a_signature=int(t.num_seqs_with_a_signature))

    def synchronize(self):
        self._check(self.lib.sigk_synchronize(self.h), "sigk_synchronize")

    def event_record(self, slot: int):
        self._check(self.lib.sigk_event_record(self.h, slot), "sigk_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self.lib.sigk_event_elapsed_ms(self.h, a, b, C.byref(ms)), "sigk_event_elapsed_ms")
        return float(ms.value)

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a uint8 numpy array (freed with the process)."""
        ptr = self.lib.sigk_host_alloc(max(1, nbytes))
        if not ptr:
            raise SigkError("sigk_host_alloc failed")
        buf = (C.c_uint8 * max(1, nbytes)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint8, count=nbytes)

    def timings(self) -> dict:
        t = SigkTimings()
        self._check(self.lib.sigk_get_timings(self.h, C.byref(t)), "sigk_get_timings")
        return t.as_dict()

    # ---- stand-alone stages (parity tests) ----
    def dbg_encode(self, p: PackedProteins):
        cap = max(1, len(p.residues))
        code = np.zeros(cap, dtype=np.uint64)
        ordinal = np.zeros(cap, dtype=np.uint32)
        offset = np.zeros(cap, dtype=np.uint16)
        n = C.c_uint64()
        st = p.as_struct()
        self._proteins = p
        self._check(self.lib.sigk_dbg_encode(self.h, C.byref(st), code.ctypes.data, ordinal.ctypes.data, offset.ctypes.data, cap, C.byref(n)), "sigk_dbg_encode")
        k = int(n.value)
        return code[:k], ordinal[:k], offset[:k]

    def dbg_sort_pairs(self, keys: np.ndarray, vals: np.ndarray, bit_lo: int, bit_hi: int):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._check(self.lib.sigk_dbg_sort_pairs(self.h, keys.ctypes.data, vals.ctypes.data, len(keys), bit_lo, bit_hi), "sigk_dbg_sort_pairs")
        return keys, vals

    # ---- FASTA bytes -> proteins on the device (sigk_fasta_parse / sigk_fasta_commit) ----
    def fasta_parse(self, files):
        """files: list of bytes objects, one per FASTA file.  Returns a dict: the record table as numpy arrays (copies),
        the concatenated buffer the positions refer to, every file's (begin, length), and the parse timings."""
        begins, lens, at = [], [], 0
        for f in files:
            begins.append(at); lens.append(len(f))
            at += (len(f) + 15) // 16 * 16
        buf = np.zeros(max(at, 16), dtype=np.uint8)
        for b, f in zip(begins, files):
            buf[b:b + len(f)] = np.frombuffer(f, dtype=np.uint8)
        fb, fl = np.asarray(begins, dtype=np.uint64), np.asarray(lens, dtype=np.uint64)
        rec = capi.SigkFastaRecords()
        self._check(self.lib.sigk_fasta_parse(self.h, buf.ctypes.data, fb.ctypes.data, fl.ctypes.data, len(files), C.byref(rec)), "sigk_fasta_parse")
        n = int(rec.n_records)

        def arr(ptr, count, dtype=np.uint64):
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True) if count else np.zeros(0, dtype=dtype)

        ne = min(int(rec.n_errors), capi.SIGK_FASTA_MAX_ERRORS)
        return {"n_records": n, "n_residues": int(rec.n_residues), "n_errors": int(rec.n_errors),
                "header_pos": arr(rec.header_pos, n), "id_end": arr(rec.id_end, n), "line_end": arr(rec.line_end, n),
                "seq_begin": arr(rec.seq_begin, n + 1), "errors": arr(rec.errors, ne), "error_record": arr(rec.error_record, ne, np.uint32),
                "bytes": buf, "file_begin": fb, "file_len": fl,
                "h2d_ms": rec.h2d_ms, "parse_ms": rec.parse_ms, "d2h_ms": rec.d2h_ms}

    def fasta_commit(self, keep, function_index, seq_id):
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        function_index = np.ascontiguousarray(function_index, dtype=np.uint16)
        seq_id = np.ascontiguousarray(seq_id, dtype=np.uint32)
        self._proteins = None
        self._check(self.lib.sigk_fasta_commit(self.h, keep.ctypes.data, function_index.ctypes.data, seq_id.ctypes.data), "sigk_fasta_commit")

    def dbg_fasta_stream(self, n_residues: int) -> np.ndarray:
        out = np.zeros(max(n_residues, 1), dtype=np.uint8)
        self._check(self.lib.sigk_dbg_fasta_stream(self.h, out.ctypes.data), "sigk_dbg_fasta_stream")
        return out[:n_residues]

    def dbg_ddiv(self, a: np.ndarray, b: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        inl, lib = np.empty_like(a), np.empty_like(a)
        self._check(self.lib.sigk_dbg_ddiv(self.h, a.ctypes.data, b.ctypes.data, len(a), inl.ctypes.data, lib.ctypes.data), "sigk_dbg_ddiv")
        return inl, lib

    def lookup(self, residues: np.ndarray, starts: np.ndarray) -> np.ndarray:
        """rows[g] = table row of the call-side window at residue position g, or 0xFFFFFFFF (sigk_lookup)."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        rows = np.full(int(starts[-1]) if len(starts) else 0, 0xFFFFFFFF, dtype=np.uint32)
        self._check(self.lib.sigk_lookup(self.h, residues.ctypes.data, starts.ctypes.data, len(starts) - 1, rows.ctypes.data), "sigk_lookup")
        return rows

    def set_table(self, kmers: np.ndarray):
        """Make a sorted k-mer column (uint8 [n,8]) the lookup target (sigk_set_table)."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8)
        t = capi.SigkTable()
        t.n_kept = len(kmers)
        t.kmer = kmers.ctypes.data
        self._check(self.lib.sigk_set_table(self.h, C.byref(t)), "sigk_set_table")


def kmer_encode(kmer: str) -> int:
    return capi.load_library().sigk_kmer_encode(kmer.encode("latin-1"))


def kmer_decode(code: int) -> str:
    buf = C.create_string_buffer(8)
    capi.load_library().sigk_kmer_decode(code, buf)
    return buf.raw.decode("latin-1")
