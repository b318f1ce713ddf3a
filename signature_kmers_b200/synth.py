"""ctypes driver for tools/libsigk_synth.so: synthetic protein sets of the
shapes BASELINE.json names, as packed arrays (struct sigk_proteins) or as the
reference's on-disk input tree.  SURVEY.md section 8d defines the generator."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from .capi import PackedProteins

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "tools", "libsigk_synth.so")


class SynthParams(C.Structure):
    _fields_ = [
        ("n_proteins", C.c_uint64),
        ("n_functions", C.c_uint32),
        ("n_genomes", C.c_uint32),
        ("seed", C.c_uint64),
        ("zipf_s", C.c_double),
        ("mut_rate", C.c_double),
        ("indel_frac", C.c_double),
        ("x_rate", C.c_double),
        ("lower_rate", C.c_double),
        ("rare_rate", C.c_double),
        ("domain_frac", C.c_double),
        ("min_reps", C.c_uint32),
        ("max_seqs_per_file", C.c_uint32),
    ]


# The configurations of BASELINE.json (seeds 1-4).
CONFIGS = {
    "config1": dict(n_proteins=20_000, n_functions=1_000, n_genomes=10, seed=1),
    "config2": dict(n_proteins=2_000_000, n_functions=20_000, n_genomes=20, seed=2),
    "config3": dict(n_proteins=20_000_000, n_functions=100_000, n_genomes=200, seed=3),
    "config4": dict(n_proteins=2_000_000, n_functions=20_000, n_genomes=20, seed=4, zipf_s=1.1, mut_rate=0.02),
}

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-C", os.path.join(_ROOT, "tools")], check=True, capture_output=True)
        lib = C.CDLL(LIB_PATH)
        lib.sigk_synth_default_params.argtypes = [C.POINTER(SynthParams)]
        lib.sigk_synth_create.argtypes = [C.POINTER(SynthParams)]
        lib.sigk_synth_create.restype = C.c_void_p
        lib.sigk_synth_destroy.argtypes = [C.c_void_p]
        lib.sigk_synth_kept_functions.argtypes = [C.c_void_p]
        lib.sigk_synth_kept_functions.restype = C.c_uint32
        lib.sigk_synth_lengths.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        lib.sigk_synth_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.sigk_synth_write_tree.argtypes = [C.c_void_p, C.c_char_p]
        _lib = lib
    return _lib


@dataclass
class SynthSlice:
    proteins: PackedProteins
    canon_lo: int
    canon_hi: int
    ordinal_base: int      # gated proteins before this slice


class Synth:
    def __init__(self, **kw):
        lib = _load()
        p = SynthParams()
        lib.sigk_synth_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        self.params = p
        self.h = C.c_void_p(lib.sigk_synth_create(C.byref(p)))
        if not self.h:
            raise ValueError(f"bad synthetic parameters {kw}")
        self.lib = lib

    @staticmethod
    def config(name: str, **override) -> "Synth":
        kw = dict(CONFIGS[name])
        kw.update(override)
        return Synth(**kw)

    def close(self):
        if self.h:
            self.lib.sigk_synth_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_proteins(self) -> int:
        return int(self.params.n_proteins)

    @property
    def kept_functions(self) -> int:
        return int(self.lib.sigk_synth_kept_functions(self.h))

    def packed(self, lo: int = 0, hi: int | None = None, n_threads: int | None = None, out_alloc=None) -> PackedProteins:
        """Packed arrays for canonical positions [lo, hi) (default: everything).

        out_alloc(nbytes) -> writable uint8 numpy array lets the caller place the
        residue buffer in pinned memory."""
        hi = self.n_proteins if hi is None else hi
        n_threads = n_threads or min(32, os.cpu_count() or 1)
        n = hi - lo
        lens = np.zeros(n, dtype=np.uint32)
        gate = np.zeros(n, dtype=np.uint8)
        rc = self.lib.sigk_synth_lengths(self.h, lo, hi, lens.ctypes.data, gate.ctypes.data, n_threads)
        if rc != 0:
            raise RuntimeError("sigk_synth_lengths failed")
        g = gate.astype(bool)
        glen = lens[g].astype(np.uint64)
        starts = np.zeros(len(glen) + 1, dtype=np.uint64)
        np.cumsum(glen, out=starts[1:])
        total = int(starts[-1])
        residues = out_alloc(total) if out_alloc else np.empty(total, dtype=np.uint8)
        func = np.zeros(len(glen), dtype=np.uint16)
        sid = np.zeros(len(glen), dtype=np.uint32)
        rc = self.lib.sigk_synth_fill(self.h, lo, hi, gate.ctypes.data, starts.ctypes.data, residues.ctypes.data,
                                      func.ctypes.data, sid.ctypes.data, n_threads)
        if rc != 0:
            raise RuntimeError("sigk_synth_fill failed")
        return PackedProteins(residues, starts, func, sid)

    def write_tree(self, directory: str):
        if self.lib.sigk_synth_write_tree(self.h, directory.encode()) != 0:
            raise RuntimeError("sigk_synth_write_tree failed")
