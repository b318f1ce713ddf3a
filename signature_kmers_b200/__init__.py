"""signature_kmers_b200 — B200-native signature k-mer builder.

The product is csrc/ (hand-written sm_100a CUDA kernels behind the C ABI of
include/sigk.h) plus the C++ host code above it; this Python package is only
the ctypes driver the tests and bench.py use.
"""
from .capi import KeptTable, PackedProteins, SigkError, load_library  # noqa: F401
