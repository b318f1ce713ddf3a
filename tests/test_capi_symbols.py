"""The C-ABI library loads on a CPU-only box and exports every symbol
include/sigk.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from signature_kmers_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return capi.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sigk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sigk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_symbol_list_agree():
    assert declared_symbols() == sorted(capi.EXPORTED_SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_host_side_code_functions(lib):
    from signature_kmers_b200.builder import kmer_decode, kmer_encode

    # group code = base-20 code of the case-folded residues << 8 | case mask (residue j = bit j)
    assert kmer_encode("AAAAAAAA") == 0
    assert kmer_encode("YYYYYYYY") == (20 ** 8 - 1) << 8
    assert kmer_encode("yyyyyyyy") == ((20 ** 8 - 1) << 8) | 0xFF
    assert kmer_encode("aAAAAAAA") == 1 and kmer_encode("AAAAAAAa") == 0x80
    assert kmer_encode("ACDEFGHX") == 2 ** 64 - 1
    for s in ("ACDEFGHI", "WYwyACac", "yyyyyyyA"):
        assert kmer_decode(kmer_encode(s)) == s
    # table order (include/sigk.h) = ascending (case mask != 0, group code)
    ks = ["ACDEFGHI", "ACDEFGHi", "aCDEFGHI", "YYYYYYYY", "AAAAAAAC", "yYYYYYYY", "AAAAAAAc"]
    by_code = sorted(ks, key=lambda k: ((kmer_encode(k) & 0xFF) != 0, kmer_encode(k)))
    assert by_code == sorted(ks, key=capi.table_order_key)
    assert by_code[:3] == ["AAAAAAAC", "ACDEFGHI", "YYYYYYYY"]


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = capi.SigkConfig(capi.SIGK_ABI_VERSION, 8, 0, 0, 1, 0)
    h = C.c_void_p()
    assert lib.sigk_create(C.byref(cfg), C.byref(h)) == -2      # SIGK_E_CUDA
    assert b"no CPU fallback" in lib.sigk_last_error(None)


def test_product_never_references_the_oracle():
    # the product path must not import, link or call anything under oracle/
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "signature_kmers_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b", text) and f != "builder.py":
                    bad.append(os.path.join(dirpath, f))
                if re.search(r"(import|from)\s+oracle", text):
                    bad.append(os.path.join(dirpath, f))
                if "refshim" in text or "_ref/" in text or "/root/reference" in text:      # stand-ins and reference builds are test-only
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_built_product_binaries_do_not_link_test_infrastructure():
    """libsigk.so and the two command lines depend on CUDA and the C/C++ runtime only."""
    import subprocess

    pkg = os.path.join(ROOT, "signature_kmers_b200")
    for name in ("libsigk.so", "kmers-build-signatures", "kmers-call-functions"):
        path = os.path.join(pkg, name)
        if not os.path.exists(path):
            continue
        deps = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "liboracle" not in deps and "libref_" not in deps and "libsigk_synth" not in deps and "libsigk_host" not in deps, (name, deps)
