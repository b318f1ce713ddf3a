"""The device-side length accumulator (csrc/length_acc.cuh: wrapping sum,
P-square median, iterative variance) compiled for the host and checked bit for
bit against the oracle's Boost restatement on random and adversarial streams."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "host_harness")


@pytest.fixture(scope="module")
def hostacc():
    so = os.path.join(HARNESS, "liblength_acc_host.so")
    subprocess.run(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wl,--exclude-libs,ALL", "-I/usr/local/cuda/include",
                    "-I" + os.path.join(ROOT, "signature_kmers_b200", "csrc"), "-o", so,
                    os.path.join(HARNESS, "length_acc_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.sigk_host_length_acc.argtypes = [C.c_void_p, C.c_uint64] + [C.POINTER(C.c_uint32)] * 3 + [C.POINTER(C.c_double)] * 2
    lib.sigk_host_u16.argtypes = [C.c_double]
    lib.sigk_host_u16.restype = C.c_uint32
    lib.sigk_host_symbol.argtypes = [C.c_uint]

    def run(samples):
        a = np.ascontiguousarray(samples, dtype=np.uint32)
        m, md, v = C.c_uint32(), C.c_uint32(), C.c_uint32()
        mf, vf = C.c_double(), C.c_double()
        lib.sigk_host_length_acc(a.ctypes.data, len(a), C.byref(m), C.byref(md), C.byref(v), C.byref(mf), C.byref(vf))
        return m.value, md.value, v.value, mf.value, vf.value

    run.lib = lib
    return run


def streams():
    rng = np.random.default_rng(0)
    yield [300] * 1000
    yield list(range(1, 200))
    yield list(range(200, 0, -1))
    yield [5, 5, 5, 5, 5, 1, 9, 5, 5, 1, 9, 9, 9, 1, 1]
    yield [0] * 40 + [3] + [0] * 40                     # zero heights: signed zeros in the equal-neighbour shortcut
    yield [300] * 200 + [290] + [300] * 300 + [310, 310] + [300] * 100
    for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 17, 100, 1000, 5000):
        yield rng.integers(50, 3000, n)
        yield rng.integers(290, 310, n)                 # many ties: upper_bound and <= paths
        yield rng.integers(60000, 70000, n)             # wrapping sums, > 16 bit samples
        yield np.round(rng.lognormal(5.7, 0.5, n)).astype(np.int64).clip(8, 40000)
    yield rng.integers(0, 2, 4000) * 65535 + 1          # variance far beyond 2^31 -> u16 gives 0


def test_device_accumulator_matches_oracle(hostacc, oracle):
    for s in streams():
        got = hostacc(s)
        want = oracle.accumulate(s)
        assert got == want, (list(s)[:12], got, want)


def test_u16_conversion(hostacc, oracle):
    for d in (0.0, 299.99, 65535.9, 65536.0, 70000.7, 2147483647.0, 2147483647.9, 2147483648.0, 1e300, float("nan"), float("inf")):
        assert hostacc.lib.sigk_host_u16(d) == oracle.u16_from_double(d), d


def test_symbol_table(hostacc):
    ok = b"ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwy"     # reference src/signature_build.h:102-103
    for c in range(256):
        # rank of the letter among the 20 amino acids, + 32 when it is lower case; -1 outside ok_prot_
        want = ok.index(bytes([c])) % 20 + (32 if ok.index(bytes([c])) >= 20 else 0) if bytes([c]) in ok else -1
        assert hostacc.lib.sigk_host_symbol(c) == want, c
