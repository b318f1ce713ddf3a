"""Full-size parity under pytest (-m gpu): the configurations BASELINE.json names, at their full sizes, against the
CPU oracle (tests/fullsize_check.py), the 100 K-query call chain of config 5 (tests/config5_check.py) and the
multi-GPU range-partitioned build on every GPU of the box (tests/multigpu_check.py).  About six minutes of the GPU
box's host cores in all; set SIGK_SKIP_FULLSIZE=1 to leave them out of a quick run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(bool(os.environ.get("SIGK_SKIP_FULLSIZE")), reason="SIGK_SKIP_FULLSIZE set")]
THREADS = os.cpu_count() or 1


def test_config2_full_size_tier_a():
    """2 M proteins / 20 K functions (593 M occurrences): every row of the kept table, order-independent columns."""
    from tests import fullsize_check

    assert fullsize_check.run("config2", THREADS) > 300_000_000


def test_config2_prefix_tier_a_and_b():
    """The first 400 K proteins of config 2 in canonical order against the 1-thread oracle: median and var too."""
    from tests import fullsize_check

    fullsize_check.run("config2", 1, prefix=400_000)


def test_config4_zipf_full_size_tier_a():
    """Zipf-skewed family sizes with heavy shared-k-mer duplication (giant groups, wrapping 16-bit sums)."""
    from tests import fullsize_check

    fullsize_check.run("config4", THREADS)


def test_config4_zipf_prefix_tier_a_and_b():
    from tests import fullsize_check

    fullsize_check.run("config4", 1, prefix=300_000)


def test_more_than_65535_functions_wrap_like_the_reference():
    """Config-3-shaped function count: 72 000 families, so that the reference's `unsigned short next` wraps
    (src/function_map.h:324-330), later functions alias earlier indices and index 0xFFFF proteins are skipped
    (src/signature_build.tcc:155-158); both halves of the distinct_functions counters are in use."""
    from tests import fullsize_check

    fullsize_check.run("wrap72k", 1, overrides=dict(n_proteins=290_000, n_functions=72_000, n_genomes=4, seed=33))


def test_config5_call_chain_100k_queries():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "config5_check.py")], cwd=ROOT, capture_output=True, text=True, timeout=1000)
    assert "CONFIG5_CHECK_PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_multi_gpu_range_partition_on_every_gpu():
    """The NCCL range-partitioned build (csrc/comm.cu) on min(8, all) GPUs of the box against the oracle and the
    one-GPU build, with every encode+route kernel instantiation (SIGK_CHECK_WIDE)."""
    import torch

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus N); pass logs of the 2- and 8-GPU runs are kept under profiles/")
    env = dict(os.environ, SIGK_CHECK_WIDE="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
                        "127.0.0.1", "--master-port", "29517", "tests/multigpu_check.py"], cwd=ROOT, capture_output=True, text=True,
                       timeout=1100, env=env)
    assert "MULTIGPU_CHECK_PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
