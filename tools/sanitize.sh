#!/bin/bash
# compute-sanitizer passes over one small build (the smoke input) and one medium build through the C ABI:
# memcheck (out-of-bounds / misaligned accesses), racecheck (shared-memory hazards between the barrier-separated
# phases of the sort and reduce kernels), synccheck.  Logs land in gpurun_out/; the summaries are copied to profiles/.
#   gpurun -- bash tools/sanitize.sh
mkdir -p gpurun_out
cat > /tmp/sigk_sanitize_case.py <<'PY'
import sys
sys.path.insert(0, ".")
from oracle import oracle_c
from signature_kmers_b200.builder import GpuSignatureBuilder
from tests.util import assert_tables_equal, pack, random_proteins
cases = [dict(seed=2024, n_families=40, members=(2, 30), length=(20, 400), sub_rate=0.08),
         dict(seed=7, n_families=6, members=(40, 90), length=(200, 600), sub_rate=0.02, lower_rate=0.05)]
b = GpuSignatureBuilder(device=0)
for kw in cases:
    seed = kw.pop("seed")
    seqs, funcs = random_proteins(seed, **kw)
    seqs = list(seqs) + [b"W" * 700, b"W" * 650, b"Y" * 3000]       # long groups: whole-warp walks, the long order-statistics kernel
    funcs = list(funcs) + [1, 1, 2]
    p = pack(seqs, funcs)
    b.set_proteins(p)
    got = b.build()
    want, _ = oracle_c.oracle_build(p)
    assert_tables_equal(got, want, tier_b=True, what=str(seed))
    print("case ok:", got.n_occurrences, "occurrences,", got.n_kept, "kept", flush=True)
b.close()
print("SANITIZE_CASE_PASSED")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sigk_sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE_PASSED|hazard" gpurun_out/sanitize_$tool.log | head -8
done
