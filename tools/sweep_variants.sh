#!/bin/bash
# Run bench.py once per built tuning variant (signature_kmers_b200/libsigk*.so) and print the stage times.
mkdir -p gpurun_out
for so in signature_kmers_b200/libsigk.so signature_kmers_b200/libsigk_[a-z].so; do [ -f $so ] || continue
  name=$(basename $so .so)
  SIGK_LIB=$PWD/$so timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.log || { echo "$name FAILED"; tail -3 gpurun_out/sweep_$name.log; continue; }
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/sweep_{sys.argv[1]}.json"))
st = d["pipeline"]["stage_ms"]
print(sys.argv[1], "total %.2f" % d["ms_per_step"], "enc %.2f hist %.2f sort %.2f red %.2f ord %.2f sq %.2f" % (st["encode_ms"], st["histogram_ms"], st["sort_ms"], st["reduce_ms"], st["order_stats_ms"], st["squeeze_ms"]), "passes", [round(x, 2) for x in d["roofline"]["pass_ms"]])
PY
done
