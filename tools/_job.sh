mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for w in config2 config4; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_lg_$w.json 2> gpurun_out/r2_lg_$w.err
  python - $w <<PY
import json,sys
d=json.load(open("gpurun_out/r2_lg_%s.json" % sys.argv[1]))
print(sys.argv[1], "ms %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["pipeline"]["stage_ms"].items() if "reduce" in k or "order" in k or "squeeze" in k})
PY
done
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "prefix or wrap" 2>&1 | tail -3
