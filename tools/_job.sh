mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for v in libsigk; do
  SIGK_LIB=$PWD/signature_kmers_b200/$v.so timeout 400 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4_$v.json 2> gpurun_out/r2_c4_$v.err
  python - $v <<PY
import json,sys
d=json.load(open("gpurun_out/r2_c4_%s.json" % sys.argv[1]))
print(sys.argv[1], "config4 ms %.2f  %.2f G/s" % (d["ms_per_step"], d["value"]/1e9), {k: round(v,2) for k,v in d["pipeline"]["stage_ms"].items()})
PY
done
bash tools/sweep_variants.sh 2>&1 | tee gpurun_out/r2_sweep12.txt
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "config4" 2>&1 | tail -5
