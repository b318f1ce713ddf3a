mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fasta.py -x -q -m gpu 2>&1 | tail -15
for combo in "0 0" "3 0" "0 3" "3 3"; do set -- $combo
  SIGK_TEST_REJ_SPREAD=$1 SIGK_TEST_META_SPREAD=$2 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_rej_$1_$2.json 2> gpurun_out/r2_rej_$1_$2.err
  python - $1 $2 <<PY
import json,sys
d=json.load(open("gpurun_out/r2_rej_%s_%s.json" % (sys.argv[1], sys.argv[2])))
print("rej spread", sys.argv[1], "meta spread", sys.argv[2], "ms %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["pipeline"]["stage_ms"].items() if "reduce" in k or "order" in k or "squeeze" in k}, "e2e %.1f" % d["e2e"]["ms_per_step"], d["e2e"].get("host_binding"))
PY
done
timeout 600 python tools/fasta_bench.py > gpurun_out/r2_fasta_bench.json 2> gpurun_out/r2_fasta_bench.err; echo "fasta bench exit $?"; cat gpurun_out/r2_fasta_bench.json; tail -3 gpurun_out/r2_fasta_bench.err
