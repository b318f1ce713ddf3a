mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fasta.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/fasta_bench.py > gpurun_out/r2_fasta_bench.json 2> gpurun_out/r2_fasta_bench.err; echo "fasta bench exit $?"; cat gpurun_out/r2_fasta_bench.json; tail -3 gpurun_out/r2_fasta_bench.err
