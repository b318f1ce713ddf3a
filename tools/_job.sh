mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass|encode_sort|group_reduce|order_stats|window_count|head_tile|squeeze" -c 16 -o gpurun_out/r2_final_kernels python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
