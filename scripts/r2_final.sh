# Evidence of the final code of round 2: GPU tests, smoke, launch list, full captures of the step's kernels, bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --workload c1 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; tail -c 600 gpurun_out/r2_bench_c1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"bounds_fat_kernel|hierarchy_kernel|overlap_kernel|radix_scatter_kernel|pair_rows_emit_kernel|gjk_filter_kernel|gjk_kernel|epa_coop_kernel|epa_init_kernel" --launch-skip 31 -c 12 -o gpurun_out/r2_step_kernels python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_step_kernels.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -c 300 gpurun_out/r2_bench_c3.json
