# Round-2 evidence of the final code: launch list, full captures of the step's kernels, racecheck, memcheck, bench lines
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"bounds_fat_kernel|hierarchy_kernel|overlap_kernel|radix_scatter_kernel|pair_rows_emit_kernel|gjk_filter_kernel|gjk_kernel|epa_coop_kernel|epa_init_kernel" --launch-skip 31 -c 12 -o gpurun_out/r2_step_kernels python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_step_kernels.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_epa_coop_final python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_epa_coop_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel --launch-skip 1 -c 1 -o gpurun_out/r2_epa_coop_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_epa_coop_c4.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_broadphase.py -m gpu -x -q -k "query_mode_pairs_exact or first_step or degenerate or sharding or c3_style or rows_overflow or read_back" > gpurun_out/r2_racecheck.log 2>&1; tail -15 gpurun_out/r2_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_broadphase.py -m gpu -x -q -k "kat or recoverable or first_step or all_shape_kinds or grazing or rows_overflow" > gpurun_out/r2_memcheck.log 2>&1; tail -8 gpurun_out/r2_memcheck.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -c 400 gpurun_out/r2_bench_c3.json
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_ref.json
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4.json 2>/dev/null
python bench.py --workload c4 --pairs 10000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4_10M.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_c4_10M.json
python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --no-cpu --manifolds > gpurun_out/r2_bench_c3_downstream.json 2>/dev/null
