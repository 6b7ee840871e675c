mkdir -p gpurun_out
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_su.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel -c 1 -o gpurun_out/r2_g10_c4_su python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_g10_ncu.log 2>&1
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_r1.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_scan_kernel -c 1 -o gpurun_out/r2_g10_c4_r1 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_g10_ncu_r1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k c3_one_million 2>&1 | tail -3
