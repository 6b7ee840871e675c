mkdir -p gpurun_out
bash scripts/r2_ab.sh
PK_AB_ARGS="--side 50" bash scripts/r2_ab.sh
PK_AB_ARGS="--workload c5 --steps 5" bash scripts/r2_ab.sh
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_disc.so timeout 900 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/g39_tests.log 2>&1; echo "variant tests rc=$?"; tail -2 gpurun_out/g39_tests.log
for lib in libpk_collide libpk_v_disc; do
PK_COLLIDE_LIB=$PWD/physkit_b200/$lib.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:epa_coop --launch-skip 2 -c 1 --clock-control none --csv --log-file gpurun_out/g39_ncu_$lib.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/g39_ncu_$lib.csv | awk -F'","' -v l=$lib '{print l, $(NF-2), $(NF-1), $NF}' | tail -4
done
