mkdir -p gpurun_out
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_timing.so timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_g5_timing.txt 2> gpurun_out/r2_g5_timing.err
grep -c "\[ec\]" gpurun_out/r2_g5_timing.txt
