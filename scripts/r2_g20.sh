mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for v in collide v_novab; do
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g20_c3_$v.json 2> gpurun_out/r2_g20_c3_$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g20_c3_$v.json').read().strip().splitlines()[-1]); print('c3 $v', d['ms_per_step'], d['roofline']['stages_ms']['epa'])"
done
