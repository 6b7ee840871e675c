mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g16_c3.json 2> gpurun_out/r2_g16_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g16_c3.json').read().strip().splitlines()[-1]); print('c3', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stages_ms'])"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g16_c4.json 2> gpurun_out/r2_g16_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g16_c4.json').read().strip().splitlines()[-1]); print('c4', d['ms_per_step'], d['roofline']['stages_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gjk_kernel|gjk_prefilter_kernel" -c 2 -o gpurun_out/r2_g16_gjk python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_g16_ncu.log 2>&1
