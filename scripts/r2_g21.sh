mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g21_c3_$i.json 2> gpurun_out/r2_g21_c3_$i.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g21_c3_$i.json').read().strip().splitlines()[-1]); print('c3 run $i', d['ms_per_step'], d['roofline']['stages_ms']['epa'], d['clocks'])"
done
