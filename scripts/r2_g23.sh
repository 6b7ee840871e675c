mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g23_c3.json 2> gpurun_out/r2_g23_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g23_c3.json').read().strip().splitlines()[-1]); print('c3', d['ms_per_step'], d['roofline']['stages_ms']['epa'], d['roofline']['stages_ms']['gjk'])"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g23_c4.json 2> gpurun_out/r2_g23_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g23_c4.json').read().strip().splitlines()[-1]); print('c4', d['ms_per_step'], d['roofline']['stages_ms'])"
timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g23_c5.json 2> gpurun_out/r2_g23_c5.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g23_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], d['roofline']['stages_ms']['epa'])"
