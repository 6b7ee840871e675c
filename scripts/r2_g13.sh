mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_g13_c5_n1.json 2> gpurun_out/r2_g13_c5_n1.err; tail -c 1500 gpurun_out/r2_g13_c5_n1.json | head -c 1500; echo
timeout 600 python bench.py --workload c5 --worlds 512 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g13_c5_512.json 2> gpurun_out/r2_g13_c5_512.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g13_c5_512.json').read().strip().splitlines()[-1]); print('c5 512 worlds', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stages_ms'])"
