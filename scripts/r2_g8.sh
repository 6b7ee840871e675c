mkdir -p gpurun_out
for v in collide v_stream; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g8_$v.json 2> gpurun_out/r2_g8_$v.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --side 50 > gpurun_out/r2_g8_side50.json 2> gpurun_out/r2_g8_side50.err
for v in collide v_stream side50; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g8_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"], d["roofline"]["stages_ms"]["gjk"], round(1e3*d["config"]["pairs_per_step"]/d["e2e"]["value"],3), d["config"]["contacts_per_step"])
except Exception as e: print("$v", "ERR", e)
PY
done
timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g8_c4.json 2> gpurun_out/r2_g8_c4.err; tail -c 600 gpurun_out/r2_g8_c4.json
timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g8_c5.json 2> gpurun_out/r2_g8_c5.err; tail -c 600 gpurun_out/r2_g8_c5.json
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_timing.so timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_g8_timing.txt 2> gpurun_out/r2_g8_timing.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
