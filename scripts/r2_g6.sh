mkdir -p gpurun_out
for v in collide v_hf0 v_hf8 v_hf20 v_nopf; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g6_$v.json 2> gpurun_out/r2_g6_$v.err
done
for v in collide v_hf0 v_hf8 v_hf20 v_nopf; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g6_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"], d["roofline"]["stages_ms"]["gjk"], round(1e3*d["config"]["pairs_per_step"]/d["e2e"]["value"],3))
except Exception as e: print("$v", "ERR", e)
PY
done
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_timing.so timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_g6_timing.txt 2> gpurun_out/r2_g6_timing.err
timeout 600 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
