mkdir -p gpurun_out
for v in collide v_nopipe; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g4_$v.json 2> gpurun_out/r2_g4_$v.err
done
for v in collide v_nopipe; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g4_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"], d["roofline"]["stages_ms"]["gjk"], round(1e3*d["config"]["pairs_per_step"]/d["e2e"]["value"],3))
except Exception as e: print("$v", "ERR", e)
PY
done
timeout 600 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_nopipe.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel -c 1 -o gpurun_out/r2_g4_coop python bench.py --steps 1 --warmup 1 > gpurun_out/r2_g4_ncu.log 2>&1
