set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_g1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g1_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2_g1_tests.log
for v in collide v_old v_f4 v_f12 v_f16; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g1_$v.json 2> gpurun_out/r2_g1_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel -c 2 -o gpurun_out/r2_g1_coop python bench.py --steps 1 --warmup 1 > gpurun_out/r2_g1_ncu.log 2>&1
tail -3 gpurun_out/r2_g1_tests.log
for v in collide v_old v_f4 v_f12 v_f16; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g1_$v.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], d.get("stages_ms"), d["e2e"]["value"])
except Exception as e: print("$v", "ERR", e)
PY
done
