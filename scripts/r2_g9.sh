mkdir -p gpurun_out
for v in collide v_su v_pf v_su_pf v_r1; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g9_c4_$v.json 2> gpurun_out/r2_g9_c4_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g9_c4_$v.json").read().strip().splitlines()[-1])
    print("c4 $v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"])
except Exception as e: print("$v", "ERR", e)
PY
done
for v in v_su v_su_pf; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g9_c3_$v.json 2> gpurun_out/r2_g9_c3_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g9_c3_$v.json").read().strip().splitlines()[-1])
    print("c3 $v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"])
except Exception as e: print("$v", "ERR", e)
PY
done
