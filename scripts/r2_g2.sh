mkdir -p gpurun_out
for v in stcs pf b2 b3 b4 k60; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_$v.so timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g2_$v.json 2> gpurun_out/r2_g2_$v.err
done
for v in stcs pf b2 b3 b4 k60; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_g2_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"], d["roofline"]["stages_ms"]["gjk"])
except Exception as e: print("$v", "ERR", e)
PY
done
