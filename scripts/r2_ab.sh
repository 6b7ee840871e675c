# A/B of library variants on the C3 step: prints ms/step and the EPA stage for every physkit_b200/libpk_v_*.so
mkdir -p gpurun_out
for lib in physkit_b200/libpk_collide.so physkit_b200/libpk_v_*.so; do
  v=$(basename $lib .so)
  PK_COLLIDE_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu ${PK_AB_ARGS} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    st=d["roofline"]["stages_ms"]
    print("$v", round(d["ms_per_step"],3), "epa", st.get("epa"), "gjk", st.get("gjk"), "sort", st.get("pair_sort"), "overlap", st.get("overlap"))
except Exception as e: print("$v", "ERR", e)
PY
done
