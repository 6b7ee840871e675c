mkdir -p gpurun_out
for v in collide v_su; do
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_g11_c3_$v.json 2> gpurun_out/r2_g11_c3_$v.err
  PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_$v.so timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g11_c4_$v.json 2> gpurun_out/r2_g11_c4_$v.err
  python - <<PY
import json
for w in ("c3","c4"):
  try:
    d=json.loads(open("gpurun_out/r2_g11_%s_$v.json"%w).read().strip().splitlines()[-1])
    print(w, "$v", round(d["ms_per_step"],3), d["roofline"]["stages_ms"]["epa"])
  except Exception as e: print(w,"$v", "ERR", e)
PY
done
timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_g11_c5.json 2> gpurun_out/r2_g11_c5.err; tail -c 330 gpurun_out/r2_g11_c5.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epa_coop_kernel -c 1 -o gpurun_out/r2_g11_coop python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_g11_ncu.log 2>&1
