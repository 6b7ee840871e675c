mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/g35_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/g35_tests.log
for v in new sync; do
  if [ $v = sync ]; then export PK_SYNC_PAIRS=1; fi
  timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/g35_c3_$v.json 2> gpurun_out/g35_c3_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/g35_c3_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["e2e"].get("stages_ms_last_step"))
PY
done
