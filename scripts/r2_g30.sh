# FP32 GJK filter + device-side pair count: parity first, then A/B on C3 / C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_broadphase.py -x -q -m gpu > gpurun_out/g30_tests_a.log 2>&1; echo "tests_a rc=$?"; tail -5 gpurun_out/g30_tests_a.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu $ARGS > gpurun_out/g30_$name.json 2> gpurun_out/g30_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/g30_$name.json").read().strip().splitlines()[-1])
    st=d["roofline"].get("stages_ms",{})
    print("$name", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("ms_per_step"), {k:st.get(k) for k in ("overlap","pair_sort","gjk","scan","epa")}, d["config"].get("pairs"), d["config"].get("contacts"))
except Exception as e: print("$name", "ERR", e)
PY
}
ARGS=""
run c3_new PK_X=1
run c3_exact PK_GJK_EXACT_PREFILTER=1
run c3_sync PK_SYNC_PAIRS=1
ARGS="--workload c4"
run c4_new PK_X=1
run c4_exact PK_GJK_EXACT_PREFILTER=1
ARGS="--workload c4 --pairs 10000000 --steps 3 --warmup 1"
run c4big_new PK_X=1
run c4big_exact PK_GJK_EXACT_PREFILTER=1
timeout 1500 python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_narrowphase.py --ignore=tests/test_gpu_broadphase.py > gpurun_out/g30_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/g30_tests.log
