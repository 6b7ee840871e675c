mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_narrowphase.py tests/test_gpu_broadphase.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/g33_tests_a.log 2>&1; echo "tests_a rc=$?"; tail -3 gpurun_out/g33_tests_a.log
run() { name=$1; shift
  env "$@" PK_DEBUG=1 timeout 600 python bench.py --no-cpu $ARGS > gpurun_out/g33_$name.json 2> gpurun_out/g33_$name.err
  grep "to gjk_kernel" gpurun_out/g33_$name.err | tail -1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/g33_$name.json").read().strip().splitlines()[-1])
    st=d["roofline"].get("stages_ms",{})
    print("$name", round(d["ms_per_step"],3), {k:st.get(k) for k in ("overlap","pair_sort","gjk","epa")})
except Exception as e: print("$name", "ERR", e)
PY
}
ARGS="--steps 10 --warmup 3"
run c3 PK_X=1
ARGS="--workload c4 --steps 5 --warmup 2"
run c4 PK_X=1
ARGS="--workload c5 --steps 5 --warmup 2"
run c5 PK_X=1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum -k regex:gjk --clock-control none --csv --log-file gpurun_out/g33_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/g33_ncu.csv | awk -F'","' '{print substr($5,1,24), $(NF-2), $NF}' | tail -6
