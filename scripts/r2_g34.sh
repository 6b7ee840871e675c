mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_broadphase.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/g34_tests_a.log 2>&1; echo "tests_a rc=$?"; tail -3 gpurun_out/g34_tests_a.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu $ARGS > gpurun_out/g34_$name.json 2> gpurun_out/g34_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/g34_$name.json").read().strip().splitlines()[-1])
    st=d["roofline"].get("stages_ms",{})
    print("$name", round(d["ms_per_step"],3), "e2e", round(d.get("e2e",{}).get("ms_per_step",0) or 0,3), {k:st.get(k) for k in ("overlap","pair_sort","gjk","epa")})
except Exception as e: print("$name", "ERR", e)
PY
}
ARGS="--steps 10 --warmup 3"
run c3_rows PK_X=1
run c3_radix PK_PAIR_RADIX=1
ARGS="--workload c5 --steps 5 --warmup 2"
run c5_rows PK_X=1
run c5_radix PK_PAIR_RADIX=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"overlap|pair_rows|tile_sum" --clock-control none --csv --log-file gpurun_out/g34_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/g34_ncu.csv | awk -F'","' '{print substr($5,1,28), $(NF-2), $NF}' | tail -12
timeout 1500 python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_fullsize.py --ignore=tests/test_gpu_broadphase.py > gpurun_out/g34_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/g34_tests.log
