mkdir -p gpurun_out
for it in 0 1 2 4 8; do
  PK_GJK_FILTER_ITERS=$it PK_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/g31_it$it.json 2> gpurun_out/g31_it$it.err
  grep "to gjk_kernel" gpurun_out/g31_it$it.err | tail -1
  python - <<PY
import json
d=json.loads(open("gpurun_out/g31_it$it.json").read().strip().splitlines()[-1]); print("iters $it", round(d["ms_per_step"],3), d["roofline"]["stages_ms"].get("gjk"))
PY
done
PK_GJK_EXACT_PREFILTER=1 PK_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | grep "to gjk_kernel" | tail -1
for it in 0 8; do
PK_GJK_FILTER_ITERS=$it timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum -k regex:gjk --clock-control none --csv --log-file gpurun_out/g31_ncu_it$it.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/g31_ncu_it$it.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -12
done
