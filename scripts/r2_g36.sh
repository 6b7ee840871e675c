mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_shim.py -x -q -m gpu > gpurun_out/g36_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/g36_tests.log; ./tests/cpp/comm_test | tail -6
show() { python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    e=d["e2e"].get("stages_ms_last_step",{})
    print("$2", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "e2e epa", e.get("epa"), "gjk", e.get("gjk"), "scan", e.get("hit_scan"), "fetch", e.get("fetch_d2h"), "gather", e.get("contact_allgather"))
except Exception as ex: print("$2 ERR", ex)
PY
}
for m in mirror nomirror; do
  if [ $m = nomirror ]; then export PK_NO_MIRROR=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/g36_n1_$m.json 2> gpurun_out/g36_n1_$m.err; show gpurun_out/g36_n1_$m.json "n1 $m"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/g36_n2_$m.json 2> gpurun_out/g36_n2_$m.err; show gpurun_out/g36_n2_$m.json "n2 $m"
done
