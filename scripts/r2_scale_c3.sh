# C3 (one world, pairs sharded, pk_comm_* all-gathers) on 2/4/8 GPUs of one box; one JSON line each
mkdir -p gpurun_out
port=29650
for n in 2 4 8; do
  port=$((port+1))
  out=gpurun_out/r2_scale_c3_n${n}.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload c3 --steps 10 --warmup 3 --no-cpu > $out 2> ${out%.json}.err
  python - <<PY
import json
try:
    d=json.loads(open("$out").read().strip().splitlines()[-1])
    st=d.get("roofline",{}).get("stages_ms",{})
    print("c3 n=$n", round(d["ms_per_step"],3), "stage sum", round(sum(st.values()),3), "e2e", round(d["e2e"]["ms_per_step"],3), "epa", st.get("epa"), "gather", d["e2e"].get("stages_ms_last_step",{}).get("contact_allgather"))
except Exception as e: print("c3 n=$n ERR", e)
PY
done
