# C3 (one world, pairs sharded) and C5 (batched worlds) on 1/2/4/8 GPUs of one box; one JSON line each
mkdir -p gpurun_out
port=29600
for wl in c5 c3; do
  for n in 1 2 4 8; do
    port=$((port+1))
    out=gpurun_out/r2_scale_${wl}_n${n}.json
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $wl --steps 10 --warmup 3 --no-cpu > $out 2> ${out%.json}.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps 10 --warmup 3 --no-cpu > $out 2> ${out%.json}.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("$out").read().strip().splitlines()[-1])
    st=d.get("roofline",{}).get("stages_ms",{})
    print("$wl n=$n", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "epa", st.get("epa"), "gjk", st.get("gjk"), d["e2e"].get("stages_ms_last_step",{}).get("contact_allgather"))
except Exception as e: print("$wl n=$n ERR", e)
PY
  done
done
