mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    e=d["e2e"].get("stages_ms_last_step",{})
    print("$2", round(d["ms_per_step"],3), "resident epa", d["roofline"]["stages_ms"]["epa"], "| e2e", round(d["e2e"]["ms_per_step"],3), "e2e epa", e.get("epa"), "gjk", e.get("gjk"), "scan", e.get("hit_scan"))
except Exception as ex: print("$2 ERR", ex)
PY
}
for side in 100 50; do
timeout 600 python bench.py --side $side --steps 10 --warmup 3 --no-cpu > gpurun_out/g38_a.json 2>/dev/null; show gpurun_out/g38_a.json "side $side host-mirror"
PK_MIRROR_DEV=1 timeout 600 python bench.py --side $side --steps 10 --warmup 3 --no-cpu > gpurun_out/g38_b.json 2>/dev/null; show gpurun_out/g38_b.json "side $side device-mirror"
done
