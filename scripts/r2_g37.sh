mkdir -p gpurun_out
for side in 50 100; do
PK_COLLIDE_LIB=$PWD/physkit_b200/libpk_v_timing.so PK_NO_E2E=1 timeout 300 python bench.py --side $side --steps 1 --warmup 1 --no-cpu > gpurun_out/g37_$side.out 2> gpurun_out/g37_$side.err
python - <<PY
import numpy as np
rows=[l.split() for l in open("gpurun_out/g37_$side.out") if l.startswith("[ec]")]
t0=min(int(r[2]) for r in rows); ends=np.array(sorted(int(r[3])-t0 for r in rows))/1e6; st=np.array([int(r[2])-t0 for r in rows])/1e6
print("side $side warps", len(rows), "start max", st.max(), "retire quantiles", [round(float(np.quantile(ends,q)),3) for q in (0,.01,.1,.25,.5,.75,.9,.99,1)])
h,_=np.histogram(ends,bins=20); print(h.tolist())
PY
done
