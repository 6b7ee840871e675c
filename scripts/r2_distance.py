"""Device time of pk_gjk_distance_batch_device over the candidate pairs of BASELINE C3 (1 M bodies, 14.25 M pairs) and
over BASELINE C4's hull pairs: one line of JSON per workload → gpurun_out/r2_distance.json (copied to profiles/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import physkit_b200 as pk  # noqa: E402
from scenes import scene_c3, scene_c4  # noqa: E402

out = []


def timed(ctx, d_a, d_b, n, reps=5):
    d_out, d_sep = ctx.device_alloc(64 * n), ctx.device_alloc(n)
    ms = [ctx.gjk_distance_batch_device(d_a, d_b, n, d_out, d_sep) for _ in range(reps + 2)][2:]
    sep = np.zeros(n, dtype=np.uint8)
    ctx.d2h(sep, d_sep)
    ctx.device_free(d_out)
    ctx.device_free(d_sep)
    return float(np.median(ms)), int(sep.sum())


side = int(os.environ.get("PK_SIDE", "100"))
sc = scene_c3(side=side)
n = sc.n
ctx = pk.Context(n, int(16.5 * n * 1.3) + 4096, mode=pk.MODE_WORLD, max_shapes=len(sc.shapes), max_contacts=int(16.5 * n * 0.33) + 4096)
ctx.add_shapes(sc.shapes)
ctx.resize(n)
ctx.upload(sc.pos, sc.quat, np.zeros_like(sc.pos), sc.shape_id, sc.flags)
ctx.collide_resident()
ctx.update_pose(sc.pos + 0.05, None, None)
r = ctx.collide_resident()
ctx.fetch()
keys = ctx.pairs()
pa = (keys >> np.uint64(32)).astype(np.uint32)
pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
d_a, d_b = ctx.device_alloc(4 * len(pa)), ctx.device_alloc(4 * len(pa))
ctx.h2d(d_a, pa)
ctx.h2d(d_b, pb)
ms, nsep = timed(ctx, d_a, d_b, len(pa))
# algorithmic bytes per pair: 8 index + 112 poses + 128 shape records + 64 record + 1 flag
out.append(dict(workload=f"c3 side {side}", bodies=n, pairs=len(pa), contacts=int(r.num_contacts), separated=nsep, ms=ms,
                pairs_per_s=len(pa) / ms * 1e3, alg_gbs=len(pa) * 313 / ms / 1e6))
ctx.close()

if os.environ.get("PK_C4_PAIRS") == "0":
    print(json.dumps(out[0]))
    sys.exit(0)
sc, pa, pb = scene_c4(n_pairs=int(os.environ.get("PK_C4_PAIRS", "2000000")), n_hulls=1024)
nh = sum(len(s[1]) for s in sc.shapes if s[0] == "hull")
ctx = pk.Context(sc.n, 16, mode=pk.MODE_QUERY, max_shapes=len(sc.shapes), max_hull_vertices=nh + 8)
ctx.add_shapes(sc.shapes)
ctx.resize(sc.n)
ctx.upload(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
d_a, d_b = ctx.device_alloc(4 * len(pa)), ctx.device_alloc(4 * len(pa))
ctx.h2d(d_a, pa)
ctx.h2d(d_b, pb)
ms, nsep = timed(ctx, d_a, d_b, len(pa))
out.append(dict(workload="c4 hull pairs", bodies=sc.n, pairs=len(pa), separated=nsep, ms=ms, pairs_per_s=len(pa) / ms * 1e3))
ctx.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r2_distance.json"), "w") as f:
    for o in out:
        f.write(json.dumps(o) + "\n")
        print(json.dumps(o))
