#!/usr/bin/env python
"""bench.py — collision stage (broadphase + GJK/EPA) throughput on BASELINE config C3.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--side S]

Workload (BASELINE.json configs[2], "C3"): side³ bodies (default 100³ = 1 M) on a jittered lattice,
even ids analytic spheres, odd ids OBBs, physkit::world pair semantics (fat AABBs).  A "step" is one
pass of the hot path: bounds/fat update → LBVH build → self-overlap traversal → pair sort → GJK → EPA.
N > 1 (torchrun, one rank per GPU): every rank rebuilds the tree, traverses its own slice of the
sorted leaves, runs GJK/EPA on its own pairs, then ONE all-gather of contact records (NCCL).

Prints one JSON line (rank 0).  `value` = candidate-pair tests per second with inputs resident in
HBM; `e2e` = the same through pk_bodies_update_pose + pk_collide with pinned HOST buffers (H2D poses,
D2H pair keys + contacts inside the timed region).  --impl reference times the CPU oracle port
(oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "colliding-pair tests/sec (collision stage: LBVH broadphase + GJK/EPA), 1M-body C3 scene"
UNIT = "pair tests/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampling during the timed region (recipe's clocks line)."""

    # started before the warm-up (nvidia-smi needs ≈0.2 s to deliver its first line, a short timed region would see none);
    # mark() brackets the timed region and stop() keeps the samples inside it — or, if the region was shorter than the
    # sampling period, those of warm-up + timed region (same workload), and says so in "window"
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def mark(self):
        if self.t0 is None:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        import datetime

        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    ts = None  # (unknown time stamp format: the sample still counts for warm-up + timed region)
                try:
                    rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        out["window"] = "timed region" if inside else "warm-up + timed region"
        if not inside:
            inside = [r for r in rows if r[0] is None or self.t1 is None or r[0] <= self.t1]
        sm, mx, reasons = [r[1] for r in inside], [r[2] for r in inside], set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_scene(side):
    from scenes import scene_c3

    return scene_c3(side=side)


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(side_sample, nthreads, steps=1, warmup=0):
    """Oracle port of the reference's collision stage on a bounded C3 sample: faithful incremental
    dynamic_bvh broadphase (1 thread, as in the reference) + gjk_epa per active pair (nthreads)."""
    import oracle

    sc = make_scene(side_sample)
    w = oracle.World(sc.shapes)
    p0 = sc.pos
    p1 = sc.pos + 0.05
    zero = np.zeros_like(p0)
    w.step(p0, sc.quat, zero, sc.shape_id, sc.flags)  # create_rigid for all (no pairs yet)
    w.step(p1, sc.quat, zero, sc.shape_id, sc.flags)  # every leaf re-inserted → pair set forms
    times, pairs = [], 0
    for s in range(warmup + steps):
        pos = p0 if s % 2 == 0 else p1
        t0 = time.perf_counter()
        w.step(pos, sc.quat, zero, sc.shape_id, sc.flags)
        keys = w.pairs()
        pa = (keys >> np.uint64(32)).astype(np.uint32)
        pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        t1 = time.perf_counter()
        hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, pos, sc.quat, sc.shape_id, pa, pb, nthreads=nthreads)
        t2 = time.perf_counter()
        if s >= warmup:
            times.append((t1 - t0, t2 - t1))
            pairs = len(keys)
    tb = float(np.mean([t[0] for t in times]))
    tn = float(np.mean([t[1] for t in times]))
    return dict(bodies=sc.n, pairs=pairs, s_broad=tb, s_narrow=tn, pairs_per_s=pairs / (tb + tn), contacts=int(hit.sum()))


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle

    oracle.build()
    cores = oracle.max_threads()
    side = args.ref_side
    r = cpu_sample(side, cores, steps=args.steps, warmup=args.warmup)
    ms = 1e3 * (r["s_broad"] + r["s_narrow"])
    sample = (f"C3 generator at side={side} ({r['bodies']} bodies, {r['pairs']} pairs/step): faithful dynamic_bvh "
              f"broadphase on 1 thread ({1e3 * r['s_broad']:.1f} ms) + gjk_epa on {cores} threads ({1e3 * r['s_narrow']:.1f} ms)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 bounded sample: {r['bodies']}-body jittered lattice, spheres+OBBs, world-mode pairs",
                   "bodies": r["bodies"], "pairs_per_step": r["pairs"]},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
STAGE_BYTES_DOC = {
    # algorithmic bytes per unit (SURVEY §8d / DESIGN.md §roofline)
    "bounds_fat": ("body", 180), "morton": ("body", 56), "body_sort": ("body", 68), "leaves": ("body", 20),
    "hierarchy_ropes": ("body", 148), "overlap": ("body+pair", (176, 8)), "pair_sort": ("pair", 144),
    "gjk": ("pair", 153), "hit_scan": ("pair", 5), "epa": ("hit", 232), "compact": ("hit", 0),
}


def source_hash():
    """Hash of the CUDA sources the library is built from: ties a committed ncu capture to the code it measured."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "physkit_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(stage, workload, n_gpus):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture summarised in profiles/ncu_traffic.json (scripts/ncu_summary.py traffic ...).
    None when there is no capture of this stage / workload, when the sources have changed since it was taken, or
    on more than one GPU (the capture is a single-GPU launch)."""
    if n_gpus != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            recs = json.load(f)
    except Exception:
        return None
    for r in recs:
        if r.get("stage") == stage and r.get("workload") == workload and r.get("source_hash") == source_hash():
            return int(r["dram_bytes_read"] + r["dram_bytes_write"])
    return None


def stage_bytes(name, n, pairs, hits):
    unit, b = STAGE_BYTES_DOC[name]
    if unit == "body":
        return n * b
    if unit == "pair":
        return pairs * b
    if unit == "hit":
        return hits * b
    return n * b[0] + pairs * b[1]


def run_ours(args, rank, world, local_rank):
    import torch

    import physkit_b200 as pk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the collision library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc = make_scene(args.side)
    n = sc.n
    est_pairs = int(16.5 * n / world) + 4096
    max_pairs = int(est_pairs * 1.3)
    max_contacts = int(max_pairs * 0.25) + 4096
    ctx = pk.Context(n, max_pairs, mode=pk.MODE_WORLD, device=local_rank, max_shapes=len(sc.shapes),
                     max_contacts=max_contacts, shard_rank=rank, shard_count=world)
    ctx.add_shapes(sc.shapes)
    ctx.resize(n)
    # per-step inputs live in pinned host memory (what a host engine would hand over every step)
    h_pos = [ctx.pinned_empty((n, 3), np.float64) for _ in range(2)]
    h_quat = ctx.pinned_empty((n, 4), np.float64)
    h_disp = ctx.pinned_empty((n, 3), np.float64)
    h_pos[0][:] = sc.pos
    h_pos[1][:] = sc.pos + 0.05
    h_quat[:] = sc.quat
    h_disp[:] = 0.0
    ctx.upload(h_pos[0], h_quat, h_disp, sc.shape_id, sc.flags)
    r0 = ctx.collide_resident()
    assert r0.num_pairs == 0, "first step must yield no pairs (reference first-step quirk)"
    ctx.update_pose(h_pos[1], None, None)
    r1 = ctx.collide_resident()
    log(f"[rank {rank}] bodies {n}  pairs {r1.num_pairs}  contacts {r1.num_contacts}  moved {r1.num_moved}  "
        f"device ms {r1.ms_total:.3f}")

    # one world over several ranks: the exchange runs inside the library (pk_comm_*: NCCL all-gathers on the context's
    # stream, straight from the contact buffer); torch.distributed only carries the communicator id and the timings
    first, count = 0, n
    if world > 1:
        cid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            cid.copy_(torch.from_numpy(pk.Context.comm_get_id()))
        dist.broadcast(cid, 0)
        ctx.comm_init(cid.cpu().numpy(), rank, world)
        first, count = ctx.comm_pose_slice()

    gather_ms = []

    def exchange():
        if world == 1:
            return None
        g = ctx.comm_allgather_contacts()
        gather_ms.append(g[3])
        return g

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident (kernel-path) timing --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        ctx.collide_resident()
        exchange()
    stage_acc = {}
    launches = 0
    dev_ms = 0.0
    sync_all()
    sampler.mark()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ctx.collide_resident()
        exchange()
        st, ln = ctx.stage_times()
        for k, v in st.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        launches += ln
        dev_ms += res.ms_total
    sync_all()
    t1 = time.perf_counter()
    sampler.mark()
    clocks = sampler.stop()
    wall_ms = 1e3 * (t1 - t0) / args.steps
    pairs, hits, contacts = int(res.num_pairs), int(res.gjk_hits), int(res.num_contacts)

    # ---- end-to-end timing: host poses in, pair keys + contacts out ------------------------------
    # every rank uploads the poses of ITS slice of the bodies (1/N of the per-step H2D) and gets the rest over NVLink
    sl = slice(first, first + count)

    def e2e_step(s):
        ctx.update_pose(h_pos[s % 2][sl], h_quat[sl], h_disp[sl], first=first, count=count)
        if world > 1:
            ctx.comm_allgather_poses()
        r = ctx.collide()
        exchange()
        return r

    for s in range(min(args.warmup, 3)):
        e2e_step(s)
    sync_all()
    gather_ms.clear()
    e0 = time.perf_counter()
    for s in range(args.steps):
        res_e = e2e_step(s)
    sync_all()
    e1 = time.perf_counter()
    e2e_ms = 1e3 * (e1 - e0) / args.steps
    e2e_stages = {k: round(v, 4) for k, v in ctx.stage_times()[0].items()}  # device times of the last end-to-end step
    if gather_ms:
        e2e_stages["contact_allgather"] = round(float(np.mean(gather_ms)), 4)
    h2d = count * (3 + 4 + 3) * 8
    d2h = int(res_e.num_pairs) * 8 + int(res_e.num_contacts) * 88

    # ---- optional: the rows downstream of the stage, not part of the headline -----------------
    manifold_info = None
    if args.manifolds and world == 1:
        ctx.manifolds_enable(min(max_contacts, 4 * max(contacts, 1) + 1024))
        ctx.dynamics_enable()
        rng = np.random.default_rng(1)
        ctx.dynamics_upload(rng.uniform(-1, 1, (n, 3)), rng.uniform(-1, 1, (n, 3)), np.full(n, 1.0), np.tile(np.eye(3).ravel(), (n, 1)))
        ms, rows_ms, nrows = [], [], 0
        for s in range(6):
            ctx.update_pose(h_pos[s % 2], h_quat, h_disp)
            ctx.collide_resident()
            ms.append(ctx.manifolds_update())
            nrows = ctx.contact_rows_setup(1.0 / 60.0, 9.81)
            rows_ms.append(ctx.contact_rows_device_ms())
        manifold_info = {"manifolds": ms[-1][0], "ms_first_step": ms[0][3], "ms_steady": float(np.median([m[3] for m in ms[2:]])),
                         "contact_rows": nrows, "contact_rows_ms": float(np.median(rows_ms[2:]))}
        # integrator loops A + B (pk_integrate_*), timed with events around both launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(ctx.stream())
        t_int = []
        for _ in range(5):
            ev0.record(st)
            ctx.integrate_velocities(1.0 / 60.0)
            ctx.integrate_positions(1.0 / 60.0)
            ev1.record(st)
            ev1.synchronize()
            t_int.append(ev0.elapsed_time(ev1))
        manifold_info["integrate_vel_pos_ms"] = float(np.median(t_int[1:]))
        # 1 M rays against the tree of the last step
        ctx.update_pose(h_pos[0], h_quat, h_disp)
        ctx.collide_resident()
        nr = 1_000_000
        o = rng.uniform(0.0, args.side * 0.8, (nr, 3))
        d = rng.uniform(-1, 1, (nr, 3))
        ray_entries = ctx.raycast(o, d, 0.5, capacity=max_pairs)
        manifold_info["raycast"] = {"rays": nr, "max_distance_m": 0.5, "entries": int(len(ray_entries)), "device_ms": ctx.raycast_device_ms()}
        ctx.raycast(o, d, 0.5, mode=pk.RAY_CLOSEST, capacity=nr)
        manifold_info["raycast"]["closest_device_ms"] = ctx.raycast_device_ms()

    # ---- reduce over ranks (max time, sum pairs) ---------------------------------------------------
    tot_pairs, tot_contacts, max_wall, max_e2e = pairs, contacts, wall_ms, e2e_ms
    if world > 1:
        t = torch.tensor([pairs, contacts, h2d, d2h], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot_pairs, tot_contacts, h2d, d2h = (int(x) for x in t.tolist())
        m = torch.tensor([wall_ms, e2e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        max_wall, max_e2e = float(m[0].item()), float(m[1].item())

    if rank == 0:
        peak, peak_src = measured_peak()
        stages = {k: v / args.steps for k, v in stage_acc.items()}
        table = []
        for k, ms in stages.items():
            if k == "fetch_d2h" or ms <= 0:
                continue
            b = stage_bytes(k, n, pairs, hits)
            table.append((k, ms, b, b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0))
        table.sort(key=lambda x: -x[1])
        log("stage                 ms/step   alg.MB   GB/s   frac-of-peak")
        for k, ms, b, gbs in table:
            log(f"  {k:18s} {ms:8.3f} {b / 1e6:9.1f} {gbs:7.1f}   {gbs / peak:6.3f}")
        dom = table[0]
        total_bytes = sum(t[2] for t in table)
        dev_step_ms = dev_ms / args.steps
        roofline = {
            "bound": "hbm", "kernel": dom[0], "achieved": dom[3], "peak": peak, "unit": "GB/s", "frac": dom[3] / peak,
            "traffic": ncu_traffic(dom[0], f"c3-side{args.side}", world), "peak_source": peak_src,
            "kernel_ms": dom[1], "kernel_share_of_step": dom[1] / max(dev_step_ms, 1e-9),
            "whole_step": {"alg_bytes": total_bytes, "device_ms": dev_step_ms,
                           "achieved": total_bytes / (dev_step_ms * 1e-3) / 1e9, "frac": total_bytes / (dev_step_ms * 1e-3) / 1e9 / peak},
            "stages_ms": {k: round(v, 4) for k, v in stages.items()},
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            import oracle

            oracle.build()
            c = cpu_sample(args.cpu_side, 1, steps=1, warmup=0)
            cpu = {"value": c["pairs_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": (f"C3 generator at side={args.cpu_side} ({c['bodies']} bodies, {c['pairs']} pairs): one steady-state "
                              f"world step of the oracle port, faithful dynamic_bvh broadphase {1e3 * c['s_broad']:.1f} ms + "
                              f"gjk_epa {1e3 * c['s_narrow']:.1f} ms, single thread like the reference")}
        line = {
            "metric": METRIC, "value": tot_pairs / (max_wall * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_wall, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: {n}-body jittered lattice (side {args.side}), analytic spheres + OBBs, world-mode fat-AABB pairs",
                       "bodies": n, "pairs_per_step": tot_pairs, "contacts_per_step": tot_contacts,
                       "epa_fallback_pairs": getattr(ctx, "epa_fallback", None),
                       **({"downstream_rows": manifold_info} if manifold_info else {}),
                       "sharding": ("pairs by sorted-leaf range, tree rebuilt per rank; pk_comm_*: each rank uploads 1/N of the poses + NCCL all-gather, "
                                    "one NCCL all-gather of contact records per step, both inside the library") if world > 1 else "none",
                       "l2": "inputs larger than L2 (per-step working set > 1 GB vs 126 MB L2); no explicit flush",
                       "timing": "wall clock around K synchronous steps bracketed by barrier+synchronize, max over ranks; "
                                 "per-kernel times from CUDA events on the library's stream",
                       "poses": "alternate P0/P1 inside the fat boxes (resting pile: no re-insertions after warm-up)"},
            "device_ms_per_step": dev_step_ms,
            "e2e": {"value": tot_pairs / (max_e2e * 1e-3), "unit": UNIT, "ms_per_step": max_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "stages_ms_last_step": e2e_stages},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_c4(args, rank, world, local_rank):
    """BASELINE C4: narrowphase only, random hull pairs from a 1024-hull library (V in 32..256)."""
    import torch

    import physkit_b200 as pk
    from scenes import scene_c4

    torch.cuda.set_device(local_rank)
    npairs = args.pairs // world
    sc, pa, pb = scene_c4(n_pairs=npairs, n_hulls=1024, seed=0x5EED0004 + rank)
    nh = sum(len(s[1]) for s in sc.shapes)
    ctx = pk.Context(sc.n, npairs, mode=pk.MODE_QUERY, device=local_rank, max_shapes=len(sc.shapes), max_contacts=npairs,
                     max_hull_vertices=nh + 8)
    ctx.add_shapes(sc.shapes)
    ctx.resize(sc.n)
    ctx.upload(sc.pos, sc.quat, None, sc.shape_id, sc.flags)
    d_a, d_b = ctx.device_alloc(4 * npairs), ctx.device_alloc(4 * npairs)
    d_out, d_hit = ctx.device_alloc(88 * npairs), ctx.device_alloc(npairs)
    ctx.h2d(d_a, pa)
    ctx.h2d(d_b, pb)
    for _ in range(args.warmup):
        ctx.gjk_epa_batch_device(d_a, d_b, npairs, d_out, d_hit)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = 0.0
    for _ in range(args.steps):
        dev += ctx.gjk_epa_batch_device(d_a, d_b, npairs, d_out, d_hit)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    st, launches = ctx.stage_times()
    hit = np.empty(npairs, np.uint8)
    ctx.d2h(hit, d_hit)
    mean_v = float(np.mean([len(s[1]) for s in sc.shapes]))
    staged = npairs * (24.0 * 2 * mean_v + 153) + int(hit.sum()) * 88
    peak, peak_src = measured_peak()
    if rank == 0:
        line = {"metric": "colliding-pair tests/sec (GJK/EPA batch, convex hulls 32-256 verts)", "value": npairs * world / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4: {npairs * world} random hull pairs, 1024-hull library", "hit_rate": float(hit.mean())},
                "device_ms_per_step": dev / args.steps, "gpu_launches": launches * args.steps,
                "roofline": {"bound": "hbm", "kernel": "gjk+epa (staged hull bytes)", "achieved": staged / (ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": staged / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "stages_ms": {k: round(v, 4) for k, v in st.items() if v > 0}}}
        print(json.dumps(line), flush=True)
    ctx.close()


def c5_worlds(first, count):
    """Worlds [first, first+count) of BASELINE C5: a C1-style 8x8x8 pile of 8-vertex box hulls + ground per world,
    jitter seeded per world (0x5EED0005 + k), so a rank's worlds do not depend on how many ranks there are."""
    from scenes import SplitMix64, scene_c1

    base = scene_c1(side=8, spacing=0.97)
    per = base.n
    pos = np.concatenate([base.pos + SplitMix64(0x5EED0005 + first + k).uniform(-0.03, 0.03, per, 3) for k in range(count)])
    quat = np.tile(base.quat, (count, 1))
    sid = np.tile(base.shape_id, count)
    flags = np.tile(base.flags, count)
    wid = np.repeat(np.arange(count, dtype=np.uint32), per)
    return base, pos, quat, sid, flags, wid


def c5_partition(worlds, world_size):
    """Block partition of the worlds over the ranks: [(first, count)] per rank, a partition of range(worlds)."""
    q, r = divmod(worlds, world_size)
    out, first = [], 0
    for k in range(world_size):
        c = q + (1 if k < r else 0)
        out.append((first, c))
        first += c
    return out


def cpu_sample_c5(nworlds):
    """Oracle port (1 thread, as the reference) on the first `nworlds` C5 worlds, one steady-state step each."""
    import oracle

    base, pos, quat, sid, flags, wid = c5_worlds(0, nworlds)
    per = base.n
    tot_pairs, tb, tn = 0, 0.0, 0.0
    for k in range(nworlds):
        sl = slice(k * per, (k + 1) * per)
        w = oracle.World(base.shapes)
        p0 = pos[sl]
        p1 = p0.copy()
        p1[:, 1] -= 0.04
        zero = np.zeros_like(p0)
        w.step(p0, quat[sl], zero, base.shape_id, base.flags)
        w.step(p1, quat[sl], zero, base.shape_id, base.flags)
        t0 = time.perf_counter()
        w.step(p0, quat[sl], zero, base.shape_id, base.flags)
        keys = w.pairs()
        pa = (keys >> np.uint64(32)).astype(np.uint32)
        pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        t1 = time.perf_counter()
        oracle.gjk_epa_pairs(base.shapes, p0, quat[sl], base.shape_id, pa, pb, nthreads=1)
        t2 = time.perf_counter()
        tot_pairs += len(keys)
        tb += t1 - t0
        tn += t2 - t1
    return dict(worlds=nworlds, pairs=tot_pairs, s_broad=tb, s_narrow=tn, pairs_per_s=tot_pairs / (tb + tn))


def run_c5(args, rank, world, local_rank):
    """BASELINE C5: independent 513-body worlds batched in one context per rank; the worlds are block-partitioned
    over the ranks (replicas of the pipeline, no collective on the data path)."""
    import torch

    import physkit_b200 as pk

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    first, nw = c5_partition(args.worlds, world)[rank]
    base, pos, quat, sid, flags, wid = c5_worlds(first, nw)
    per = base.n
    n = len(pos)
    max_pairs, max_contacts = int(14 * n), int(7 * n)
    ctx = pk.Context(n, max_pairs, mode=pk.MODE_WORLD, device=local_rank, max_shapes=8, max_contacts=max_contacts,
                     max_hull_vertices=64, num_worlds=nw)
    ctx.add_shapes(base.shapes)
    ctx.resize(n)
    h_pos = [ctx.pinned_empty((n, 3), np.float64) for _ in range(2)]
    h_quat = ctx.pinned_empty((n, 4), np.float64)
    h_disp = ctx.pinned_empty((n, 3), np.float64)
    h_pos[0][:] = pos
    h_pos[1][:] = pos
    h_pos[1][:, 1] -= 0.04
    h_quat[:] = quat
    h_disp[:] = 0.0
    ctx.upload(h_pos[0], h_quat, h_disp, sid, flags, wid)
    ctx.collide_resident()
    ctx.update_pose(h_pos[1], None, None)
    r = ctx.collide_resident()
    log(f"[rank {rank}] worlds {first}..{first + nw - 1}  bodies {n}  pairs {r.num_pairs}  contacts {r.num_contacts}")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        ctx.collide_resident()
    stage_acc, launches, dev_ms = {}, 0, 0.0
    sync_all()
    sampler.mark()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = ctx.collide_resident()
        st, ln = ctx.stage_times()
        for k, v in st.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        launches += ln
        dev_ms += r.ms_total
    sync_all()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    sampler.mark()
    clocks = sampler.stop()
    pairs, hits, contacts = int(r.num_pairs), int(r.gjk_hits), int(r.num_contacts)
    # end to end: every rank uploads the poses of ITS worlds from pinned host memory and reads its pair keys and
    # contact records back
    for s in range(3):
        ctx.update_pose(h_pos[s % 2], None, h_disp)
        ctx.collide()
    sync_all()
    e0 = time.perf_counter()
    for s in range(args.steps):
        ctx.update_pose(h_pos[s % 2], h_quat, h_disp)
        re_ = ctx.collide()
    sync_all()
    e2e_ms = 1e3 * (time.perf_counter() - e0) / args.steps
    h2d = n * (3 + 4 + 3) * 8
    d2h = int(re_.num_pairs) * 8 + int(re_.num_contacts) * 88
    tot = [float(pairs), float(contacts), float(h2d), float(d2h)]
    mx = [ms, e2e_ms]
    if world > 1:
        t = torch.tensor(tot, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = [float(x) for x in t.tolist()]
        m = torch.tensor(mx, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        mx = [float(x) for x in m.tolist()]
    if rank == 0:
        peak, peak_src = measured_peak()
        stages = {k: v / args.steps for k, v in stage_acc.items()}
        table = sorted(((k, v, stage_bytes(k, n, pairs, hits)) for k, v in stages.items() if k != "fetch_d2h" and v > 0), key=lambda x: -x[1])
        dom = table[0]
        total_bytes = sum(t[2] for t in table)
        dev_step = dev_ms / args.steps
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(dom[0], f"c5-{nw}worlds", world),
                    "peak_source": peak_src, "kernel_ms": dom[1], "kernel_share_of_step": dom[1] / max(dev_step, 1e-9),
                    "whole_step": {"alg_bytes": total_bytes, "device_ms": dev_step, "achieved": total_bytes / (dev_step * 1e-3) / 1e9,
                                   "frac": total_bytes / (dev_step * 1e-3) / 1e9 / peak},
                    "stages_ms": {k: round(v, 4) for k, v in stages.items()},
                    "note": "rank 0's launches; algorithmic bytes as for C3 (SURVEY 8d) with this rank's bodies / pairs / hits"}
        cpu = None
        if world == 1 and not args.no_cpu:
            import oracle

            oracle.build()
            c = cpu_sample_c5(64)
            cpu = {"value": c["pairs_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": (f"the first {c['worlds']} C5 worlds ({c['pairs']} pairs), one steady-state world step each of the oracle port: "
                              f"dynamic_bvh broadphase {1e3 * c['s_broad']:.1f} ms + gjk_epa {1e3 * c['s_narrow']:.1f} ms, single thread like the reference")}
        line = {"metric": "colliding-pair tests/sec (batched independent worlds)", "value": tot[0] / (mx[0] * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mx[0], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5: {args.worlds} worlds x {per} bodies (8-vertex box hulls + ground), worlds block-partitioned over the ranks, no collective",
                           "worlds_per_rank": [c for _, c in c5_partition(args.worlds, world)], "pairs_per_step": int(tot[0]),
                           "contacts_per_step": int(tot[1]), "worlds_per_s": args.worlds / (mx[0] * 1e-3),
                           "l2": "per-step working set of a rank exceeds L2 at <= 4 ranks (no explicit flush); 512 worlds per rank: 0.3 GB",
                           "timing": "wall clock around K synchronous steps bracketed by barrier+synchronize, max over ranks"},
                "device_ms_per_step": dev_step,
                "e2e": {"value": tot[0] / (mx[1] * 1e-3), "unit": UNIT, "ms_per_step": mx[1], "h2d_bytes_per_step": int(tot[2]),
                        "d2h_bytes_per_step": int(tot[3])},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def c1_trajectory(steps, dt=1.0 / 60.0):
    """BASELINE C1: ground + 10x10x10 unit boxes (8-vertex hulls, as every body of a physkit::world demo) dropped from a
    lattice with 0.2 m gaps.  The sweep that would stop them is out of scope, so the drop is prescribed: free fall under
    gravity until a box reaches its place in the settled pile (layers 0.98 m apart: resting contacts of 2 cm).
    → scene, pos[steps, n, 3], disp[steps, n, 3] (disp = the step's motion, what world::step_impl passes as vel*dt)."""
    from scenes import scene_c1

    sc = scene_c1(side=10)
    n = sc.n
    layer = np.round((sc.pos[:, 1] - 1.0) / 1.2).astype(np.int64)
    rest = 0.5 + 0.98 * layer
    rest[0] = sc.pos[0, 1]
    pos = np.empty((steps, n, 3))
    disp = np.zeros((steps, n, 3))
    y = sc.pos[:, 1].copy()
    v = np.zeros(n)
    for k in range(steps):
        v[1:] -= 9.81 * dt
        ny = np.maximum(rest, y + v * dt)
        ny[0] = y[0]
        v[ny <= rest] = 0.0
        pos[k] = sc.pos
        pos[k, :, 1] = y
        disp[k, :, 1] = ny - y
        y = ny
    return sc, pos, disp


def run_c1(args, rank, world, local_rank):
    """BASELINE C1 (the reference's own CPU-runnable case): one 1001-body world, 600 steps at 60 Hz, every step through
    pk_bodies_update_pose + pk_collide with host buffers, as a host engine would drive it.  N > 1: replicas."""
    import torch

    import physkit_b200 as pk

    torch.cuda.set_device(local_rank)
    steps = args.steps if args.steps != 20 else 600
    sc, pos, disp = c1_trajectory(steps + args.warmup)
    n = sc.n
    ctx = pk.Context(n, 64 * n, mode=pk.MODE_WORLD, device=local_rank, max_shapes=8, max_contacts=32 * n, max_hull_vertices=64)
    ctx.add_shapes(sc.shapes)
    ctx.resize(n)
    h_pos, h_disp = ctx.pinned_empty((n, 3), np.float64), ctx.pinned_empty((n, 3), np.float64)
    ctx.upload(pos[0], sc.quat, disp[0], sc.shape_id, sc.flags)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for k in range(args.warmup):
        h_pos[:] = pos[k]
        h_disp[:] = disp[k]
        ctx.update_pose(h_pos, None, h_disp)
        ctx.collide()
    torch.cuda.synchronize()
    sampler.mark()
    pairs = contacts = launches = 0
    dev_ms = 0.0
    stage_acc = {}
    t0 = time.perf_counter()
    for k in range(args.warmup, args.warmup + steps):
        h_pos[:] = pos[k]
        h_disp[:] = disp[k]
        ctx.update_pose(h_pos, None, h_disp)
        r = ctx.collide()
        pairs += int(r.num_pairs)
        contacts += int(r.num_contacts)
        dev_ms += r.ms_total
        st, ln = ctx.stage_times()
        launches += ln
        for key, val in st.items():
            stage_acc[key] = stage_acc.get(key, 0.0) + val
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sampler.mark()
    clocks = sampler.stop()
    last_keys = ctx.pairs()
    if rank == 0:
        peak, peak_src = measured_peak()
        stages = {k: v / steps for k, v in stage_acc.items()}
        hits = contacts
        table = sorted(((k, v, stage_bytes(k, n, pairs / steps, hits / steps)) for k, v in stages.items() if k != "fetch_d2h" and v > 0), key=lambda x: -x[1])
        dom = table[0]
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": dom[1],
                    "stages_ms": {k: round(v, 4) for k, v in stages.items()},
                    "note": "a 1001-body world is bound by launch latency (about 35 launches per step), not by any roof"}
        cpu = None
        if not args.no_cpu:
            import oracle

            oracle.build()
            w = oracle.World(sc.shapes)
            ck = min(steps + args.warmup, 120)
            cp, tb, tn = 0, 0.0, 0.0
            for k in range(ck):
                c0 = time.perf_counter()
                w.step(pos[k], sc.quat, disp[k], sc.shape_id, sc.flags)
                keys = w.pairs()
                c1 = time.perf_counter()
                oracle.gjk_epa_pairs(sc.shapes, pos[k], sc.quat, sc.shape_id, (keys >> np.uint64(32)).astype(np.uint32),
                                     (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), nthreads=1)
                c2 = time.perf_counter()
                if k >= args.warmup:
                    cp += len(keys)
                    tb += c1 - c0
                    tn += c2 - c1
            cpu = {"value": cp / max(tb + tn, 1e-12), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": (f"steps {args.warmup}..{ck - 1} of the same trajectory ({cp} pairs): oracle port of the collision stage, dynamic_bvh broadphase "
                              f"{1e3 * tb:.1f} ms + gjk_epa {1e3 * tn:.1f} ms, single thread like the reference"),
                   "ms_per_step": 1e3 * (tb + tn) / max(ck - args.warmup, 1)}
        line = {"metric": "colliding-pair tests/sec (PhysKit demo world, collision stage)", "value": pairs / max(dev_ms * 1e-3, 1e-12), "unit": UNIT,
                "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C1: ground + 1000 unit boxes (8-vertex hulls) dropped onto it, {steps} steps at 60 Hz, prescribed fall (no solver)",
                           "pairs_total": pairs, "contacts_total": contacts, "pairs_last_step": int(len(last_keys)),
                           "timing": "value: device time of the collision stage summed over the steps (CUDA events); e2e: wall clock of the loop "
                                     "pk_bodies_update_pose + pk_collide with pinned host buffers"},
                "e2e": {"value": pairs / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / steps, "h2d_bytes_per_step": n * 6 * 8,
                        "d2h_bytes_per_step": int((pairs * 8 + contacts * 88) / steps)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=100, help="C3 lattice side (bodies = side^3)")
    ap.add_argument("--cpu-side", type=int, default=50, help="lattice side of the cpu_baseline sample")
    ap.add_argument("--ref-side", type=int, default=40, help="lattice side of the --impl reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--manifolds", action="store_true", help="also time the downstream rows (manifolds, contact rows, integrator, ray casts; reported under config)")
    ap.add_argument("--workload", default="c3", choices=["c1", "c3", "c4", "c5"], help="c3 is the headline; c1/c4/c5 are the other BASELINE configs")
    ap.add_argument("--pairs", type=int, default=2_000_000, help="c4: number of hull pairs")
    ap.add_argument("--worlds", type=int, default=4096, help="c5: number of independent worlds")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        log(f"bench.py: --gpus {args.gpus} without torchrun: running rank 0 of 1 (launch with torch.distributed.run for N>1)")
    if args.workload == "c1":
        return run_c1(args, rank, world, local_rank)
    if args.workload == "c4":
        return run_c4(args, rank, world, local_rank)
    if args.workload == "c5":
        return run_c5(args, rank, world, local_rank)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
