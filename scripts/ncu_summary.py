#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into profiles/ (tracked).

  python scripts/ncu_summary.py launches <launches.csv> <out.md>
  python scripts/ncu_summary.py kernel <report.ncu-rep> <kernel-substr> <out.md> [lib.so]
  python scripts/ncu_summary.py table <report.ncu-rep> <out.md>
  python scripts/ncu_summary.py traffic <report.ncu-rep> <kernel-substr> <stage> <workload> [profiles/ncu_traffic.json]

`kernel` needs ncu, cuobjdump and nvdisasm (all in the CUDA toolkit; no GPU).  Per-line stall
attribution joins ncu's SASS-level source page with nvdisasm's line table of the library."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {os.path.basename(path)}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |\n")
        f.write(f"| total | {sum(v[0] for v in agg.values())} | {tot:.3f} | |\n")
    print(open(out).read())


def kernel(rep, kern, out, lib):
    # first launch whose (template-argument-normalised) name contains `kern`, e.g. "epa_scan_kernel<0>"
    base = kern.split("<")[0]
    allraw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + base], capture_output=True, text=True).stdout
    names = [r[4] for r in list(csv.reader(io.StringIO(allraw)))[2:] if len(r) > 4]
    norm = [re.sub(r"\((?:bool|int)\)", "", n).replace("pk::", "") for n in names]
    skip = next(i for i, n in enumerate(norm) if kern in n)
    sel = ["-k", "regex:" + base, "--launch-skip", str(skip), "-c", "1"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rr[0], rr[1], rr[2]
    lines = [f"# ncu --set full: `{vals[4].split('(')[0]}`  ({os.path.basename(rep)})\n", f"grid {vals[8]} block {vals[7]}\n", "| metric | value | unit |\n|---|---:|---|\n"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| `{k}` | {vals[i]} | {units[i]} |\n")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
            try:
                stalls.append((float(vals[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(s[0] for s in stalls) or 1.0
    lines.append("\nStall reasons (PC samples): " + ", ".join(f"{n} {100 * v / tot:.1f} %" for v, n in sorted(stalls, reverse=True)[:6]) + "\n")
    # per-line attribution
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
        cub = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
        dis = subprocess.run(["nvdisasm", "-g", "-c"] + cub, capture_output=True, text=True).stdout.split("\n")
    mangled = re.sub(r"<(\d)>", lambda m: "ILb%sE" % m.group(1), kern)  # epa_scan_kernel<0> → epa_scan_kernelILb0E
    start = next((i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l and l.endswith(":")), None)
    off2line, cur = {}, None
    if start is not None:
        for l in dis[start + 1:]:
            if l.startswith("//-----"):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+\S", l)
            if m and cur:
                off2line[int(m.group(1), 16)] = cur
    sr = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(sr) if r and r[0] == "Address")
    sr = sr[hi - 1:]  # [kernel name row, header, instructions ...] of the first matching launch
    sh = sr[1]
    ia, isamp, iinst, ithr, ilsb = sh.index("Address"), sh.index("# Samples"), sh.index("Instructions Executed"), sh.index("Thread Instructions Executed"), sh.index("stall_long_sb")
    base = int(sr[2][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    t = [0, 0, 0, 0]
    for r in sr[2:]:
        if not r or not r[ia].startswith("0x"):
            break  # next launch
        key = off2line.get(int(r[ia], 16) - base, ("?", 0))
        v = [int(r[isamp]), int(r[iinst]), int(r[ithr]), int(r[ilsb])]
        for k in range(4):
            agg[key][k] += v[k]
            t[k] += v[k]
    cache = {}

    def text(f, n):
        p = os.path.join(ROOT, "physkit_b200", "csrc", f)
        if not os.path.exists(p):
            return ""
        if p not in cache:
            cache[p] = open(p).read().split("\n")
        return cache[p][n - 1].strip()[:100] if 0 < n <= len(cache[p]) else ""

    lines.append(f"\nWarp instructions {t[1]:,}; thread instructions {t[2]:,}; average active lanes {t[2] / max(t[1], 1):.2f} / 32.\n")
    lines.append("\n| samples | instr | lanes | long_sb | source |\n|---:|---:|---:|---:|---|\n")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        lines.append(f"| {100 * v[0] / t[0]:.1f} % | {100 * v[1] / t[1]:.1f} % | {v[2] / max(v[1], 1):.1f} | {100 * v[3] / max(v[0], 1):.0f} % | `{key[0]}:{key[1]}` `{text(*key)}` |\n")
    open(out, "w").write("".join(lines))
    print("".join(lines))


def traffic(rep, kern, stage, workload, out):
    """Adds / replaces the (stage, workload) record of profiles/ncu_traffic.json: DRAM bytes of one launch of `kern`
    in an `ncu --set full` report, with the hash of the sources it was built from (bench.py only trusts a record
    whose hash matches the current sources)."""
    import json

    sys.path.insert(0, ROOT)
    from bench import source_hash

    base = kern.split("<")[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + base], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units = rr[0], rr[1]
    row = next(r for r in rr[2:] if kern in re.sub(r"\((?:bool|int)\)", "", r[4]).replace("pk::", ""))

    def gb(name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]

    rec = {"stage": stage, "workload": workload, "kernel": kern, "report": os.path.basename(rep), "source_hash": source_hash(),
           "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
           "duration_ms": float(row[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns"), 1.0)}
    recs = []
    if os.path.exists(out):
        recs = [r for r in json.load(open(out)) if not (r["stage"] == stage and r["workload"] == workload)]
    recs.append(rec)
    json.dump(recs, open(out, "w"), indent=1)
    print(json.dumps(rec))


def table(rep, out):
    """One row per profiled launch of an `ncu --set full` report with the figures the roofline discussion uses."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units = rr[0], rr[1]

    def val(row, name, scale=None):
        if name not in hdr:
            return float("nan")
        i = hdr.index(name)
        v = float(row[i].replace(",", "")) if row[i] not in ("", "n/a") else float("nan")
        if scale:
            v *= scale.get(units[i], 1.0)
        return v

    B = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    T = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
    lines = [f"# ncu --set full, one steady-state step: {os.path.basename(rep)}\n\n",
             "Cold-cache, serialised launches (`--clock-control none`): compare shares and per-kernel figures, not the sum.\n\n",
             "| kernel | grid × block | ms | regs | lanes / 32 | warps active % | issue active % | DRAM read MB | DRAM write MB | DRAM % of peak | L2 hit % | FP64 pipe % |\n",
             "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n"]
    for row in rr[2:]:
        if len(row) < 10:
            continue
        name = re.sub(r"\((?:bool|int)\)", "", row[4].split("(")[0] if "<" not in row[4] else row[4][:row[4].index(">") + 1]).replace("void ", "").replace("pk::", "")
        lines.append(
            f"| `{name}` | {row[8]} × {row[7]} | {val(row, 'gpu__time_duration.sum', T):.3f} | {val(row, 'launch__registers_per_thread'):.0f} | "
            f"{val(row, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {val(row, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
            f"{val(row, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {val(row, 'dram__bytes_read.sum', B) / 1e6:.1f} | "
            f"{val(row, 'dram__bytes_write.sum', B) / 1e6:.1f} | {val(row, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
            f"{val(row, 'lts__t_sector_hit_rate.pct'):.1f} | {val(row, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} |\n")
    open(out, "w").write("".join(lines))
    print("".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "table":
        table(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6] if len(sys.argv) > 6 else os.path.join(ROOT, "profiles", "ncu_traffic.json"))
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "physkit_b200", "libpk_collide.so"))
