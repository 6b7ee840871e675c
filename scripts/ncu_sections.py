#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go, by SECTION of its source.

  python scripts/ncu_sections.py <report.ncu-rep> <kernel-substr> <kernel-file.cuh> name:lo-hi [name:lo-hi ...] [--lib lib.so] [--skip N]

Every SASS instruction of the first matching launch (ncu --page source) is attributed to the OUTERMOST line of
<kernel-file> in its inline chain (nvdisasm -gi), i.e. to the statement of the kernel body it was inlined into, and
the lines are summed over the given ranges.  Per section: share of warp instructions, of thread instructions,
average active lanes, share of the stall samples and the top stall reasons."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    lib = os.path.join(ROOT, "physkit_b200", "libpk_collide.so")
    skip = 0
    if "--lib" in args:
        i = args.index("--lib")
        lib = args[i + 1]
        del args[i:i + 2]
    split = []
    if "--split" in args:  # lines of the kernel file (call sites of the inlined body) to report separately
        i = args.index("--split")
        split = [int(x) for x in args[i + 1].split(",")]
        del args[i:i + 2]
    if "--skip" in args:
        i = args.index("--skip")
        skip = int(args[i + 1])
        del args[i:i + 2]
    rep, kern, kfile = args[0], args[1], os.path.basename(args[2])
    sections = []
    for a in args[3:]:
        name, rng = a.split(":")
        lo, hi = rng.split("-")
        sections.append((name, int(lo), int(hi)))
    base = kern.split("<")[0]
    allraw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + base], capture_output=True, text=True).stdout
    names = [r[4] for r in list(csv.reader(io.StringIO(allraw)))[2:] if len(r) > 4]
    norm = [re.sub(r"\((?:bool|int)\)", "", n).replace("pk::", "") for n in names]
    idx = [i for i, n in enumerate(norm) if kern in n][skip]
    sel = ["-k", "regex:" + base, "--launch-skip", str(idx), "-c", "1"]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
        cub = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
        dis = subprocess.run(["nvdisasm", "-gi", "-c"] + cub, capture_output=True, text=True).stdout.split("\n")
    m = re.match(r"(\w+)<(.*)>", kern)
    mangled = base
    if m:
        mangled = m.group(1) + "I" + "".join("Lb%sE" % x.strip() for x in m.group(2).split(",")) + "E"
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l and l.endswith(":"))
    off2line, chain = {}, []
    fresh = True
    for l in dis[start + 1:]:
        if l.startswith("//-----"):
            break
        if "//## File" in l:
            if fresh:
                chain = []
                fresh = False
            chain += [(os.path.basename(f), int(n)) for f, n in re.findall(r'"([^"]+)", line (\d+)', l)]
            continue
        mm = re.search(r"/\*([0-9a-f]{4,})\*/\s+\S", l)
        if mm:
            fresh = True
            # outermost line of the kernel file that lies inside one of the sections (a kernel that only calls an
            # inlined body has its own call line outermost: --split tags the instruction with it instead)
            inner = [n for f, n in chain if f == kfile]
            sec_lines = [n for n in inner if any(lo <= n <= hi_ for _, lo, hi_ in sections)]
            tag = next((str(n) for n in reversed(inner) if n in split), "")
            off2line[int(mm.group(1), 16)] = (sec_lines[-1] if sec_lines else 0, tag)
    sr = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(sr) if r and r[0] == "Address")
    sh = sr[hi]
    body = sr[hi + 1:]
    col = {n: sh.index(n) for n in ("Address", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    stall_cols = [(i, h[6:]) for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
    base_addr = int(body[0][col["Address"]], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot = [0, 0, 0]
    for r in body:
        if not r or not r[0].startswith("0x"):
            break
        line, tag = off2line.get(int(r[0], 16) - base_addr, (0, ""))
        sec = next((n for n, lo, hi_ in sections if lo <= line <= hi_), "other")
        v = [int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]])]
        for k in range(3):
            agg[(tag, sec)][k] += v[k]
            tot[k] += v[k]
        for i, n in stall_cols:
            agg[(tag, sec)][3][n] += int(r[i])
    print(f"# {kern}: {tot[1]:,} warp instructions, {tot[2]:,} thread instructions, {tot[2] / max(tot[1], 1):.2f} lanes, {tot[0]} samples\n")
    for tag in sorted({k[0] for k in agg}):
        if tag:
            print(f"\n## inlined at line {tag}\n")
        print("| section | lines | warp instr | thread instr | lanes | samples | top stalls |\n|---|---|---:|---:|---:|---:|---|")
        for name, lo, hi_ in sections + [("other", 0, 0)]:
            v = agg.get((tag, name))
            if not v:
                continue
            st = ", ".join(f"{n} {100 * c / max(sum(v[3].values()), 1):.0f}%" for n, c in v[3].most_common(3))
            print(f"| {name} | {lo}-{hi_} | {100 * v[1] / tot[1]:.1f} % | {100 * v[2] / tot[2]:.1f} % | {v[2] / max(v[1], 1):.1f} | {100 * v[0] / tot[0]:.1f} % | {st} |")


if __name__ == "__main__":
    main()
