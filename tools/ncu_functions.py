"""Aggregate an ncu report's source page per inlined device function of csrc/pt_device.cuh / csrc/ptgpu.cu: share of issued warp
instructions, share of stall samples, active lanes per instruction.  Usage: python tools/ncu_functions.py report.ncu-rep
The report must come from the sources in the working tree (line numbers are matched against them).  Works without a GPU."""
import bisect, collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fpath, per = None, None, collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if r and r[0] == "File Path":
        fpath = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        n, t, s = r[ix["Instructions Executed"]], r[ix["Thread Instructions Executed"]], r[ix["# Samples"]]
        if n.isdigit():
            k = (os.path.basename(fpath), int(r[0]))
            per[k][0] += int(n) if n.isdigit() else 0; per[k][1] += int(t) if t.isdigit() else 0; per[k][2] += int(s) if s.isdigit() else 0
tot = sum(v[0] for v in per.values()); totT = sum(v[1] for v in per.values()); totS = sum(v[2] for v in per.values())
print(f"warp instructions {tot / 1e6:.1f} M, active lanes per instruction {totT / max(tot, 1):.2f}, stall samples {totS}")
funcs = {}
for name in ("pt_device.cuh", "ptgpu.cu"):
    src = open(os.path.join(ROOT, "ptsharp_b200", "csrc", name)).read().split("\n")
    fl = []
    for i, line in enumerate(src, 1):
        m = re.match(r"^(?:PT_D[NC]?|__global__|static)\s+[\w:<>\*&\s\(\),]+?\s+\*?(\w+)\(", line)
        if m:
            fl.append((i, m.group(1)))
    funcs[name] = fl
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in per.items():
    name = f
    if f in funcs:
        starts = [a for a, _ in funcs[f]]
        j = bisect.bisect_right(starts, l) - 1
        name = funcs[f][j][1] if j >= 0 else f
    for q in range(3):
        agg[name][q] += v[q]
print("| device function | issued warp instructions | stall samples | active lanes / instruction |\n|---|---|---|---|")
for name, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    if v[0] / max(tot, 1) >= 0.004:
        print(f"| `{name}` | {v[0] / tot * 100:.1f} % | {v[2] / max(totS, 1) * 100:.1f} % | {v[1] / max(v[0], 1):.1f} |")
