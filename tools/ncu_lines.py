"""Aggregate an ncu report's source page by CUDA source line: share of issued warp instructions, active lanes per
instruction and stall samples.  Usage: python tools/ncu_lines.py report.ncu-rep [launch-index] [top-n]
(runs `ncu -i ... --page source --csv --print-source cuda,sass`; works without a GPU)."""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel launch: each launch starts with a "Function Name"/"Kernel Name" block sequence; files repeat inside a launch
launches, cur, seen_files = [], [], set()
fpath = None
for r in rows:
    if r and r[0] == "File Path":
        if r[1] in seen_files:
            launches.append(cur); cur = []; seen_files = set()
        seen_files.add(r[1]); fpath = r[1]
        continue
    if r and r[0] in ("Function Name", "Kernel Name"):
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if r and r[0] not in ("",) and len(r) > 8:
        cur.append((fpath, r))
launches.append(cur)
print(f"{len(launches)} launch(es) in report; showing #{launch}")
L = launches[launch]
ix = {h: i for i, h in enumerate(hdr)}
iI, iT, iS = ix["Instructions Executed"], ix["Thread Instructions Executed"], ix["# Samples"]
tot = sum(int(r[iI]) for _, r in L if r[iI].isdigit()); totS = sum(int(r[iS]) for _, r in L if r[iS].isdigit())
totT = sum(int(r[iT]) for _, r in L if r[iT].isdigit())
print(f"warp inst {tot/1e6:.1f} M, thread inst {totT/1e6:.1f} M, lanes/inst {totT/max(tot,1):.2f}, samples {totS}")
L2 = sorted(L, key=lambda fr: -int(fr[1][iI]) if fr[1][iI].isdigit() else 0)
for f, r in L2[:topn]:
    n, t, s = int(r[iI]), int(r[iT]), int(r[iS])
    print(f"{n/tot*100:5.2f}% inst  {s/max(totS,1)*100:5.2f}% smp  lanes {t/max(n,1):5.1f}  {f.split('/')[-1]}:{r[0]:>4}  {r[1].strip()[:110]}")
