"""Aggregate an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X) per kernel: launches, total ms, share.
Usage: python tools/launch_summary.py launches.csv"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    name = re.sub(r"\(.*", "", r[ik]).strip()
    name = re.sub(r"^void ", "", name)
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].strip(), 1e-6)
    tot[name] += v * scale; cnt[name] += 1
all_ms = sum(tot.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.2f} | {tot[k] / all_ms:.3f} |")
print(f"| all | {sum(cnt.values())} | {all_ms:.2f} | 1.000 |")
