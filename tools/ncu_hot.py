"""Top stall sites of one kernel from an ncu report: python tools/ncu_hot.py <report.ncu-rep> [N]
Uses `ncu --page source --csv --print-source sass`: per-SASS-instruction warp-stall samples with their dominant reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r["# Samples"] or 0) for r in rows)
print(f"{rep}: {len(rows)} SASS instructions, {tot} samples")
agg = {k: sum(int(r[k] or 0) for r in rows) for k in reasons}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    n = int(r["# Samples"] or 0)
    rs = sorted(((int(r[k] or 0), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"{i:5d} {100 * n / max(tot, 1):5.1f}%  {r['Source'][:90]:90s} {rs[0][1]}:{rs[0][0]} {rs[1][1]}:{rs[1][0]}")
