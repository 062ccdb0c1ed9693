"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
   python tools/summarize_ncu.py <tag> <launches.csv> [<report.ncu-rep> ...]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("endo::", "").replace("void ", "")
    return name[:100]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    out = ["| share | total ms | launches | kernel |", "|---:|---:|---:|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total < 0.0005:
            continue
        out.append(f"| {100 * t / total:.2f}% | {t / 1e6:.3f} | {n} | `{k}` |")
    return out, total


def details(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append(f"\n**`{short(r[idx['Kernel Name']])}`**\n")
        out.append("| metric | value |")
        out.append("|---|---:|")
        for k in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                out.append(f"| {k} | {r[idx[k]]} {units[idx[k]]} |")
    return out


def main():
    tag, launch_csv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    md = [f"# ncu summary `{tag}`", "",
          "Source: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of ONE optimisation step "
          "(`tools/profile_step.py`, bs8 256x320, fused pair forward) and `ncu --set full --clock-control none` "
          "captures of selected launches.  Times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", ""]
    tab, total = launches(launch_csv)
    md.append(f"## Launch list (sum {total / 1e6:.2f} ms)\n")
    md += tab
    for rep in reps:
        md.append(f"\n## `--set full` capture `{os.path.basename(rep)}`")
        md += details(rep)
    path = os.path.join(ROOT, "profiles", f"{tag}.md")
    with open(path, "w") as fh:
        fh.write("\n".join(md) + "\n")
    print(path)


if __name__ == "__main__":
    main()
