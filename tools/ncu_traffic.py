"""profiles/ncu_traffic.json from `ncu --set full` captures of THIS tree (bench.py reads it for roofline.traffic):
    python tools/ncu_traffic.py <class>=<report.ncu-rep> ...
Each capture is the first matching launch of one bs8 256x320 step (tools/profile_step.py): the Cin = 180 full-resolution layer
denseBlocksUp.4.layers.3 for the DenseLayer classes.  `algorithmic_bytes` is that same launch's share of bench.py's byte model."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 16 * 256 * 320                     # pixels of the captured launch: both images of 8 pairs at 256x320
CIN, COUT = 180, 12
C8 = (CIN + 7) // 8 * 8
# the same per-launch figures as bench.py's conv_bytes_per_image (design_* entries for the two backward kernels: the data
# gradient also writes the bf16 operand planes, the weight-gradient GEMM reads them)
ALG = {"conv_dense_dgrad": (4.0 * (3 * CIN + COUT) + 2.0 * C8 + 2.0 * 16) * P, "conv_dense_wgrad": (2.0 * C8 + 2.0 * 16) * P,
       "conv_dense_fwd": 4.0 * (CIN + COUT) * P}
out = {}
for arg in sys.argv[1:]:
    cls, rep = arg.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]

    def val(k):
        v = float(r[h.index(k)].replace(",", ""))
        unit = u[h.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(unit, 1.0)

    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out[cls] = {"launch": "denseBlocksUp.4.layers.3 (Cin 180, 16 x 256x320), kernel " + r[h.index("Kernel Name")].split("(")[0],
                "duration_us": round(val("gpu__time_duration.sum"), 1), "dram_bytes": dram, "algorithmic_bytes": ALG[cls],
                "traffic_over_algorithmic": round(dram / ALG[cls], 3), "report": os.path.basename(rep),
                "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()}
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
