"""Developer tool: the few counters profiles/k1_traffic.json holds, from ncu reports:  python tools/ncu_counts.py a.ncu-rep ..."""
import csv
import json
import subprocess
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def get(k):
        u, v = d[k]
        return float(v.replace(",", "")) * SCALE.get(u, 1)
    print(json.dumps({"report": rep, "kernel_ms": round(get("gpu__time_duration.sum"), 4),
                      "dram_bytes": int(get("dram__bytes_read.sum") + get("dram__bytes_write.sum")),
                      "warp_instructions": int(get("smsp__inst_executed.sum")),
                      "issue_active_pct": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                      "lanes_per_instruction": round(get("smsp__thread_inst_executed_per_inst_executed.ratio"), 1)}))
