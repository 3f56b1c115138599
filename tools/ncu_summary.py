"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the committed summaries under profiles/:
a CSV of selected raw-page columns per captured launch and a JSON of DRAM bytes per launch that
bench.py reports as roofline.traffic.

    python tools/ncu_summary.py <tag> <report.ncu-rep> [<report2.ncu-rep> ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def raw_rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1]
    table, traffic = [], {}
    units_out = None
    for rep in sys.argv[2:]:
        hdr, units, rows = raw_rows(rep)
        idx = [hdr.index(c) if c in hdr else None for c in COLS]
        units_out = [units[i] if i is not None else "" for i in idx]
        for r in rows:
            table.append([r[i] if i is not None else "" for i in idx] + [os.path.basename(rep)])
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
            def val(col):
                i = hdr.index(col)
                return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            t = traffic.setdefault(name, {"launches_captured": 0, "dram_read": 0.0, "dram_write": 0.0, "duration_ms": 0.0,
                                          "report": os.path.basename(rep)[:-8], "grid": r[hdr.index("Grid Size")]})
            t["launches_captured"] += 1
            t["dram_read"] += val("dram__bytes_read.sum")
            t["dram_write"] += val("dram__bytes_write.sum")
            t["duration_ms"] += val("gpu__time_duration.sum")
    for t in traffic.values():
        n = t["launches_captured"]
        for k in ("dram_read", "dram_write", "duration_ms"):
            t[k] /= n
        t["dram_bytes_per_launch"] = t["dram_read"] + t["dram_write"]
    with open(os.path.join(ROOT, "profiles", f"{tag}_kernels_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLS + ["report"])
        w.writerow(units_out + [""])
        w.writerows(table)
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    old.update(traffic)
    json.dump(old, open(path, "w"), indent=1)
    for k, t in traffic.items():
        print(f"{k}: {t['duration_ms']:.4f} ms, DRAM {t['dram_bytes_per_launch']/1e9:.3f} GB ({t['dram_read']/1e9:.3f} r + {t['dram_write']/1e9:.3f} w), grid {t['grid']}")


if __name__ == "__main__":
    main()
