"""Summarise an `ncu --csv` launch list (gpu__time_duration + dram bytes per launch) per kernel name.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_target.py ...
    python tools/summarize_ncu.py gpurun_out/launches.csv profiles/r01_kernel_summary.json
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = defaultdict(lambda: {"launches": set(), "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    for r in rd:
        name = re.sub(r"\(.*", "", r.get("Kernel Name", ""))
        name = re.sub(r"^void\s+", "", name).strip()
        m, v, unit = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", ""), r.get("Metric Unit", "")
        try:
            v = float(v)
        except ValueError:
            continue
        d = per[name]
        d["launches"].add(r.get("ID"))
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if m == "gpu__time_duration.sum":
            d["time_us"] += v * scale
        elif m == "dram__bytes_read.sum":
            d["dram_read"] += v * scale
        elif m == "dram__bytes_write.sum":
            d["dram_write"] += v * scale
    total = sum(d["time_us"] for d in per.values())
    summ = {}
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["time_us"]):
        n = len(d["launches"])
        summ[k] = {"launches": n, "time_us": round(d["time_us"], 1), "share": round(d["time_us"] / total, 4) if total else 0,
                   "dram_bytes_per_launch": round((d["dram_read"] + d["dram_write"]) / n) if n else 0,
                   "dram_bytes_total": round(d["dram_read"] + d["dram_write"])}
    json.dump({"source": path, "total_time_us": round(total, 1), "kernels": summ}, open(out, "w"), indent=1)
    for k, v in list(summ.items())[:12]:
        print(f"{v['share']*100:5.1f}%  {v['time_us']:10.1f} us  x{v['launches']:4d}  {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
