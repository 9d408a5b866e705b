"""profiles/traffic.json from an ncu csv with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch:
    python tools/measure_traffic.py gpurun_out/r02_traffic.csv profiles/traffic.json "<how it was captured>" """
import csv, json, sys
from collections import defaultdict
src, dst, how = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))
for r in csv.DictReader(lines):
    k = "conv_tc" if "k_conv_tma" in r["Kernel Name"] else "wgrad_tc" if "k_wgrad_tma" in r["Kernel Name"] else None
    if not k:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[k]["bytes"] += v
    elif r["Metric Name"] == "gpu__time_duration.sum":
        per[k]["ms"] += v / ({"ns": 1e6, "nsecond": 1e6, "us": 1e3, "usecond": 1e3, "ms": 1}.get(u, 1e6))
        per[k]["launches"] += 1
out = {k: v["bytes"] / max(v["launches"], 1) for k, v in per.items()}
out["launches"] = {k: int(v["launches"]) for k, v in per.items()}
out["dram_gbs_under_ncu"] = {k: v["bytes"] / (v["ms"] * 1e-3) / 1e9 for k, v in per.items() if v["ms"]}
out["source"] = how
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
