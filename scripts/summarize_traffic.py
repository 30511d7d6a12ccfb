"""Summarise an ncu csv with gpu__time_duration.sum, dram__bytes_{read,write}.sum (+ tensor pipe) per kernel family
into profiles/r01_gemm_dram_traffic.json (read by bench.py for roofline.traffic)."""
import collections, csv, json, sys
path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if l.startswith('"')]
per = collections.defaultdict(dict)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per[row["ID"]]["name"] = row["Kernel Name"]
    per[row["ID"]][m] = v
fam = collections.defaultdict(lambda: collections.defaultdict(float))
for k in per.values():
    f = "tapconv" if "tapconv" in k["name"] else ("wgrad" if "wgrad" in k["name"] else None)
    if not f:
        continue
    a = fam[f]
    a["launches"] += 1
    a["us"] += k["gpu__time_duration.sum"]
    a["rd"] += k.get("dram__bytes_read.sum", 0.0)
    a["wr"] += k.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * k["gpu__time_duration.sum"]
res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... "
                 "-k regex:tapconv|wgrad over one train step (profiles/r01_gemm_dram_traffic.csv)"}
for f, a in fam.items():
    n = a["launches"]
    res[f] = {"launches": int(n), "avg_us": round(a["us"] / n, 1), "avg_dram_read_bytes": int(a["rd"] / n),
              "avg_dram_write_bytes": int(a["wr"] / n), "avg_dram_bytes_per_launch": int((a["rd"] + a["wr"]) / n),
              "tensor_pipe_active_pct_time_weighted": round(a["tensor_w"] / a["us"], 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
