"""Summarise an ncu csv with gpu__time_duration.sum, dram__bytes_{read,write}.sum (+ tensor pipe) PER LAUNCH CLASS into a json
that bench.py reads for roofline.traffic / by_kernel[*].traffic. Classes follow bench.py's names: the GEMM kernels carry their
class in the template arguments (tapconv_kernel<T, kEpi, kTaps>, wgrad_kernel<T, kTaps>), the fused graph-conv kernels in
their names. Usage: summarize_traffic.py <ncu.csv> <out.json>"""
import collections, csv, json, sys

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if l.startswith('"') or l.startswith('ID,') or l[:1].isdigit()]
per = collections.defaultdict(dict)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u, m = row["Metric Unit"], row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per[row["ID"]]["name"] = row["Kernel Name"]
    per[row["ID"]][m] = v


# keep only the last complete train step: the launches after the second-to-last optimizer kernel up to the last one
ids = sorted(per, key=int)
opt = [i for i in ids if "rmsprop_kernel" in per[i]["name"]]
if len(opt) >= 2:
    lo, hi = int(opt[-2]), int(opt[-1])
    per = {i: per[i] for i in ids if lo < int(i) <= hi}
    print(f"last step: launches {lo + 1}..{hi} ({len(per)} kernels)")


def klass(name):
    if "gcn_fwd_kernel" in name: return "gcn_fwd"
    if "gcn_wgrad_kernel" in name: return "gcn_wgrad"
    if "gcn_bwd_kernel" in name: return "gcn_bwd"
    def arg(n, key, i):        # "...tapconv_kernel<__nv_bfloat16, 8, 1, 0>(...)": kTaps is template argument i (1 / true)
        a = n.split(key + "<", 1)[1].split(">", 1)[0].split(",")[i].strip()
        return a in ("1", "true", "(bool)1")
    if "fmm::tapconv_kernel<" in name: return "tapconv_taps" if arg(name, "tapconv_kernel", 2) else "tapconv_1x1"
    if "fmm::wgrad_kernel<" in name: return "wgrad_taps" if arg(name, "wgrad_kernel", 1) else "wgrad_1x1"
    if "fmm::wgrad_tma_kernel<" in name: return "wgrad_taps" if arg(name, "wgrad_tma_kernel", 0) else "wgrad_1x1"
    return None


fam = collections.defaultdict(lambda: collections.defaultdict(float))
for k in per.values():
    f = klass(k["name"])
    if not f:
        continue
    a = fam[f]
    a["launches"] += 1
    a["us"] += k["gpu__time_duration.sum"]
    a["rd"] += k.get("dram__bytes_read.sum", 0.0)
    a["wr"] += k.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * k["gpu__time_duration.sum"]
res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg."
                 "pct_of_peak_sustained_elapsed over one eager train step of bench.py (" + path + ")"}
for f, a in sorted(fam.items()):
    n = a["launches"]
    res[f] = {"launches": int(n), "avg_us": round(a["us"] / n, 1), "avg_dram_read_bytes": int(a["rd"] / n),
              "avg_dram_write_bytes": int(a["wr"] / n), "avg_dram_bytes_per_launch": int((a["rd"] + a["wr"]) / n),
              "tensor_pipe_active_pct_time_weighted": round(a["tensor_w"] / a["us"], 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
