"""DRAM bytes per kernel family from an ncu launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv (one training iteration):
python tools/dram_traffic.py launches.csv out.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.defaultdict(dict)
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[d["ID"]]["name"] = d["Kernel Name"]
    per[d["ID"]][d["Metric Name"]] = v * scale


def family(pred):
    sel = [k for k in per.values() if pred(k["name"])]
    return {"launches": len(sel), "dram_read_bytes": sum(k.get("dram__bytes_read.sum", 0.0) for k in sel),
            "dram_write_bytes": sum(k.get("dram__bytes_write.sum", 0.0) for k in sel),
            "kernel_ms_under_ncu": sum(k.get("gpu__time_duration.sum", 0.0) for k in sel)}


fprop = family(lambda n: "conv_halo_kernel" in n or "conv_fprop_kernel" in n or "conv_splitk_kernel" in n)
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one "
                 "train256 iteration (B=32, style mixing on): all conv_halo_kernel / conv_splitk_kernel / conv_fprop_kernel launches "
                 "(fprop, dgrad, tangent; plain, pooled 4x4-stride-2, transposed and fused-style forms)"}
out.update(fprop)
out["wgrad_family"] = family(lambda n: "wgrad" in n and "unpack" not in n)
out["all_kernels"] = family(lambda n: True)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
