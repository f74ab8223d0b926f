"""Builds profiles/rNN_gemm_traffic.json (the `roofline.traffic` figure of bench.py) from the ncu launch lists of ONE denoiser
evaluation per mode (tests/scripts/gpu_launch_lists.sh -> gpurun_out/nfe_launches_{fp32,bf16}.csv):
average dram__bytes_read.sum + dram__bytes_write.sum per gemm_tc_kernel launch.   usage: make_gemm_traffic.py r02"""
import csv
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one denoiser evaluation "
                 f"(tests/gpu_nfe_once.py, B=64, T=864): profiles/{tag}_ncu_launches_one_nfe_{{fp32,bf16}}.csv; average over the "
                 f"gemm_tc_kernel launches of the LAST evaluation in the list"}
for prec in ("fp32", "bf16"):
    rows = [r for r in csv.reader(open(f"gpurun_out/nfe_launches_{prec}.csv")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if r is not hdr and r[ix["ID"]].isdigit()]
    # per launch ID: collect metrics
    per = {}
    for r in data:
        d = per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
        d[r[ix["Metric Name"]]] = val * scale
    ids = sorted(per)
    n_eval = sum(1 for i in ids if "transpose" in per[i]["name"]) // 2 or 1          # two boundary transposes per lds_denoise
    last = ids[len(ids) - len(ids) // n_eval:] if n_eval > 1 else ids
    g = [per[i] for i in last if "gemm_tc_kernel" in per[i]["name"]]
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in g) / max(1, len(g))
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in g) / max(1, len(g))
    us = sum(d.get("gpu__time_duration.sum", 0) for d in g) / max(1, len(g))
    out[prec] = {"launches": len(g), "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "avg_us_under_ncu": us,
                 "traffic_bytes_per_launch": rd + wr}
json.dump(out, open(f"profiles/{tag}_gemm_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
