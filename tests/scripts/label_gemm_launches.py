"""Labels every gemm_tc_kernel launch of a one-evaluation ncu launch list (tests/gpu_nfe_once.py, B=64, T=864) with the layer it
belongs to, by replaying the launch order of run_unet (csrc/lds_api.cu), and prints a per-shape table: measured time (cold-cache,
under ncu), algorithmic FLOPs, operand + result bytes, and the tensor / HBM bounds.   usage: label_gemm_launches.py CSV fp32|bf16"""
import csv
import sys
from collections import OrderedDict

path, prec = sys.argv[1], sys.argv[2]
B, T = 64, 864
parts = 2 if prec == "fp32" else 1
mma_per_product = 3 if prec == "fp32" else 1
PEAK_TF, HBM_GBS = 1359.1, 6552.0
ch = [256, 384, 512, 512]
Tl = [864, 432, 216, 108]

seq = []          # (label, M, N, K, out_bytes_per_elem, n_out, has_residual)


def g(label, M, N, K, out="f32", res=False, n_out=None):
    seq.append((label, M, N, K, out, n_out if n_out is not None else N, res))


def resnet(lvl, cin, cout):
    M = B * Tl[lvl]
    g(f"L{lvl} conv1 k3 {cin}->{cout}", M, cout, 3 * cin)
    if cin != cout:
        g(f"L{lvl} shortcut {cin}->{cout}", M, cout, cin)
    g(f"L{lvl} conv2 k3 {cout}->{cout} +res", M, cout, 3 * cout, res=True)


def xf(lvl, C):
    M = B * Tl[lvl]
    d = C // 8
    dpad = 32 if d <= 32 else 64
    g(f"L{lvl} proj_in {C}", M, C, C)
    for _ in range(2):
        g(f"L{lvl} qkv {C}->{3 * 8 * dpad}", M, 3 * 8 * dpad, C, out="planes")
        g(f"L{lvl} attn_out {C} +res", M, C, C, res=True)
    g(f"L{lvl} ff1 geglu {C}->{8 * C}", M, 8 * C, C, out="planes", n_out=4 * C)
    g(f"L{lvl} ff2 {4 * C}->{C} +res", M, C, 4 * C, out="planes", res=True)
    g(f"L{lvl} proj_out {C} +res", M, C, C, res=True)


g("conv_in cond half k3 256->256 (once per call)", B * T, 256, 3 * 256)
g("conv_in k3 128->256 +cond", B * T, 256, 3 * 128, res=True)
for i in range(4):
    cin = ch[i - 1] if i else ch[0]
    for j in range(2):
        resnet(i, cin if j == 0 else ch[i], ch[i])
        if i < 3:
            xf(i, ch[i])
    if i < 3:
        g(f"L{i} down k3s2 {ch[i]}", B * Tl[i + 1], ch[i], 3 * ch[i])
resnet(3, 512, 512); xf(3, 512); resnet(3, 512, 512)
skips = [256, 256, 256, 256, 384, 384, 384, 512, 512, 512, 512, 512]
cur = 512
for i in range(4):
    lvl = 3 - i
    for j in range(3):
        s = skips.pop()
        resnet(lvl, cur + s, ch[lvl]); cur = ch[lvl]
        if i > 0:
            xf(lvl, ch[lvl])
    if i < 3:
        g(f"L{lvl} up k3 {ch[lvl]}", B * Tl[lvl - 1], ch[lvl], 3 * ch[lvl])
g("conv_out k3 256->128", B * T, 128, 3 * 256)

rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
per = OrderedDict()
for r in rows:
    if r is hdr or not r[ix["ID"]].isdigit():
        continue
    d = per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3}.get(r[ix["Metric Unit"]], 1)
    d[r[ix["Metric Name"]]] = v * sc
gl = [d for d in per.values() if "gemm_tc_kernel" in d["name"]]
assert len(gl) == len(seq), (len(gl), len(seq))
agg = OrderedDict()
for (label, M, N, K, out, n_out, res), d in zip(seq, gl):
    a = agg.setdefault(label, dict(n=0, us=0.0, M=M, N=N, K=K, out=out, n_out=n_out, res=res, dram=0.0))
    a["n"] += 1; a["us"] += d["gpu__time_duration.sum"]
    a["dram"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
print(f"| layer ({prec} mode, B={B} x T={T}) | launches | avg us | alg TFLOP/s | tensor bound us | HBM bound us | time / max(bounds) | DRAM MB (ncu) / algorithmic MB |")
print("|---|---|---|---|---|---|---|---|")
tot = 0.0
for label, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    us = a["us"] / a["n"]; tot += a["us"]
    fl = 2.0 * a["M"] * a["N"] * a["K"]
    esz = 2 * parts
    by = a["M"] * (a["K"] if "k3" not in label or "s2" in label else a["K"] / 3) * esz + a["N"] * a["K"] * esz \
        + a["M"] * a["n_out"] * (4 if a["out"] == "f32" else esz) + (a["M"] * a["n_out"] * 4 if a["res"] else 0)
    tb = fl * mma_per_product / (PEAK_TF * 1e12) * 1e6
    hb = by / (HBM_GBS * 1e9) * 1e6
    print(f"| {label} | {a['n']} | {us:.1f} | {fl / us / 1e6:.0f} | {tb:.1f} | {hb:.1f} | {us / max(tb, hb):.2f} | {a['dram'] / a['n'] / 1e6:.0f} / {by / 1e6:.0f} |")
print(f"\ntotal gemm_tc time of the evaluation: {tot:.0f} us over {len(gl)} launches")
