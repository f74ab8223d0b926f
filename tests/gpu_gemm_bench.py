"""GPU experiment (not a test): gemm_tc throughput at the denoiser's GEMM shapes (SURVEY.md Appendix D, B=64, T=864),
both precision modes, against cuBLAS bf16 on the same shape, next to each shape's own BOUNDS:
  tensor_us : logical FLOPs x (3 plane products in split-f16 mode | 1) / measured sustained bf16 peak (MEASURED_PEAKS.json)
  hbm_us    : the launch's own bytes (A planes + W planes + output + fp32 residual) / measured copy bandwidth
  cublas_plus_epi_us : cuBLAS bf16 time of the bare matmul + (our output/residual bytes beyond a bf16 C) / copy bandwidth —
                       what a library GEMM followed by a perfect elementwise pass would take in bf16 mode
Epilogues are the ones the sampler uses (bias everywhere, fp32 residual on the out-projection / proj_out / FF2 / conv2 shapes).
Writes gpurun_out/gemm_bench.json and gpurun_out/gemm_bounds.md.

Logical TFLOP/s are reported; the split (fp32-accurate) mode issues 6 bf16 MMAs per logical product, so its
tensor-pipe occupancy is 6x its logical rate."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
B = 64
SHAPES = {  # name: (batches, rows, cin, N, taps, epilogue, calls per evaluation)
    "T0 lin 256x256": (1, B * 864, 256, 256, 1, 0, 50),
    "T0 qkv 768x256": (1, B * 864, 256, 768, 1, 0, 10),
    "T0 geglu 2048x256": (1, B * 864, 256, 2048, 1, 2, 5),
    "T0 ff2 256x1024": (1, B * 864, 1024, 256, 1, 0, 5),
    "T0 conv3 256->256": (B, 864, 256, 256, 3, 0, 8),
    "T0 conv3 512->256": (B, 864, 512, 256, 3, 0, 2),
    "T0 conv3 384->384": (B, 864, 384, 384, 3, 0, 1),
    "T1 lin 384x384": (1, B * 432, 384, 384, 1, 0, 50),
    "T1 geglu 3072x384": (1, B * 432, 384, 3072, 1, 2, 5),
    "T1 ff2 384x1536": (1, B * 432, 1536, 384, 1, 0, 5),
    "T1 conv3 384->384": (B, 432, 384, 384, 3, 0, 6),
    "T1 conv3 768->384": (B, 432, 768, 384, 3, 0, 1),
    "T2 lin 512x512": (1, B * 216, 512, 512, 1, 0, 50),
    "T2 geglu 4096x512": (1, B * 216, 512, 4096, 1, 2, 5),
    "T2 ff2 512x2048": (1, B * 216, 2048, 512, 1, 0, 5),
    "T2 conv3 512->512": (B, 216, 512, 512, 3, 0, 7),
    "T2 conv3 1024->512": (B, 216, 1024, 512, 3, 0, 2),
    "T3 conv3 512->512": (B, 108, 512, 512, 3, 0, 11),
    "T3 conv3 1024->512": (B, 108, 1024, 512, 3, 0, 3),
}


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


PEAKS = {"hbm_gbs": 6552.3, "bf16_tflops_sustained": 1359.1}
_pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(_pk):
    PEAKS.update({k: v for k, v in json.load(open(_pk)).items() if k in PEAKS})
rows_out = []
tot = {1: 0.0, 2: 0.0, "flops": 0.0}
for name, (batches, rows, cin, N, taps, epi, calls) in SHAPES.items():
    M, K = batches * rows, taps * cin
    A, W = rnd(M, cin, seed=3), rnd(N, K, seed=4, scale=K ** -0.5)
    n_out = N // 2 if epi == 2 else N
    bias = rnd(N, seed=5)
    with_res = epi == 0 and "qkv" not in name            # out-proj / proj_out / FF2 / conv2 carry the fp32 residual stream
    R = rnd(M, n_out, seed=6) if with_res else None
    fl = 2.0 * M * N * K
    res = dict(name=name, M=M, N=N, K=K, calls=calls, residual=with_res)
    for parts in (1, 2):
        a, w = G.op_split_cast(A, parts), G.pack_w_parts(W, taps, parts)
        out_kind = 0 if epi == 0 else (1 if parts == 1 else 2)
        t = timeit(lambda: G.op_gemm_tc(a, batches, rows, cin, parts, w, N, taps=taps, bias=bias, R=R, epilogue=epi, out_kind=out_kind))
        out_bytes = M * n_out * (4 if out_kind == 0 else 2 * parts)
        byts = M * cin * 2 * parts + N * K * 2 * parts + out_bytes + (M * n_out * 4 if with_res else 0)
        tensor_us = fl * (3 if parts == 2 else 1) / (PEAKS["bf16_tflops_sustained"] * 1e12) * 1e6
        hbm_us = byts / (PEAKS["hbm_gbs"] * 1e9) * 1e6
        res[f"tc{parts}_us"] = round(t * 1e3, 2)
        res[f"tc{parts}_tflops"] = round(fl / t / 1e9, 1)
        res[f"tc{parts}_tensor_bound_us"] = round(tensor_us, 2)
        res[f"tc{parts}_hbm_bound_us"] = round(hbm_us, 2)
        res[f"tc{parts}_bytes"] = byts
        res[f"tc{parts}_frac_of_bound"] = round(max(tensor_us, hbm_us) / (t * 1e3), 3)
        tot[parts] += t * calls
    tot["flops"] += fl * calls
    if taps == 1:
        A16, W16 = A.bfloat16(), W.bfloat16().contiguous()
        t = timeit(lambda: A16 @ W16.t())
        res["cublas_bf16_us"] = round(t * 1e3, 2)
        res["cublas_bf16_tflops"] = round(fl / t / 1e9, 1)
        extra = M * n_out * (4 - 2 if epi == 0 else 0) + (M * n_out * 4 if with_res else 0) + (M * N * 2 if epi == 2 else 0)
        res["cublas_plus_epi_us"] = round(t * 1e3 + extra / (PEAKS["hbm_gbs"] * 1e9) * 1e6, 2)
        res["tc1_vs_cublas_plus_epi"] = round(res["cublas_plus_epi_us"] / res["tc1_us"], 3)
    rows_out.append(res)
    print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gemm_bounds%s.md" % (sys.argv[1] if len(sys.argv) > 1 else ""), "w") as f:
    f.write("| shape | M | N | K | calls/NFE | split: measured us | tensor bound | HBM bound | frac of bound | bf16: measured us | tensor bound | HBM bound | "
            "frac of bound | cuBLAS bf16 us | cuBLAS + epilogue bytes us | ours / that |\n|" + "---|" * 16 + "\n")
    for r in rows_out:
        f.write("| {name} | {M} | {N} | {K} | {calls} | {tc2_us} | {tc2_tensor_bound_us} | {tc2_hbm_bound_us} | {tc2_frac_of_bound} | {tc1_us} | "
                "{tc1_tensor_bound_us} | {tc1_hbm_bound_us} | {tc1_frac_of_bound} | {cb} | {ce} | {rt} |\n".format(
                    cb=r.get("cublas_bf16_us", "-"), ce=r.get("cublas_plus_epi_us", "-"), rt=r.get("tc1_vs_cublas_plus_epi", "-"), **r))
summary = dict(weighted_ms_bf16=tot[1], weighted_ms_split=tot[2], weighted_tflops_bf16=tot["flops"] / tot[1] / 1e9,
               weighted_tflops_split=tot["flops"] / tot[2] / 1e9)
print(json.dumps(summary))
os.makedirs("gpurun_out", exist_ok=True)
TAG = sys.argv[1] if len(sys.argv) > 1 else ""
json.dump(dict(shapes=rows_out, summary=summary), open(f"gpurun_out/gemm_bench{TAG}.json", "w"), indent=1)
