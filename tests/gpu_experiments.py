"""GPU experiments (not a test): accuracy of the split-bf16 tensor-core GEMM vs the FFMA GEMM, and kernel timings.
Writes gpurun_out/experiments.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
out = {"accuracy": [], "timing": []}


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()


def timeit(fn, n=10):
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


for (M, N, K) in [(2048, 256, 256), (2048, 256, 768), (2048, 512, 1536), (2048, 512, 3072), (2048, 256, 4096)]:
    for dist in ("normal", "positive"):
        A, W = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
        if dist == "positive":        # same-sign products: worst case for truncating accumulation
            A, W = A.abs(), W.abs()
        want = A.double() @ W.double().t()
        ffma = G.errs(G.op_gemm(A, W), want)
        a3, w3 = G.op_split_cast(A, 3), G.pack_w_parts(W, 1, 3)
        split = G.errs(G.op_gemm_tc(a3, 1, M, K, 3, w3, N), want)
        a1, w1 = G.op_split_cast(A, 1), G.op_split_cast(W, 1)
        bf16 = G.errs(G.op_gemm_tc(a1, 1, M, K, 1, w1, N), want)
        bf16_exact = G.errs(G.op_gemm_tc(a1, 1, M, K, 1, w1, N), a1.double() @ w1.double().t())
        cublas = G.errs(A @ W.t(), want)
        row = dict(M=M, N=N, K=K, dist=dist, ffma=ffma, split=split, bf16=bf16, bf16_vs_exact_products=bf16_exact, cublas_fp32=cublas)
        out["accuracy"].append(row)
        print(f"K={K:5d} {dist:8s} rel_l2: ffma {ffma['rel_l2']:.2e} split {split['rel_l2']:.2e} cublas {cublas['rel_l2']:.2e} "
              f"bf16 {bf16['rel_l2']:.2e} | max_abs: ffma {ffma['max_abs']:.2e} split {split['max_abs']:.2e} "
              f"bf16-vs-exact {bf16_exact['max_abs']:.2e}", flush=True)

# timings at the workload's shapes (M = 64*864 level-0 rows etc.)
for name, (B, T, cin, N, taps) in {"lin256": (1, 55296, 256, 256, 1), "qkv256": (1, 55296, 256, 768, 1), "geglu256": (1, 55296, 256, 2048, 1),
                                   "ff2_256": (1, 55296, 1024, 256, 1), "conv256": (64, 864, 256, 256, 3), "conv512_T216": (64, 216, 512, 512, 3),
                                   "conv1024_T108": (64, 108, 1024, 512, 3), "lin512_T216": (1, 13824, 512, 512, 1)}.items():
    M = B * T
    K = taps * cin
    A, W = rnd(M, cin, seed=3), rnd(N, K, seed=4, scale=K ** -0.5)
    fl = 2.0 * M * N * K
    t_f = timeit(lambda: G.op_gemm(A, W, M=M, taps=taps, cin=cin, t_out=T if taps == 3 else M, t_in=T if taps == 3 else M))
    res = dict(name=name, M=M, N=N, K=K, ffma_ms=t_f, ffma_tflops=fl / t_f / 1e9)
    for parts in (1, 3):
        a, w = G.op_split_cast(A, parts), G.pack_w_parts(W, taps, parts)
        t = timeit(lambda: G.op_gemm_tc(a, B, T, cin, parts, w, N, taps=taps))
        res[f"tc{parts}_ms"] = t
        res[f"tc{parts}_tflops"] = fl / t / 1e9
    A16 = A.bfloat16()
    W16 = W[:, :cin].bfloat16().contiguous()
    if taps == 1:
        t = timeit(lambda: A16 @ W16.t())
        res["cublas_bf16_ms"] = t
        res["cublas_bf16_tflops"] = fl / t / 1e9
    out["timing"].append(res)
    print(json.dumps(res), flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/experiments.json", "w"), indent=1)
