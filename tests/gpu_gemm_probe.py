"""GPU experiment (not a test): a few gemm_tc shapes in split-f16 mode, with / without the fp32 residual — epilogue cost probe.
usage: gpu_gemm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402

B = 64


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, (batches, rows, cin, N, taps) in {"T3 conv3 512->512": (B, 108, 512, 512, 3), "T0 conv3 256->256": (B, 864, 256, 256, 3),
                                             "T0 lin 256x256": (1, B * 864, 256, 256, 1), "T2 lin 512x512": (1, B * 216, 512, 512, 1),
                                             "T0 ff2 256x1024": (1, B * 864, 1024, 256, 1)}.items():
    M, K = batches * rows, taps * cin
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, cin, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    R = torch.randn(M, N, generator=g).cuda()
    a, w = G.op_split_cast(A, 2), G.pack_w_parts(W, taps, 2)
    t_r = timeit(lambda: G.op_gemm_tc(a, batches, rows, cin, 2, w, N, taps=taps, bias=bias, R=R))
    t_n = timeit(lambda: G.op_gemm_tc(a, batches, rows, cin, 2, w, N, taps=taps, bias=bias))
    fl = 2.0 * M * N * K
    print(f"{name:20s} nkb={K // 64:3d}  with residual {t_r:7.2f} us  without {t_n:7.2f} us   tensor bound (3 products) {fl * 3 / 1359.1e12 * 1e6:6.2f} us", flush=True)
