"""GPU experiment (not a test): one gemm_tc shape, a few launches — the target of an ncu capture.
usage: gpu_gemm_one.py batches rows cin N taps epilogue parts out_kind [R]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402

batches, rows, cin, N, taps, epi, parts, out_kind = (int(v) for v in sys.argv[1:9])
g = torch.Generator().manual_seed(0)
M, K = batches * rows, taps * cin
A = torch.randn(M, cin, generator=g).cuda()
W = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
R = torch.randn(M, N, generator=g).cuda() if (len(sys.argv) > 9 and sys.argv[9] == "R") else None
a, w = G.op_split_cast(A, parts), G.pack_w_parts(W, taps, parts)
for _ in range(5):
    G.op_gemm_tc(a, batches, rows, cin, parts, w, N, taps=taps, epilogue=epi, out_kind=out_kind, R=R)
torch.cuda.synchronize()
print("ok")
