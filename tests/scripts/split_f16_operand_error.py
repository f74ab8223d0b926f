"""CPU experiment behind the split-f16 operand form (csrc/planes.cuh): operand-truncation error of
  (a) round 1's three bf16 planes / six products and (b) two fp16 planes / three products, with and without the power-of-two scale,
against fp64 of the fp32 operands (products and sums in fp64, so only the operand representation is measured).
Output committed as profiles/r02_split_f16_operand_error.txt."""
import torch

torch.manual_seed(0)


def split_bf16x3(x):
    p, r = [], x.clone()
    for _ in range(3):
        h = r.to(torch.bfloat16).float()
        p.append(h.double())
        r = r - h
    return p


def split_f16x2(x, scale):
    xs = x * scale
    h1 = xs.to(torch.float16).float()
    h2 = (xs - h1).to(torch.float16).float()
    return [h1.double() / scale, h2.double() / scale]


for K, dist in [(256, "normal"), (1024, "normal"), (3072, "normal"), (256, "positive"), (1024, "positive")]:
    M, N = 512, 256
    A, W = torch.randn(M, K), torch.randn(N, K) * K ** -0.5
    if dist == "positive":
        A, W = A.abs(), W.abs()
    ref = A.double() @ W.double().t()
    a, w = split_bf16x3(A), split_bf16x3(W)
    six = a[2] @ w[0].t() + a[0] @ w[2].t() + a[1] @ w[1].t() + a[1] @ w[0].t() + a[0] @ w[1].t() + a[0] @ w[0].t()
    e6 = (six - ref).norm() / ref.norm()
    for sa, sw in [(1.0, 1.0), (16.0, 4096.0)]:
        a2, w2 = split_f16x2(A, sa), split_f16x2(W, sw)
        three = a2[0] @ w2[1].t() + a2[1] @ w2[0].t() + a2[0] @ w2[0].t()
        e3 = (three - ref).norm() / ref.norm()
        print(f"K={K:5d} {dist:8s}: bf16x3 / 6 products rel-L2 {e6:.2e} | f16x2 / 3 products, scales ({sa:g}, {sw:g}) rel-L2 {e3:.2e} "
              f"max-abs {float((three - ref).abs().max()):.2e}   [fp32 FFMA GEMM on the same data: 2.9e-7 (K=256) ... 9.9e-7 (K=3072)]")
