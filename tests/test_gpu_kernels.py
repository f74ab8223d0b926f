"""GPU: every hand-written kernel against the PyTorch op it replaces (fp64 ground truth), through the
stateless C-ABI operator entry points.  Tolerances are fp32 round-off for the reduction length."""
import math

import pytest
import torch
import torch.nn.functional as F

import gpu_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (1000, 384, 1536), (77, 128, 768), (4096, 768, 256), (1, 1024, 256)])
def test_linear_bias_residual(M, N, K):
    A, W, b, R = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3), _rand(M, N, seed=4)
    got = G.op_gemm(A, W, b, R)
    want = A.double() @ W.double().t() + b.double() + R.double()
    e = G.errs(got, want)
    G.report(test="linear", M=M, N=N, K=K, **e)
    assert e["max_abs"] <= 2e-5 * math.sqrt(K / 256) * max(1.0, e["scale"] / 4), e


def test_linear_silu_and_broadcast_residual():
    B, T, K, N = 3, 50, 1280, 256
    A, W, b, R = _rand(B * T, K, seed=5), _rand(N, K, seed=6, scale=K ** -0.5), _rand(N, seed=7), _rand(B, N, seed=8)
    got = G.op_gemm(A, W, b, R, r_div=T)
    want = (A.double() @ W.double().t() + b.double()).view(B, T, N) + R.double()[:, None, :]
    e = G.errs(got, want.reshape(B * T, N))
    G.report(test="linear_bcast_residual", **e)
    assert e["max_abs"] <= 5e-5, e
    got = G.op_gemm(A, W, b, epilogue=1)
    e = G.errs(got, F.silu(A.double() @ W.double().t() + b.double()))
    G.report(test="linear_silu", **e)
    assert e["max_abs"] <= 5e-5, e


@pytest.mark.parametrize("C", [256, 384])
def test_geglu_epilogue(C):
    M = 333
    A, W, b, R = _rand(M, C, seed=9), _rand(8 * C, C, seed=10, scale=C ** -0.5), _rand(8 * C, seed=11), None
    # interleave rows per 128-tile: [64 value | 64 gate]
    idx = []
    for t in range(4 * C // 64):
        idx += list(range(t * 64, t * 64 + 64)) + list(range(4 * C + t * 64, 4 * C + t * 64 + 64))
    idx = torch.tensor(idx, device=DEV)
    got = G.op_gemm(A, W[idx].contiguous(), b[idx].contiguous(), R, epilogue=2)
    y = A.double() @ W.double().t() + b.double()
    want = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
    e = G.errs(got, want)
    G.report(test="geglu", C=C, **e)
    assert got.shape == (M, 4 * C) and e["max_abs"] <= 5e-5, e


@pytest.mark.parametrize("B,T,Cin,Cout,stride", [(2, 37, 256, 384, 1), (3, 100, 384, 128, 1), (2, 37, 256, 256, 2),
                                                 (2, 40, 512, 512, 2), (1, 864, 640, 256, 1)])
def test_conv3(B, T, Cin, Cout, stride):
    x = _rand(B, T, Cin, seed=12)                       # channels-last
    w, b = _rand(Cout, Cin, 3, seed=13, scale=(3 * Cin) ** -0.5), _rand(Cout, seed=14)
    t_out = (T - 1) // stride + 1
    R = _rand(B * t_out, Cout, seed=15)
    got = G.op_gemm(x.view(B * T, Cin), G.pack_conv3(w), b, R, M=B * t_out, taps=3, cin=Cin, t_out=t_out, t_in=T,
                    t_conv=T, stride=stride)
    want = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), stride=stride, padding=1).transpose(1, 2)
    e = G.errs(got, want.reshape(B * t_out, Cout) + R.double())
    G.report(test="conv3", B=B, T=T, Cin=Cin, Cout=Cout, stride=stride, **e)
    assert e["max_abs"] <= 5e-5, e


@pytest.mark.parametrize("t_in,t_up", [(13, 26), (108, 216), (54, 107), (108, 215), (323, 646)])
def test_conv3_fused_nearest_upsample(t_in, t_up):
    B, Cc = 2, 384
    x = _rand(B, t_in, Cc, seed=16)
    w, b = _rand(Cc, Cc, 3, seed=17, scale=(3 * Cc) ** -0.5), _rand(Cc, seed=18)
    exact = t_up == 2 * t_in
    scale = 0.5 if exact else float(torch.tensor(t_in, dtype=torch.float32) / torch.tensor(t_up, dtype=torch.float32))
    got = G.op_gemm(x.view(B * t_in, Cc), G.pack_conv3(w), b, M=B * t_up, taps=3, cin=Cc, t_out=t_up, t_in=t_in,
                    t_conv=t_up, stride=1, upsample=1, up_scale=scale)
    xc = x.double().transpose(1, 2)
    up = F.interpolate(xc, scale_factor=2.0, mode="nearest") if exact else F.interpolate(xc, size=(t_up,), mode="nearest")
    want = F.conv1d(up, w.double(), b.double(), padding=1).transpose(1, 2).reshape(B * t_up, Cc)
    e = G.errs(got, want)
    G.report(test="conv3_upsample", t_in=t_in, t_up=t_up, **e)
    assert e["max_abs"] <= 5e-5, e


@pytest.mark.parametrize("B,T,C", [(2, 100, 256), (1, 864, 256), (2, 431, 384), (3, 216, 512), (2, 64, 512), (1, 1, 256)])
def test_attention(B, T, C):
    heads = 8
    qkv = _rand(B * T, 3 * C, seed=19)
    got = G.op_attention(qkv, B, T, C, heads)
    q, k, v = (z.view(B, T, heads, C // heads).transpose(1, 2) for z in qkv.double().view(B, T, 3, C).unbind(2))
    want = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, C)
    e = G.errs(got, want)
    G.report(test="attention", B=B, T=T, C=C, **e)
    assert e["max_abs"] <= 2e-5, e


@pytest.mark.parametrize("B,T,c1,c2,ss,silu,eps", [(2, 37, 256, 0, False, 1, 1e-5), (2, 100, 512, 384, False, 1, 1e-5),
                                                   (3, 33, 384, 256, False, 1, 1e-5), (2, 64, 384, 0, True, 1, 1e-5),
                                                   (2, 216, 512, 0, False, 0, 1e-6), (1, 864, 512, 512, False, 1, 1e-5),
                                                   (2, 861, 256, 0, True, 1, 1e-5), (1, 2584, 384, 256, False, 1, 1e-5), (3, 5, 512, 0, False, 1, 1e-5),
                                                   (2, 2584, 256, 0, False, 1, 1e-5), (70, 108, 512, 0, True, 1, 1e-5)])
@pytest.mark.parametrize("fused", [0, 1, 2])        # 0 stats + apply, 1 one CTA per (utterance, group), 2 cluster (the sampler's default)
def test_groupnorm(B, T, c1, c2, ss, silu, eps, fused):
    C = c1 + c2
    if fused == 1 and T * (C // 8) * 4 > 200 * 1024:
        pytest.skip("slab exceeds shared memory: single-CTA form not applicable to this shape")
    if fused == 2 and 2 * ((T + 7) // 8) * (C // 8) * 4 > 200 * 1024:
        pytest.skip("two buffers of an eighth of the slab exceed shared memory: the sampler uses stats + apply for this shape")
    x1 = _rand(B * T, c1, seed=20, scale=3.0) + 50.0          # large mean: exercises the variance formulation
    x2 = _rand(B * T, c2, seed=21, scale=0.5) if c2 else None
    gamma, beta = _rand(C, seed=22), _rand(C, seed=23)
    sst = _rand(2 * C, seed=24, scale=0.3) if ss else None
    got = (G.op_groupnorm, G.op_groupnorm_fused, G.op_groupnorm_cluster)[fused](x1, x2, B, T, 8, eps, gamma, beta, sst, silu)
    x = x1 if x2 is None else torch.cat([x1, x2], dim=-1)
    xc = x.double().view(B, T, C).transpose(1, 2)
    y = F.group_norm(xc, 8, gamma.double(), beta.double(), eps)
    if ss:
        y = y * (1 + sst.double()[:C, None]) + sst.double()[C:, None]
    if silu:
        y = F.silu(y)
    e = G.errs(got, y.transpose(1, 2).reshape(B * T, C))
    G.report(test="groupnorm", B=B, T=T, c1=c1, c2=c2, fused=fused, **e)
    assert e["max_abs"] <= 3e-5 * max(1.0, e["scale"]), e


@pytest.mark.parametrize("rows,C", [(1000, 256), (333, 384), (7, 512)])
def test_layernorm(rows, C):
    x, gamma, beta = _rand(rows, C, seed=25, scale=2.0) + 3.0, _rand(C, seed=26), _rand(C, seed=27)
    got = G.op_layernorm(x, gamma, beta)
    e = G.errs(got, F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5))
    G.report(test="layernorm", rows=rows, C=C, **e)
    assert e["max_abs"] <= 2e-5, e


# ------------------------------------------------------------------------------------------------------------
# tcgen05 / TMEM / TMA GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 256), (1000, 384, 1536), (4096, 768, 256), (77, 128, 1280)])
def test_tc_gemm_bf16_exact_products(M, N, K):
    """Plain bf16 mode against the exact product of the bf16-rounded operands (isolates the kernel mechanics)."""
    A, W, b, R = _rand(M, K, seed=31), _rand(N, K, seed=32, scale=K ** -0.5), _rand(N, seed=33), _rand(M, N, seed=34)
    a16, w16 = G.op_split_cast(A, 1), G.op_split_cast(W, 1)
    got = G.op_gemm_tc(a16, 1, M, K, 1, w16, N, bias=b, R=R)
    want = a16.double() @ w16.double().t() + b.double() + R.double()
    e = G.errs(got, want)
    G.report(test="tc_gemm_bf16", M=M, N=N, K=K, **e)
    assert e["max_abs"] <= 3e-5 * math.sqrt(K / 256) * max(1.0, e["scale"] / 4), e


@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (1000, 384, 1536), (555, 512, 3072)])
def test_tc_gemm_split_fp32_accuracy(M, N, K):
    """Split-f16 (two fp16 planes per operand, 3 plane products, main + small accumulator) against fp64 of the fp32 operands:
    must be fp32-grade — no worse than the FFMA GEMM on the same data."""
    A, W, b = _rand(M, K, seed=35), _rand(N, K, seed=36, scale=K ** -0.5), _rand(N, seed=37)
    a3, w3 = G.op_split_cast(A, 2), G.pack_w_parts(W, 1, 2)
    got = G.op_gemm_tc(a3, 1, M, K, 2, w3, N, bias=b)
    want = A.double() @ W.double().t() + b.double()
    e = G.errs(got, want)
    ffma = G.errs(G.op_gemm(A, W, b), want)
    G.report(test="tc_gemm_split", M=M, N=N, K=K, split=e, ffma=ffma)
    assert e["max_abs"] <= 5e-5 * math.sqrt(K / 256), (e, ffma)


@pytest.mark.parametrize("B,T,Cin,Cout,parts", [(2, 37, 256, 384, 1), (3, 100, 384, 128, 1), (2, 300, 640, 256, 2), (1, 864, 256, 256, 2),
                                                  (3, 96, 256, 256, 2), (2, 864, 256, 256, 1), (5, 160, 384, 384, 2)])   # last three: flat 32-row block tiling
def test_tc_conv3(B, T, Cin, Cout, parts):
    x = _rand(B, T, Cin, seed=38)
    w, b = _rand(Cout, Cin, 3, seed=39, scale=(3 * Cin) ** -0.5), _rand(Cout, seed=40)
    R = _rand(B * T, Cout, seed=41)
    a = G.op_split_cast(x.view(B * T, Cin), parts)
    wp = G.pack_w_parts(G.pack_conv3(w), 3, parts)
    got = G.op_gemm_tc(a, B, T, Cin, parts, wp, Cout, taps=3, bias=b, R=R)
    if parts == 1:
        xr, wr = a.view(B, T, Cin).double(), w.bfloat16().double()
    else:
        xr, wr = x.double(), w.double()
    want = F.conv1d(xr.transpose(1, 2), wr, b.double(), padding=1).transpose(1, 2).reshape(B * T, Cout) + R.double()
    e = G.errs(got, want)
    G.report(test="tc_conv3", B=B, T=T, Cin=Cin, Cout=Cout, parts=parts, **e)
    assert e["max_abs"] <= 6e-5, e


@pytest.mark.parametrize("out_kind,parts", [(1, 1), (2, 2)])
def test_tc_gemm_geglu_and_output_kinds(out_kind, parts):
    M, C = 333, 256
    A, W, b = _rand(M, C, seed=42), _rand(8 * C, C, seed=43, scale=C ** -0.5), _rand(8 * C, seed=44)
    idx = []
    for t in range(4 * C // 64):
        idx += list(range(t * 64, t * 64 + 64)) + list(range(4 * C + t * 64, 4 * C + t * 64 + 64))
    idx = torch.tensor(idx, device=DEV)
    a = G.op_split_cast(A, parts)
    wp = G.pack_w_parts(W[idx].contiguous(), 1, parts)
    got = G.op_gemm_tc(a, 1, M, C, parts, wp, 8 * C, bias=b[idx].contiguous(), out_kind=out_kind, epilogue=2)
    if parts == 1:
        y = a.double() @ G.op_split_cast(W, 1).double().t() + b.double()
        want = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
        e = G.errs(got.float(), want)
        tol = 2e-2 * max(1.0, e["scale"])            # bf16 output rounding
    else:
        y = A.double() @ W.double().t() + b.double()
        want = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
        e = G.errs(G.planes_to_double(got, 2, 4 * C), want)
        tol = 5e-5
    G.report(test="tc_geglu", out_kind=out_kind, parts=parts, **e)
    assert e["max_abs"] <= tol, e


@pytest.mark.parametrize("B,T,C,parts", [(2, 100, 256, 2), (1, 864, 256, 2), (2, 431, 384, 2), (3, 216, 512, 2), (2, 64, 512, 2),
                                         (1, 1, 256, 2), (2, 300, 256, 1), (2, 216, 384, 1), (1, 129, 512, 1)])
def test_tc_qkv_attention(B, T, C, parts):
    """Fused QKV projection (attention-operand epilogue) + tcgen05 flash attention vs fp64 SDPA."""
    heads = 8
    x = _rand(B * T, C, seed=51)
    wq, wk, wv = (_rand(C, C, seed=52 + i, scale=C ** -0.5) for i in range(3))
    got = G.op_qkv_attention_tc(x, wq, wk, wv, B, T, heads, parts)
    xr = x.double() if parts == 2 else x.bfloat16().double()
    ws = [w.double() if parts == 2 else w.bfloat16().double() for w in (wq, wk, wv)]
    q, k, v = ((xr @ w.t()).view(B, T, heads, C // heads).transpose(1, 2) for w in ws)
    if parts == 1:      # the projection results are rounded to bf16 before the attention in bf16 mode
        q, k, v = (z.float().bfloat16().double() for z in (q, k, v))
    want = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, C)
    e = G.errs(got, want)
    G.report(test="tc_qkv_attention", B=B, T=T, C=C, parts=parts, **e)
    assert e["max_abs"] <= (3e-5 if parts == 2 else 3e-2), e
