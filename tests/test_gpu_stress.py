"""GPU: what stands in for compute-sanitizer on this pool (the tool is closed here, profiles/r02_sanitizer_closed.txt).

  * guard bands   — every tensor-core / TMA kernel writes into the middle of a larger NaN-filled allocation at ragged shapes;
                    the bands on both sides must stay untouched (out-of-bounds st.global / TMA store) and the payload NaN-free
                    (rows that were never written);
  * determinism   — 8 runs on the same inputs must be bit-identical (races in the mbarrier / TMEM hand-offs of the hand-rolled
                    pipelines show up as run-to-run differences)."""
import ctypes as ct

import pytest
import torch

import gpu_util as G
from conftest import gpu_model_for
from oracle import unit2mel_oracle as O

pytestmark = pytest.mark.gpu
GUARD = 4096        # elements on each side


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()


def _guarded(n, dtype):
    buf = torch.full((n + 2 * GUARD,), float("nan"), device="cuda", dtype=dtype)
    return buf, buf[GUARD:GUARD + n]


def _bands_clean(buf, n):
    return bool(torch.isnan(buf[:GUARD]).all()) and bool(torch.isnan(buf[GUARD + n:]).all())


@pytest.mark.parametrize("parts", [1, 2])
@pytest.mark.parametrize("batches,rows,cin,N,taps,res", [
    (1, 300, 256, 256, 1, True),      # ragged M, TMA epilogue with residual
    (1, 55, 384, 384, 1, False),      # one partial tile + ghost tile
    (3, 37, 256, 256, 3, True),       # per-utterance tails (conv, 8 does not divide T), residual
    (5, 96, 128, 512, 3, False),      # flat 32-frame block tiling, odd tile count
    (2, 431, 512, 128, 3, True),
])
def test_gemm_tc_fp32_output_guard_bands_and_determinism(parts, batches, rows, cin, N, taps, res):
    M, K = batches * rows, taps * cin
    A, W = _rand(M, cin, seed=1), _rand(N, K, seed=2, scale=K ** -0.5)
    bias = _rand(N, seed=3)
    R = _rand(M, N, seed=4) if res else None
    a, w = G.op_split_cast(A, parts), G.pack_w_parts(W, taps, parts)
    outs = []
    for _ in range(8):
        buf, out = _guarded(M * N, torch.float32)
        G.check(G.lib().lds_op_gemm_tc(G.ptr(a), batches, rows, cin, parts, G.ptr(w), N, taps, G.ptr(bias), G.ptr(R), 0 if R is None else N, 1,
                                       ct.c_void_p(out.data_ptr()), N, 0, 0, G.stream()), "lds_op_gemm_tc")
        torch.cuda.synchronize()
        assert _bands_clean(buf, M * N), "write outside the output tensor"
        assert not torch.isnan(out).any(), "output rows left unwritten"
        outs.append(out.clone())
    assert all(torch.equal(outs[0], o) for o in outs[1:]), "run-to-run difference"
    # and it is the right answer (fp64, conv = shifted rows with zero padding per utterance)
    x = A.double().view(batches, rows, cin)
    if taps == 3:
        xp = torch.nn.functional.pad(x, (0, 0, 1, 1))
        x = torch.cat([xp[:, t:t + rows] for t in range(3)], dim=-1)
    want = x.reshape(M, K) @ W.double().t() + bias.double() + (0 if R is None else R.double())
    e = G.errs(outs[0].view(M, N), want)
    assert e["max_abs"] <= (5e-5 if parts == 2 else 6e-2) * max(1.0, e["scale"] / 4), e


@pytest.mark.parametrize("parts", [1, 2])
@pytest.mark.parametrize("B,T,C", [(2, 40, 256), (3, 37, 384), (1, 300, 512)])
def test_attention_tc_guard_bands_and_determinism(parts, B, T, C):
    heads, d = 8, C // 8
    dpad = 32 if d <= 32 else 64
    x = _rand(B * T, C, seed=5)
    ws = [_rand(C, C, seed=6 + i, scale=C ** -0.5) for i in range(3)]
    rows = []
    for wmat in ws:
        for hh in range(heads):
            rows.append(wmat[hh * d:(hh + 1) * d])
            if dpad > d:
                rows.append(torch.zeros(dpad - d, C, device="cuda"))
    wp = G.pack_w_parts(torch.cat(rows, 0).contiguous(), 1, parts)
    xp = G.op_split_cast(x, parts)
    t_pad = (T + 7) // 8 * 8
    ap = parts
    n_qk, n_vt, n_out = B * T * ap * heads * dpad, B * ap * heads * dpad * t_pad, B * T * parts * C
    outs = []
    for _ in range(8):
        bq, q = _guarded(n_qk, torch.bfloat16)
        bk, k = _guarded(n_qk, torch.bfloat16)
        bv, vt = _guarded(n_vt, torch.bfloat16)
        bo, out = _guarded(n_out, torch.bfloat16)
        vt.zero_()                      # the key padding [T, T_pad) of V^T is never written by the projection and never read beyond T
        G.check(G.lib().lds_op_qkv_attention_tc(G.ptr(xp), G.ptr(wp), B, T, C, heads, dpad, parts, ct.c_void_p(q.data_ptr()),
                                                ct.c_void_p(k.data_ptr()), ct.c_void_p(vt.data_ptr()), ct.c_void_p(out.data_ptr()), G.stream()), "op")
        torch.cuda.synchronize()
        for buf, n in ((bq, n_qk), (bk, n_qk), (bv, n_vt), (bo, n_out)):
            assert _bands_clean(buf, n), "write outside an attention operand / output tensor"
        assert not torch.isnan(out.float()).any()
        outs.append(out.clone())
    assert all(torch.equal(outs[0], o) for o in outs[1:]), "run-to-run difference"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampler_run_to_run_bit_identical(precision, host_model):
    model = gpu_model_for(host_model, precision)
    B, T = 3, 216
    units, spk, noise, _, _ = O.synthetic_inputs(B, T)
    outs = []
    with torch.no_grad():
        for _ in range(4):
            outs.append(model(units.cuda(), None, spk_id=spk.cuda(), infer=True, infer_speedup=200, method="unipc", noise=noise.cuda()).clone())
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
