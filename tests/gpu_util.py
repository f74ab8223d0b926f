"""Helpers for the GPU parity tests: ctypes calls into the stateless operator entry points and a
JSON-lines report (gpurun_out/parity_report.jsonl) of every measured error."""
import ctypes as C
import json
import os
import time

import torch

from conftest import ROOT

REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    kw["ts"] = time.time()
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def lib():
    from latent_diffusion_speech_b200 import capi
    return capi.load_library()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, lib().lds_last_error().decode()))


def pack_conv3(w):
    """[Cout, Cin, 3] -> [Cout, 3*Cin] tap-major."""
    return w.permute(0, 2, 1).contiguous().reshape(w.shape[0], -1)


def op_gemm(A, w, bias=None, R=None, r_div=1, M=None, N=None, K=None, taps=1, cin=None, t_out=None, t_in=None,
            t_conv=None, stride=1, upsample=0, up_scale=1.0, epilogue=0, out_cols=None):
    N = N or w.shape[0]
    K = K or w.shape[1]
    cin = cin or K // taps
    M = M if M is not None else A.shape[0]
    t_out = t_out or M
    t_in = t_in or t_out
    t_conv = t_conv or t_in
    oc = out_cols or (N // 2 if epilogue == 2 else N)
    out = torch.empty(M, oc, device=A.device, dtype=torch.float32)
    rc = lib().lds_op_gemm(ptr(A), A.shape[-1], ptr(w), ptr(bias), ptr(R), 0 if R is None else R.shape[-1], r_div,
                           ptr(out), oc, M, N, K, taps, cin, t_out, t_in, t_conv, stride, upsample, float(up_scale),
                           epilogue, stream())
    check(rc, "lds_op_gemm")
    return out


def op_attention(qkv, B, T, Cc, heads):
    out = torch.empty(B * T, Cc, device=qkv.device, dtype=torch.float32)
    check(lib().lds_op_attention(ptr(qkv), ptr(out), B, T, Cc, heads, stream()), "lds_op_attention")
    return out


def op_groupnorm(x1, x2, B, T, groups, eps, gamma, beta, ss=None, silu=0):
    c1 = x1.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    part = torch.empty(B * ((T + 31) // 32) * groups * 3, device=x1.device, dtype=torch.float32)
    y = torch.empty(B * T, c1 + c2, device=x1.device, dtype=torch.float32)
    check(lib().lds_op_groupnorm(ptr(x1), c1, ptr(x2), c2, B, T, groups, float(eps), ptr(gamma), ptr(beta), ptr(ss), silu,
                                 ptr(part), ptr(y), stream()), "lds_op_groupnorm")
    return y


def op_groupnorm_fused(x1, x2, B, T, groups, eps, gamma, beta, ss=None, silu=0):
    c1 = x1.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    y = torch.empty(B * T, c1 + c2, device=x1.device, dtype=torch.float32)
    check(lib().lds_op_groupnorm_fused(ptr(x1), c1, ptr(x2), c2, B, T, groups, float(eps), ptr(gamma), ptr(beta), ptr(ss), silu,
                                       ptr(y), stream()), "lds_op_groupnorm_fused")
    return y


def op_groupnorm_cluster(x1, x2, B, T, groups, eps, gamma, beta, ss=None, silu=0):
    c1 = x1.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    y = torch.empty(B * T, c1 + c2, device=x1.device, dtype=torch.float32)
    check(lib().lds_op_groupnorm_cluster(ptr(x1), c1, ptr(x2), c2, B, T, groups, float(eps), ptr(gamma), ptr(beta), ptr(ss), silu,
                                         ptr(y), stream()), "lds_op_groupnorm_cluster")
    return y


def op_layernorm(x, gamma, beta, eps=1e-5):
    y = torch.empty_like(x)
    check(lib().lds_op_layernorm(ptr(x), ptr(gamma), ptr(beta), float(eps), x.shape[0], x.shape[1], ptr(y), stream()),
          "lds_op_layernorm")
    return y


def errs(got, want):
    got, want = got.double(), want.double()
    d = (got - want).abs()
    return dict(max_abs=float(d.max()), rel_l2=float((got - want).norm() / want.norm().clamp_min(1e-30)),
                scale=float(want.abs().max()))


def op_split_cast(x, parts):
    rows, Cc = x.shape
    out = torch.empty(rows, parts * Cc, device=x.device, dtype=torch.bfloat16)
    check(lib().lds_op_split_cast(ptr(x), ptr(out), rows, Cc, parts, stream()), "lds_op_split_cast")
    return out


PLANE_SCALE = 16.0     # csrc/planes.cuh: split-f16 planes (parts = 2) hold fp16(x * 16) and fp16(x * 16 - h1)


def planes_to_double(planes, parts, Cc):
    """Operand planes [rows, parts*C] (stored in a bf16-typed tensor) -> the value they represent, fp64 [rows, C]."""
    rows = planes.shape[0]
    if parts == 2:
        return planes.view(torch.float16).view(rows, 2, Cc).double().sum(1) / PLANE_SCALE
    return planes.float().view(rows, parts, Cc).double().sum(1)


def pack_w_parts(w, taps, parts):
    """fp32 [N, taps*cin] (tap-major) -> 16-bit planes [N, taps*parts*cin] (tap, plane, channel)."""
    N, K = w.shape
    cin = K // taps
    sp = op_split_cast(w.reshape(N * taps, cin).contiguous(), parts)       # [N*taps, parts*cin]
    return sp.reshape(N, taps * parts * cin).contiguous()


def op_gemm_tc(a_bf16, batches, rows, cin, parts, w_bf16, N, taps=1, bias=None, R=None, r_div=1, out_kind=0, epilogue=0):
    n_out = N // 2 if epilogue == 2 else N
    if out_kind == 0:
        out = torch.empty(batches * rows, n_out, device=a_bf16.device, dtype=torch.float32)
        c_ld = n_out
    elif out_kind == 1:
        out = torch.empty(batches * rows, n_out, device=a_bf16.device, dtype=torch.bfloat16)
        c_ld = n_out
    else:                      # split-f16 planes [h1 | h2]
        out = torch.empty(batches * rows, 2 * n_out, device=a_bf16.device, dtype=torch.bfloat16)
        c_ld = 2 * n_out
    check(lib().lds_op_gemm_tc(ptr(a_bf16), batches, rows, cin, parts, ptr(w_bf16), N, taps, ptr(bias), ptr(R),
                               0 if R is None else R.shape[-1], r_div, ptr(out), c_ld, out_kind, epilogue, stream()),
          "lds_op_gemm_tc")
    return out


def op_conv1d_tc(a_planes, batches, rows, cin, parts, w_planes, N, taps, dil, bias=None, R=None, out_kind=0, epilogue=0, act_slope=0.0):
    if out_kind == 0:
        out = torch.empty(batches * rows, N, device=a_planes.device, dtype=torch.float32)
        c_ld = N
    else:                      # 16-bit planes (bf16: 1 plane, split-f16: 2)
        out = torch.empty(batches * rows, parts * N, device=a_planes.device, dtype=torch.bfloat16)
        c_ld = parts * N
    check(lib().lds_op_conv1d_tc(ptr(a_planes), batches, rows, cin, parts, ptr(w_planes), N, taps, dil, ptr(bias), ptr(R),
                                 0 if R is None else R.shape[-1], ptr(out), c_ld, out_kind, epilogue, act_slope, stream()),
          "lds_op_conv1d_tc")
    return out


def op_qkv_attention_tc(x, wq, wk, wv, B, T, heads, parts):
    """x fp32 [B*T, C]; returns attention output as fp32 (planes summed) [B*T, C]."""
    Cc = x.shape[1]
    d = Cc // heads
    dpad = 32 if d <= 32 else 64
    rows = []
    for w in (wq, wk, wv):
        for hh in range(heads):
            rows.append(w[hh * d:(hh + 1) * d])
            if dpad > d:
                rows.append(torch.zeros(dpad - d, Cc, device=x.device))
    wp = pack_w_parts(torch.cat(rows, 0).contiguous(), 1, parts)
    xp = op_split_cast(x, parts)
    t_pad = (T + 7) // 8 * 8
    ap = parts          # attention operands: three bf16 planes in the fp32-accurate mode
    q = torch.zeros(B * T * ap * heads * dpad, device=x.device, dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    vt = torch.zeros(B * ap * heads * dpad * t_pad, device=x.device, dtype=torch.bfloat16)
    out = torch.zeros(B * T, parts * Cc, device=x.device, dtype=torch.bfloat16)
    check(lib().lds_op_qkv_attention_tc(ptr(xp), ptr(wp), B, T, Cc, heads, dpad, parts, ptr(q), ptr(k), ptr(vt), ptr(out), stream()),
          "lds_op_qkv_attention_tc")
    return planes_to_double(out, parts, Cc)


# ---- solver / layout kernels (csrc/solver.cu) ----
def _f(v):
    return float(v)


def op_x0_pred(x, eps, sigma, alpha):
    m = torch.empty_like(x)
    check(lib().lds_op_x0_pred(ptr(x), ptr(eps), _f(sigma), _f(alpha), ptr(m), x.numel(), stream()), "lds_op_x0_pred")
    return m


def op_dpm_update(x, m0, m1, cx, cm, hcm, ir0, order):
    x = x.clone()
    check(lib().lds_op_dpm_update(ptr(x), ptr(m0), ptr(m1), _f(cx), _f(cm), _f(hcm), _f(ir0), order, x.numel(), stream()),
          "lds_op_dpm_update")
    return x


def op_unipc_predict(x, m0, m1, cx, cmE, aB, rk, rho_p, order):
    xb, xp = torch.empty_like(x), torch.empty_like(x)
    check(lib().lds_op_unipc_predict(ptr(x), ptr(m0), ptr(m1), _f(cx), _f(cmE), _f(aB), _f(rk), _f(rho_p), order, ptr(xb), ptr(xp),
                                     x.numel(), stream()), "lds_op_unipc_predict")
    return xb, xp


def op_unipc_correct(xb, m0, m1, mt, aB, rk, rho_c0, rho_c1, order):
    x = torch.empty_like(xb)
    check(lib().lds_op_unipc_correct(ptr(xb), ptr(m0), ptr(m1), ptr(mt), _f(aB), _f(rk), _f(rho_c0), _f(rho_c1), order, ptr(x),
                                     xb.numel(), stream()), "lds_op_unipc_correct")
    return x


def op_ddpm_step(x_btm, eps_btm, noise_bmt, cr, crm1, pm1, pm2, sig):
    B, T, M = x_btm.shape
    x = x_btm.clone()
    check(lib().lds_op_ddpm_step(ptr(x), ptr(eps_btm), ptr(noise_bmt), _f(cr), _f(crm1), _f(pm1), _f(pm2), _f(sig), B, T, M, stream()),
          "lds_op_ddpm_step")
    return x


def op_ddim_step(x, eps, sqrt_at, coef, sqrt_aprev):
    x = x.clone()
    check(lib().lds_op_ddim_step(ptr(x), ptr(eps), _f(sqrt_at), _f(coef), _f(sqrt_aprev), x.numel(), stream()), "lds_op_ddim_step")
    return x


def op_pndm_update(x, e, h1, h2, h3, d, k1, k2, mode):
    out = torch.empty_like(x)
    check(lib().lds_op_pndm_update(ptr(x), ptr(e), ptr(h1), ptr(h2), ptr(h3), _f(d), _f(k1), _f(k2), mode, ptr(out), x.numel(), stream()),
          "lds_op_pndm_update")
    return out


def op_q_sample(gt_btm, noise_bmt, ascale, sa, sb):
    B, T, M = gt_btm.shape
    x = torch.empty_like(gt_btm)
    check(lib().lds_op_q_sample(ptr(x), ptr(gt_btm), ptr(noise_bmt), _f(ascale), _f(sa), _f(sb), B, T, M, stream()), "lds_op_q_sample")
    return x


def op_cast_gather(x_btc, t_out, parts, mode, scale=0.0):
    B, t_in, Cc = x_btc.shape
    width = parts * Cc * (3 if mode == 2 else 1)
    out = torch.empty(B, t_out, width, device=x_btc.device, dtype=torch.bfloat16)
    check(lib().lds_op_cast_gather(ptr(x_btc), ptr(out), B, t_in, t_out, Cc, parts, mode, float(scale), stream()), "lds_op_cast_gather")
    return out


def op_transpose(x, to_channels_last, scale=1.0):
    if to_channels_last:
        B, Cc, T = x.shape
        out = torch.empty(B, T, Cc, device=x.device, dtype=torch.float32)
    else:
        B, T, Cc = x.shape
        out = torch.empty(B, Cc, T, device=x.device, dtype=torch.float32)
    check(lib().lds_op_transpose(ptr(x), ptr(out), B, Cc, T, float(scale), int(to_channels_last), stream()), "lds_op_transpose")
    return out


def op_div_copy(x, d):
    out = torch.empty_like(x)
    check(lib().lds_op_div_copy(ptr(x), ptr(out), x.numel(), float(d), stream()), "lds_op_div_copy")
    return out


def split_planes_ref(x, parts):
    """fp32 [..., C] -> 16-bit planes [..., parts*C] (as a bf16-typed tensor) by the definition of the operand planes:
    parts 1 / 3: hi = bf16(x), mid = bf16(x - hi), ...; parts 2 (split-f16): h1 = f16(16 x), h2 = f16(16 x - h1)."""
    if parts == 2:
        xs = x.float() * PLANE_SCALE
        h1 = xs.clamp(-65504.0, 65504.0).to(torch.float16)
        h2 = (xs - h1.float()).to(torch.float16)
        return torch.cat([h1, h2], dim=-1).view(torch.bfloat16)
    planes, r = [], x.float()
    for _ in range(parts):
        p = r.to(torch.bfloat16)
        planes.append(p)
        r = r - p.float()
    return torch.cat(planes, dim=-1)
