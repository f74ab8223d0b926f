"""GPU: the per-step sampler kernels and the layout kernels of csrc/solver.cu, each through its own C-ABI entry point
(lds_op_*), must be BIT-IDENTICAL (torch.equal) to the eager PyTorch expressions of the reference on the same device:

    x0 prediction                dpm_solver_pytorch.py:433-442, uni_pc.py:285-294
    DPM-Solver++ updates         dpm_solver_pytorch.py:569-576 (first order), 813-831 (multistep second order)
    UniPC-bh2 predictor/corrector uni_pc.py:545-568
    DDPM ancestral step          diffusion.py:95-121
    DDIM / PLMS steps            diffusion.py:123-167
    q_sample (shallow start)     diffusion.py:169-171,208-212 (+ norm_spec :86)
    nearest upsample / stride-2  resnet.py:157-160,200 (as gathering casts into bf16 operand planes)

The reference evaluates these with [1]- or [B,1,1,1]-shaped DEVICE tensors as coefficients (separate mul / sub / div kernels,
round-to-nearest each), which is what the expressions below do."""
import pytest
import torch
import torch.nn.functional as F

import gpu_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = [(2, 40, 128), (3, 37, 128), (1, 864, 128)]


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(DEV)


def _c(v):
    """A coefficient as the reference holds it: a one-element fp32 device tensor."""
    return torch.tensor([float(v)], dtype=torch.float32, device=DEV)


def _coefs(n, seed):
    g = torch.Generator().manual_seed(seed)
    return [float(v) for v in (torch.rand(n, generator=g) * 1.5 + 0.05).float()]


@pytest.mark.parametrize("shape", SHAPES)
def test_x0_pred_bit_exact(shape):
    x, eps = _rand(*shape, seed=1, scale=30.0), _rand(*shape, seed=2)
    sigma, alpha = _coefs(2, 3)
    got = G.op_x0_pred(x, eps, sigma, alpha)
    want = (x - _c(sigma) * eps) / _c(alpha)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("order", [1, 2])
def test_dpm_update_bit_exact(shape, order):
    x, m0, m1 = _rand(*shape, seed=4, scale=20.0), _rand(*shape, seed=5, scale=20.0), _rand(*shape, seed=6, scale=20.0)
    cx, cm, hcm, ir0 = _coefs(4, 7)
    got = G.op_dpm_update(x, m0, m1, cx, cm, hcm, ir0, order)
    if order == 1:      # x_t = sigma_t/sigma_s * x - alpha_t*phi_1 * model_s
        want = _c(cx) * x - _c(cm) * m0
    else:               # D1_0 = (1/r0)*(m0 - m1); x_t = ... - 0.5*alpha_t*phi_1 * D1_0
        d1 = _c(ir0) * (m0 - m1)
        want = _c(cx) * x - _c(cm) * m0 - _c(hcm) * d1
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("order", [1, 2])
def test_unipc_predict_correct_bit_exact(shape, order):
    x, m0, m1, mt = (_rand(*shape, seed=s, scale=20.0) for s in (8, 9, 10, 11))
    cx, cmE, aB, rk, rho_p, rho_c0, rho_c1 = _coefs(7, 12)
    xb, xp = G.op_unipc_predict(x, m0, m1, cx, cmE, aB, rk, rho_p, order)
    want_b = _c(cx) * x - _c(cmE) * m0
    assert torch.equal(xb, want_b)
    if order == 1:
        want_p = want_b
    else:
        d1 = (m1 - m0) / _c(rk)
        want_p = want_b - _c(aB) * (_c(rho_p) * d1)
    assert torch.equal(xp, want_p)
    got = G.op_unipc_correct(want_b, m0, m1, mt, aB, rk, rho_c0, rho_c1, order)
    d1_t = mt - m0
    if order == 1:
        want = want_b - _c(aB) * (_c(rho_c1) * d1_t)
    else:
        want = want_b - _c(aB) * (_c(rho_c0) * ((m1 - m0) / _c(rk)) + _c(rho_c1) * d1_t)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
def test_ddpm_step_bit_exact(shape):
    B, T, M = shape
    x, eps = _rand(B, T, M, seed=13), _rand(B, T, M, seed=14)
    noise = _rand(B, M, T, seed=15)                       # reference layout [B,1,M,T]
    cr, crm1, pm1, pm2, sig = _coefs(5, 16)
    got = G.op_ddpm_step(x, eps, noise, cr, crm1, pm1, pm2, sig)
    x0 = (_c(cr) * x - _c(crm1) * eps).clamp(-1.0, 1.0)
    want = (_c(pm1) * x0 + _c(pm2) * x) + _c(sig) * noise.transpose(1, 2)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES)
def test_ddim_step_bit_exact(shape):
    x, eps = _rand(*shape, seed=17, scale=10.0), _rand(*shape, seed=18)
    sat, coef, sap = _coefs(3, 19)
    got = G.op_ddim_step(x, eps, sat, coef, sap)
    want = _c(sap) * (x / _c(sat) + _c(coef) * eps)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES[:2])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_pndm_update_bit_exact(shape, mode):
    x, e, h1, h2, h3 = (_rand(*shape, seed=s) for s in (20, 21, 22, 23, 24))
    d, k1, k2 = _coefs(3, 25)
    got = G.op_pndm_update(x, e, h1, h2, h3, d, k1, k2, mode)
    # Evaluated on the CPU, where the goldens of the executed reference come from: `tensor / python_scalar` is a true division
    # there, while PyTorch's CUDA kernel multiplies by the rounded reciprocal (1/12, 1/24 are inexact) — the library follows
    # the CPU semantics (diffusion.py:158-165).
    xc, ec, a, b, c = (t.cpu() for t in (x, e, h1, h2, h3))
    if mode == 0:
        ep = ec
    elif mode == 1:
        ep = (ec + a) / 2
    elif mode == 2:
        ep = (3 * ec - a) / 2
    elif mode == 3:
        ep = (23 * ec - 16 * a + 5 * b) / 12
    else:
        ep = (55 * ec - 59 * a + 37 * b - 9 * c) / 24
    cc = lambda v: torch.tensor([float(v)], dtype=torch.float32)
    want = xc + cc(d) * (cc(k1) * xc - cc(k2) * ep)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ascale", [1.0, 0.37])
def test_q_sample_bit_exact(shape, ascale):
    B, T, M = shape
    gt = _rand(B, T, M, seed=26, scale=4.0) - 5.0
    noise = _rand(B, 1, M, T, seed=27)
    sa, sb = _coefs(2, 28)
    got = G.op_q_sample(gt, noise.reshape(B, M, T), ascale, sa, sb)
    x0 = (gt * ascale).transpose(1, 2)[:, None, :, :]                    # norm_spec(gt_spec).transpose(1, 2)[:, None]
    want = _c(sa).reshape(1, 1, 1, 1) * x0 + _c(sb).reshape(1, 1, 1, 1) * noise    # q_sample
    assert torch.equal(got, want[:, 0].transpose(1, 2).contiguous())


@pytest.mark.parametrize("parts", [1, 2, 3])
@pytest.mark.parametrize("B,t_in,C,t_out", [(2, 20, 256, 40), (3, 27, 384, 54), (2, 108, 512, 215), (1, 54, 512, 107)])
def test_cast_gather_nearest_upsample_bit_exact(parts, B, t_in, C, t_out):
    x = _rand(B, t_in, C, seed=29, scale=7.0)
    if t_out == 2 * t_in:        # F.interpolate(scale_factor=2.0, mode="nearest")  (resnet.py:157-160)
        up = F.interpolate(x.transpose(1, 2), scale_factor=2.0, mode="nearest")
        scale = 0.5
    else:                        # F.interpolate(size=t_out, mode="nearest") (8 does not divide T, unet_1d_condition.py:795-797)
        up = F.interpolate(x.transpose(1, 2), size=t_out, mode="nearest")
        scale = float(torch.tensor(t_in, dtype=torch.float32) / torch.tensor(t_out, dtype=torch.float32))
    got = G.op_cast_gather(x, t_out, parts, 1, scale)
    want = G.split_planes_ref(up.transpose(1, 2).contiguous(), parts)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))        # bit patterns (parts = 2 holds fp16 bits)


@pytest.mark.parametrize("parts", [1, 2, 3])
@pytest.mark.parametrize("B,t_in,C", [(2, 40, 256), (3, 37, 384), (1, 431, 512)])
def test_cast_gather_stride2_im2col_bit_exact(parts, B, t_in, C):
    x = _rand(B, t_in, C, seed=30, scale=7.0)
    t_out = (t_in - 1) // 2 + 1
    got = G.op_cast_gather(x, t_out, parts, 2)
    xp = F.pad(x, (0, 0, 1, 2))                                         # frames -1 .. t_in+1 (zeros outside)
    taps = torch.stack([xp[:, tap:tap + 2 * t_out:2] for tap in range(3)], dim=2)       # [B, t_out, 3, C]: frame 2*to - 1 + tap
    r = taps.reshape(B, t_out, 3 * C).float()
    want = G.split_planes_ref(r, parts)                                 # element (p, tap, c) at p*3C + tap*C + c
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # and the k=3 / stride-2 / pad-1 convolution it feeds equals F.conv1d on the hi+mid+lo sum
    if parts == 2:
        w = _rand(64, C, 3, seed=31, scale=(3 * C) ** -0.5)
        y_ref = F.conv1d(x.transpose(1, 2).double(), w.double(), stride=2, padding=1).transpose(1, 2)
        cols = G.planes_to_double(got.reshape(B * t_out, 2 * 3 * C), 2, 3 * C).view(B, t_out, 3 * C)   # planes -> values [B, t_out, 3C] (tap-major)
        y = cols @ w.permute(0, 2, 1).reshape(64, 3 * C).double().t()
        assert float((y - y_ref).abs().max()) < 1e-4


@pytest.mark.parametrize("B,C,T", [(2, 128, 40), (1, 128, 861), (3, 96, 33)])
def test_transpose_and_div_copy_bit_exact(B, C, T):
    x = _rand(B, C, T, seed=32)
    y = G.op_transpose(x, True)
    assert torch.equal(y, x.transpose(1, 2).contiguous())
    assert torch.equal(G.op_transpose(y, False), x)
    if (B * C * T) % 4 == 0:
        # denorm_spec = x / acoustic_scale with a Python-float scale (diffusion.py:87,343): a true division on the CPU (where the
        # reference's goldens come from); PyTorch's CUDA kernel would multiply by the rounded reciprocal instead.
        assert torch.equal(G.op_div_copy(x, 0.37).cpu(), x.cpu() / 0.37)
        assert torch.equal(G.op_div_copy(x, 1.0), x)
