"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A compact, functional PyTorch restatement of the reference's Unit2Mel diffusion-sampling
hot path (bfloat16/latent-diffusion-speech).  It exists solely so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs can
check (and time) the CUDA path against the reference's arithmetic on machines where
/root/reference does not exist.  Nothing under ``latent_diffusion_speech_b200/`` imports it.

Parity status: **pinned against the executed reference** — ``oracle/make_golden.py`` runs the
unmodified reference (imported from /root/reference) on seeded inputs and stores its outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those
fixtures bit-for-bit on CPU (same torch build), and ``tests/test_oracle_vs_reference.py``
compares it live against the reference when the tree is present.  The reference itself ships
no golden vectors / tests, so the executed reference is the only authority.

The restatement works on a plain ``state_dict`` (key names of the reference checkpoint) and is
dtype/device generic (fp32 = parity oracle, fp64 = accuracy yardstick).

Reference code followed (file:line relative to the reference tree):
  diffusion/unit2mel.py:73-89                       cond = unit_embed(units) + spk_embed(spk_id-1)
  diffusion/diffusion.py:18-21,46-87,95-121,169-171,189-343   schedule buffers, DDPM step, q_sample, dispatch
  diffusion/dpm_solver_pytorch.py:94-154,271-280,433-442,474,547-576,796-831,1171-1213,1253-1292
  diffusion/uni_pc.py:76-87,118-138,285-294,471-588,606-658
  diffusion/unet1d/unet_1d_condition.py:743-1036    U-Net forward
  diffusion/unet1d/unet_1d_blocks.py:602-623,949-1015,1070-1096,2069-2130,2181-2206
  diffusion/unet1d/resnet.py:150-173,207-223,591-641
  diffusion/unet1d/transformer_1d.py:256-295
  diffusion/unet1d/attention.py:130-203,280-301
  diffusion/unet1d/attention_processor.py:980-1052
  diffusion/unet1d/embeddings.py:24-64,189-201
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
DEFAULT_CFG = dict(out_dims=128, n_hidden=256, block_out_channels=(256, 384, 512, 512),
                   n_layers=2, n_heads=8, groups=8, acoustic_scale=1.0, n_spk=323)
PFX = "decoder.denoise_fn."


# --------------------------------------------------------------------------------------
# U-Net denoiser
# --------------------------------------------------------------------------------------
def timestep_sinusoid(t: Tensor, dim: int) -> Tensor:
    """embeddings.py:24-64 with flip_sin_to_cos=True, freq_shift=0, max_period=1e4."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - 0)
    freq = torch.exp(exponent)
    arg = t[:, None].float() * freq[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def _gn(sd, key, x, groups, eps):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps)


def _conv(sd, key, x, stride=1, padding=1):
    return F.conv1d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride, padding)


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def resnet_block(sd, key, x, emb, groups):
    """resnet.py:591-641 (scale_shift conditioning, output_scale_factor 1)."""
    h = F.silu(_gn(sd, key + ".norm1", x, groups, 1e-5))
    h = _conv(sd, key + ".conv1", h)
    temb = _lin(sd, key + ".time_emb_proj", F.silu(emb))[:, :, None]
    h = _gn(sd, key + ".norm2", h, groups, 1e-5)
    scale, shift = torch.chunk(temb, 2, dim=1)
    h = h * (1 + scale) + shift
    h = _conv(sd, key + ".conv2", F.silu(h))
    if key + ".conv_shortcut.weight" in sd:
        x = _conv(sd, key + ".conv_shortcut", x, padding=0)
    return (x + h) / 1.0


def attention(sd, key, x, heads):
    """attention_processor.py:980-1052 with encoder_hidden_states=None (self-attention)."""
    b, n, c = x.shape
    q, k, v = (_lin(sd, key + s, x) for s in (".to_q", ".to_k", ".to_v"))
    d = c // heads
    q, k, v = (z.view(b, -1, heads, d).transpose(1, 2) for z in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(b, -1, heads * d)
    return _lin(sd, key + ".to_out.0", o)


def transformer_block(sd, key, x, heads):
    """attention.py:130-203: LN->attn1->+, LN->attn2 (self, encoder states are None)->+, LN->GEGLU FF->+."""
    c = x.shape[-1]
    ln = lambda name, z: F.layer_norm(z, (c,), sd[key + name + ".weight"], sd[key + name + ".bias"], 1e-5)
    x = attention(sd, key + ".attn1", ln(".norm1", x), heads) + x
    x = attention(sd, key + ".attn2", ln(".norm2", x), heads) + x
    hv, gate = _lin(sd, key + ".ff.net.0.proj", ln(".norm3", x)).chunk(2, dim=-1)
    ff = _lin(sd, key + ".ff.net.2", hv * F.gelu(gate))
    return ff + x


def transformer_1d(sd, key, x, heads, groups):
    """transformer_1d.py:256-295 (continuous input, conv projections)."""
    res = x
    h = _gn(sd, key + ".norm", x, groups, 1e-6)
    h = _conv(sd, key + ".proj_in", h, padding=0).permute(0, 2, 1)
    h = transformer_block(sd, key + ".transformer_blocks.0", h, heads)
    h = _conv(sd, key + ".proj_out", h.permute(0, 2, 1).contiguous(), padding=0)
    return h + res


def unet_forward(sd: Dict[str, Tensor], cfg: dict, sample: Tensor, timestep: Tensor, pfx: str = PFX) -> Tensor:
    """One denoiser evaluation: sample [B, out_dims+n_hidden, T], timestep [B] -> eps [B, out_dims, T]."""
    sd = _View(sd, pfx)
    ch = list(cfg["block_out_channels"])
    nblk, L, heads, G = len(ch), cfg["n_layers"], cfg["n_heads"], cfg["groups"]
    T = sample.shape[-1]
    force_size = T % (2 ** (nblk - 1)) != 0            # unet_1d_condition.py:789-797

    if timestep.ndim == 0:
        timestep = timestep[None]
    timestep = timestep.expand(sample.shape[0])
    t_emb = timestep_sinusoid(timestep, ch[0]).to(sample.dtype)
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t_emb)))

    h = _conv(sd, "conv_in", sample)
    skips: List[Tensor] = [h]
    for i in range(nblk):
        last = i == nblk - 1
        for j in range(L):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, emb, G)
            if not last:
                h = transformer_1d(sd, f"down_blocks.{i}.attentions.{j}", h, heads, G)
            skips.append(h)
        if not last:
            h = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2)
            skips.append(h)

    h = resnet_block(sd, "mid_block.resnets.0", h, emb, G)
    h = transformer_1d(sd, "mid_block.attentions.0", h, heads, G)
    h = resnet_block(sd, "mid_block.resnets.1", h, emb, G)

    for i in range(nblk):
        last = i == nblk - 1
        for j in range(L + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, emb, G)
            if i > 0:
                h = transformer_1d(sd, f"up_blocks.{i}.attentions.{j}", h, heads, G)
        if not last:
            if force_size:
                h = F.interpolate(h, size=skips[-1].shape[2:], mode="nearest")
            else:
                h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h)

    h = F.silu(_gn(sd, "conv_norm_out", h, G, 1e-5))
    return _conv(sd, "conv_out", h)


class _View(dict):
    """Prefix view over a state_dict."""

    def __init__(self, sd, pfx):
        super().__init__()
        self._sd, self._pfx = sd, pfx

    def __getitem__(self, k):
        return self._sd[self._pfx + k]

    def __contains__(self, k):
        return (self._pfx + k) in self._sd

    def get(self, k, default=None):
        return self._sd.get(self._pfx + k, default)


# --------------------------------------------------------------------------------------
# Noise schedule and samplers
# --------------------------------------------------------------------------------------
def diffusion_buffers(timesteps: int = 1000, max_beta: float = 0.02) -> Dict[str, Tensor]:
    """diffusion.py:46-84 — fp64 numpy, cast to fp32 buffers."""
    betas = np.linspace(1e-4, max_beta, timesteps)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(
        betas=f32(betas), alphas_cumprod=f32(ac), alphas_cumprod_prev=f32(ac_prev),
        sqrt_alphas_cumprod=f32(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1.0 - ac)),
        log_one_minus_alphas_cumprod=f32(np.log(1.0 - ac)),
        sqrt_recip_alphas_cumprod=f32(np.sqrt(1.0 / ac)), sqrt_recipm1_alphas_cumprod=f32(np.sqrt(1.0 / ac - 1)),
        posterior_variance=f32(pv), posterior_log_variance_clipped=f32(np.log(np.maximum(pv, 1e-20))),
        posterior_mean_coef1=f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        posterior_mean_coef2=f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    )


def piecewise_linear(x: Tensor, xp: Tensor, yp: Tensor) -> Tensor:
    """dpm_solver_pytorch.py:1253-1292 for C=1: x [N], xp/yp [K]; linear extrapolation at both ends.
    A query equal to a knot sorts *before* the knot (x is element 0 of the concatenation)."""
    K = xp.shape[0]
    idx = torch.searchsorted(xp, x, right=False)          # number of knots strictly below x
    lo = torch.where(idx == 0, torch.zeros_like(idx), torch.where(idx == K, torch.full_like(idx, K - 2), idx - 1))
    x0, x1, y0, y1 = xp[lo], xp[lo + 1], yp[lo], yp[lo + 1]
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0)


class VPSchedule:
    """Discrete VP noise schedule (dpm_solver_pytorch.py:94-154; uni_pc.py:76-138).
    ``clip`` reproduces numerical_clip_alpha (DPM copy only; inactive for the linear betas)."""

    def __init__(self, betas: Tensor, clip: bool, dtype=torch.float32):
        log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        if clip:
            log_sig = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
            lambs = log_alphas - log_sig
            n_drop = int(torch.searchsorted(torch.flip(lambs, [0]), torch.tensor(-5.1, dtype=lambs.dtype, device=lambs.device)))
            if n_drop > 0:
                log_alphas = log_alphas[:-n_drop]
        self.log_alpha = log_alphas.to(dtype)
        self.N = self.log_alpha.shape[0]
        self.t_knots = torch.linspace(0.0, 1.0, self.N + 1)[1:].to(dtype)

    def log_mean(self, t):
        return piecewise_linear(t.reshape(-1), self.t_knots.to(t.device), self.log_alpha.to(t.device))

    def alpha(self, t):
        return torch.exp(self.log_mean(t))

    def std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_mean(t)))

    def lam(self, t):
        lm = self.log_mean(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))


def _x0_pred(eps_fn, ns: VPSchedule, x, t):
    """data_prediction_fn (dpm_solver_pytorch.py:433-442) with the discrete-time wrapper (:271-280)."""
    t_in = (t.expand(x.shape[0]) - 1.0 / ns.N) * ns.N
    eps = eps_fn(x, t_in)
    return (x - ns.std(t) * eps) / ns.alpha(t)


def sample_dpm_solver_pp(eps_fn, betas: Tensor, x: Tensor, steps: int) -> Tensor:
    """DPM-Solver++ multistep order 2, time_uniform (dpm_solver_pytorch.py:1171-1213,547-576,796-831)."""
    ns = VPSchedule(betas, clip=True)
    ts = torch.linspace(1.0, 1.0 / ns.N, steps + 1).to(x.device)
    assert steps >= 2

    def first(x, s, t, m_s):
        h = ns.lam(t) - ns.lam(s)
        return ns.std(t) / ns.std(s) * x - torch.exp(ns.log_mean(t)) * torch.expm1(-h) * m_s

    def second(x, m1, m0, s1, s0, t):
        l1, l0, lt = ns.lam(s1), ns.lam(s0), ns.lam(t)
        alpha_t = torch.exp(ns.log_mean(t))
        h0, h = l0 - l1, lt - l0
        r0 = h0 / h
        d1 = (1.0 / r0) * (m0 - m1)
        phi = torch.expm1(-h)
        return (ns.std(t) / ns.std(s0)) * x - (alpha_t * phi) * m0 - 0.5 * (alpha_t * phi) * d1

    t_hist = [ts[0]]
    m_hist = [_x0_pred(eps_fn, ns, x, ts[0])]
    x = first(x, t_hist[-1], ts[1], m_hist[-1])
    t_hist.append(ts[1])
    m_hist.append(_x0_pred(eps_fn, ns, x, ts[1]))
    for k in range(2, steps + 1):
        t = ts[k]
        order = min(2, steps + 1 - k) if steps < 10 else 2
        if order == 1:
            x = first(x, t_hist[-1], t, m_hist[-1])
        else:
            x = second(x, m_hist[-2], m_hist[-1], t_hist[-2], t_hist[-1], t)
        t_hist = [t_hist[-1], t]
        m_hist = [m_hist[-1], m_hist[-1]]
        if k < steps:
            m_hist[-1] = _x0_pred(eps_fn, ns, x, t)
    return x


def sample_unipc_bh2(eps_fn, betas: Tensor, x: Tensor, steps: int) -> Tensor:
    """UniPC-bh2, data prediction, multistep order 2, lower_order_final (uni_pc.py:471-588,606-658)."""
    ns = VPSchedule(betas, clip=False)
    ts = torch.linspace(1.0, 1.0 / ns.N, steps + 1).to(x.device)
    assert steps >= 2

    def update(x, m_hist, t_hist, t, order, corrector):
        t = t.view(-1)
        s0, m0 = t_hist[-1], m_hist[-1]
        lam0, lamt = ns.lam(s0), ns.lam(t)
        sig0, sigt = ns.std(s0), ns.std(t)
        alpha_t = torch.exp(ns.log_mean(t))
        h = lamt - lam0
        rks, d1s = [], []
        for i in range(1, order):
            rk = (ns.lam(t_hist[-(i + 1)]) - lam0) / h
            rks.append(rk)
            d1s.append((m_hist[-(i + 1)] - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks, device=x.device)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        R, b = torch.stack(R), torch.cat(b)
        rho_p = torch.tensor([0.5], device=b.device) if order == 2 else None
        if corrector:
            rho_c = torch.tensor([0.5], device=b.device) if order == 1 else torch.linalg.solve(R, b)
        x_base = sigt / sig0 * x - alpha_t * h_phi_1 * m0
        if d1s:
            D = torch.stack(d1s, dim=1)
            pred = torch.einsum("k,bkchw->bchw", rho_p, D)
        else:
            D, pred = None, 0
        x_t = x_base - alpha_t * B_h * pred
        m_t = None
        if corrector:
            m_t = _x0_pred(eps_fn, ns, x_t, t)
            corr = torch.einsum("k,bkchw->bchw", rho_c[:-1], D) if D is not None else 0
            x_t = x_base - alpha_t * B_h * (corr + rho_c[-1] * (m_t - m0))
        return x_t, m_t

    t_hist = [ts[0]]
    m_hist = [_x0_pred(eps_fn, ns, x, ts[0])]
    x, m = update(x, m_hist, t_hist, ts[1], 1, True)
    t_hist.append(ts[1])
    m_hist.append(m)
    for k in range(2, steps + 1):
        order = min(2, steps + 1 - k)
        x, m = update(x, m_hist, t_hist, ts[k], order, k < steps)
        t_hist = [t_hist[-1], ts[k]]
        m_hist = [m_hist[-1], m_hist[-1]]
        if k < steps:
            m_hist[-1] = m
    return x


def sample_ddpm(eps_fn_int, buf: Dict[str, Tensor], x: Tensor, k_step: int, noises: Sequence[Tensor]) -> Tensor:
    """Ancestral sampling (diffusion.py:104-121,335-341).  noises[j] is the draw of loop iteration j."""
    b = x.shape[0]
    for j, i in enumerate(reversed(range(0, k_step))):
        t = torch.full((b,), i, device=x.device, dtype=torch.long)
        ex = lambda name: buf[name].to(x.device).gather(-1, t).reshape(b, 1, 1, 1)
        eps = eps_fn_int(x, t)
        x0 = ex("sqrt_recip_alphas_cumprod") * x - ex("sqrt_recipm1_alphas_cumprod") * eps
        x0.clamp_(-1.0, 1.0)
        mean = ex("posterior_mean_coef1") * x0 + ex("posterior_mean_coef2") * x
        logvar = ex("posterior_log_variance_clipped")
        mask = (1 - (t == 0).float()).reshape(b, 1, 1, 1)
        x = mean + mask * (0.5 * logvar).exp() * noises[j].to(x.device)
    return x


def sample_ddim(eps_fn_int, buf: Dict[str, Tensor], x: Tensor, t_total: int, interval: int) -> Tensor:
    """DDIM (eta = 0) over t = reversed(range(0, t_total, interval))  (diffusion.py:123-132,317-332).
    Note a_prev = alphas_cumprod[max(t - interval, 0)]: the last step lands on alphas_cumprod[0], not on 1."""
    b = x.shape[0]
    acp = buf["alphas_cumprod"].to(x.device)
    for i in reversed(range(0, t_total, interval)):
        t = torch.full((b,), i, device=x.device, dtype=torch.long)
        a_t = acp.gather(-1, t).reshape(b, 1, 1, 1)
        a_prev = acp.gather(-1, torch.max(t - interval, torch.zeros_like(t))).reshape(b, 1, 1, 1)
        eps = eps_fn_int(x, t)
        x = a_prev.sqrt() * (x / a_t.sqrt() + (((1 - a_prev) / a_prev).sqrt() - ((1 - a_t) / a_t).sqrt()) * eps)
    return x


def sample_pndm(eps_fn_int, buf: Dict[str, Tensor], x: Tensor, t_total: int, interval: int) -> Tensor:
    """PLMS / PNDM (diffusion.py:134-167,300-316): Adams-Bashforth combination of up to four noise predictions, the
    first step bootstrapped with a second evaluation at t - interval.  The reference evaluates
    ``max(t - interval, 0)`` on a [B] tensor (diffusion.py:155), which only works for B == 1; the restatement applies
    the same element-wise maximum for any B (identical for B == 1)."""
    b = x.shape[0]
    acp = buf["alphas_cumprod"].to(x.device)
    hist: list = []                                    # deque(maxlen=4) of the reference; only the last 3 are read

    def x_pred(x, noise_t, t):
        a_t = acp.gather(-1, t).reshape(b, 1, 1, 1)
        a_prev = acp.gather(-1, torch.max(t - interval, torch.zeros_like(t))).reshape(b, 1, 1, 1)
        a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
        x_delta = (a_prev - a_t) * ((1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x - 1 / (
            a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * noise_t)
        return x + x_delta

    for i in reversed(range(0, t_total, interval)):
        t = torch.full((b,), i, device=x.device, dtype=torch.long)
        e = eps_fn_int(x, t)
        if len(hist) == 0:
            xp = x_pred(x, e, t)
            e_prev = eps_fn_int(xp, torch.max(t - interval, torch.zeros_like(t)))
            e_prime = (e + e_prev) / 2
        elif len(hist) == 1:
            e_prime = (3 * e - hist[-1]) / 2
        elif len(hist) == 2:
            e_prime = (23 * e - 16 * hist[-1] + 5 * hist[-2]) / 12
        else:
            e_prime = (55 * e - 59 * hist[-1] + 37 * hist[-2] - 9 * hist[-3]) / 24
        x = x_pred(x, e_prime, t)
        hist.append(e)
        hist = hist[-4:]
    return x


def q_sample(buf, x_start, t_index: int, noise):
    """diffusion.py:169-171 at a single integer t for the whole batch."""
    return buf["sqrt_alphas_cumprod"][t_index].to(x_start.device) * x_start + \
        buf["sqrt_one_minus_alphas_cumprod"][t_index].to(x_start.device) * noise


# --------------------------------------------------------------------------------------
# Unit2Mel front door
# --------------------------------------------------------------------------------------
def unit2mel_cond(sd, units: Tensor, spk_id: Optional[Tensor], n_spk: Optional[int]) -> Tensor:
    """unit2mel.py:79-82 (volume None -> +0)."""
    x = F.linear(units, sd["unit_embed.weight"], sd["unit_embed.bias"]) + 0
    if n_spk is not None and n_spk > 1:
        x = x + F.embedding(spk_id - 1, sd["spk_embed.weight"])
    return x


def unit2mel_infer(sd: Dict[str, Tensor], cfg: dict, units: Tensor, spk_id: Optional[Tensor], noise: Tensor,
                   method: Optional[str] = "dpm-solver", infer_speedup: int = 50,
                   gt_spec: Optional[Tensor] = None, k_step: Optional[int] = None,
                   step_noises: Optional[Sequence[Tensor]] = None, dtype=None) -> Tensor:
    """Unit2Mel.forward(infer=True) -> [B, T, out_dims]  (unit2mel.py:73-89, diffusion.py:189-343).

    noise       [B,1,M,T]: the initial ``randn`` (or the q_sample noise for shallow diffusion).
    step_noises list of [B,1,M,T], one per DDPM iteration (infer_speedup == 1)."""
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        units, noise = units.to(dtype), noise.to(dtype)
        gt_spec = None if gt_spec is None else gt_spec.to(dtype)
    buf = diffusion_buffers()
    cond = unit2mel_cond(sd, units, spk_id, cfg.get("n_spk")).transpose(1, 2)      # [B,H,T]
    scale = cfg.get("acoustic_scale", 1.0)
    if gt_spec is None or k_step is None:
        t_total, x = 1000, noise
    else:
        t_total = k_step
        x0 = (gt_spec * scale).transpose(1, 2)[:, None, :, :]
        x = q_sample(buf, x0, k_step - 1, noise)
    calc = x.dtype

    def eps_fn(xx, t_in):
        inp = torch.cat([xx[:, 0, :, :], cond], dim=-2)
        return unet_forward(sd, cfg, inp, t_in)[:, None, :, :]

    betas = buf["betas"][:t_total].to(x.device)
    if calc == torch.float64:
        betas = torch.tensor(np.linspace(1e-4, 0.02, 1000)[:t_total], dtype=torch.float64, device=x.device)
    if method is not None and infer_speedup > 1:
        steps = t_total // infer_speedup
        if method == "dpm-solver":
            x = sample_dpm_solver_pp(eps_fn, betas, x, steps)
        elif method == "unipc":
            x = sample_unipc_bh2(eps_fn, betas, x, steps)
        elif method in ("ddim", "pndm"):
            bb = buf if calc != torch.float64 else {k: v.double() for k, v in buf.items()}
            x = (sample_ddim if method == "ddim" else sample_pndm)(eps_fn, bb, x, t_total, infer_speedup)
        else:
            raise NotImplementedError(method)
    else:
        x = sample_ddpm(eps_fn, buf if calc != torch.float64 else {k: v.double() for k, v in buf.items()},
                        x, t_total, step_noises)
    return x.squeeze(1).transpose(1, 2) / scale


def unit2mel_train_loss(sd: Dict[str, Tensor], cfg: dict, units: Tensor, spk_id: Optional[Tensor], gt_spec: Tensor, t: Tensor, noise: Tensor,
                        loss_type: str = "l2", dtype=None, return_eps: bool = False):
    """Unit2Mel.forward(infer=False) -> scalar loss (unit2mel.py:73-89 -> diffusion.py:193-201 -> p_losses :173-187) with the
    ``randint`` timesteps t [B] (int64) and the ``randn_like`` noise [B,1,M,T] injected."""
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        units, noise, gt_spec = units.to(dtype), noise.to(dtype), gt_spec.to(dtype)
    buf = diffusion_buffers()
    cond = unit2mel_cond(sd, units, spk_id, cfg.get("n_spk")).transpose(1, 2)      # [B,H,T]
    x_start = (gt_spec * cfg.get("acoustic_scale", 1.0)).transpose(1, 2)[:, None, :, :]
    tt = t.to("cpu").long()
    sa = buf["sqrt_alphas_cumprod"][tt].reshape(-1, 1, 1, 1).to(device=x_start.device, dtype=x_start.dtype)       # extract (diffusion.py:18-21)
    sb = buf["sqrt_one_minus_alphas_cumprod"][tt].reshape(-1, 1, 1, 1).to(device=x_start.device, dtype=x_start.dtype)
    x_noisy = sa * x_start + sb * noise                                              # q_sample (diffusion.py:169-171)
    denoise_input = torch.cat([x_noisy[:, 0, :, :], cond], dim=-2)
    x_recon = unet_forward(sd, cfg, denoise_input, t.to(x_start.device))[:, None, :, :]
    if loss_type == "l1":
        loss = (noise - x_recon).abs().mean()
    elif loss_type == "l2":
        loss = F.mse_loss(noise, x_recon)
    else:
        raise NotImplementedError()
    return (loss, x_recon[:, 0]) if return_eps else loss


def synthetic_inputs(B: int, T: int, seed: int = 7, in_dims: int = 1280, n_spk: int = 323, out_dims: int = 128,
                     noise_seed: int = 1000, n_step_noises: int = 0, gt: bool = False):
    """SURVEY.md §8(d) synthetic inputs with per-utterance seeded noise (shard-invariant)."""
    g = torch.Generator().manual_seed(seed)
    units = torch.randn(B, T, in_dims, generator=g)
    spk_id = torch.randint(1, n_spk + 1, (B, 1), generator=g)
    noise = torch.empty(B, 1, out_dims, T)
    steps = [torch.empty(B, 1, out_dims, T) for _ in range(n_step_noises)]
    for b in range(B):
        gb = torch.Generator().manual_seed(noise_seed + b)
        noise[b] = torch.randn(1, out_dims, T, generator=gb)
        for s in steps:
            s[b] = torch.randn(1, out_dims, T, generator=gb)
    gt_spec = (torch.rand(B, T, out_dims, generator=g) * 14 - 12) if gt else None
    return units, spk_id, noise, steps, gt_spec
