"""Host-side, batch-invariant scalar work of the samplers.

In the reference every per-step coefficient is recomputed on the device with dozens of tiny
kernels (sort/gather based interpolation, ``expm1``, a 2x2 ``linalg.solve`` ...), although the
timestep is one scalar shared by the whole batch.  Here the complete sampler
program — the timestep fed to the denoiser at every evaluation and the scalar coefficients of
every state update — is computed once per ``(method, steps, k_step)`` on the host and handed to
the CUDA library (``lds_plan``), which then runs one fused update kernel per step.

The arithmetic is done with fp32 torch CPU ops in the reference's own evaluation order so the
coefficients are the very floats the reference's CPU run produces:
  diffusion/dpm_solver_pytorch.py:94-154,271-280,474,547-576,796-831,1171-1213,1253-1292
  diffusion/uni_pc.py:76-138,471-588,606-658
  diffusion/diffusion.py:95-121 (DDPM ancestral step), :123-167 (DDIM, PLMS), embeddings.py:24-64 (timestep sinusoid)
Row layouts are documented in include/lds_b200.h.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

COEF_STRIDE = 12
SAMPLER_DPMPP_2M, SAMPLER_UNIPC_BH2, SAMPLER_DDPM, SAMPLER_DDIM, SAMPLER_PNDM = 0, 1, 2, 3, 4


def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """[n] timesteps -> [n, dim] = [cos(t w) | sin(t w)], w_i = exp(-ln(1e4) i / (dim/2))."""
    half = dim // 2
    w = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    arg = t.detach().to("cpu")[:, None].float() * w[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).contiguous()


class DiscreteVPSchedule:
    """log(alpha_t) as a piecewise-linear function of continuous time over knots t_i = (i+1)/N."""

    def __init__(self, betas: torch.Tensor, clip_lambda: float | None):
        betas = betas.detach().to("cpu", torch.float32)
        log_alpha = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        if clip_lambda is not None:   # numerical_clip_alpha of the DPM-Solver copy
            lam = log_alpha - 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alpha))
            cut = int(torch.searchsorted(torch.flip(lam, [0]), torch.tensor(clip_lambda)))
            if cut > 0:
                log_alpha = log_alpha[:-cut]
        self.log_alpha = log_alpha
        self.n = int(log_alpha.shape[0])
        self.knots = torch.linspace(0.0, 1.0, self.n + 1)[1:]

    def log_alpha_at(self, t: torch.Tensor) -> torch.Tensor:
        t = t.reshape(-1)
        k = self.n
        below = torch.searchsorted(self.knots, t, right=False)   # ties: the query sorts first
        seg = torch.clamp(below - 1, 0, k - 2)
        xa, xb = self.knots[seg], self.knots[seg + 1]
        ya, yb = self.log_alpha[seg], self.log_alpha[seg + 1]
        return ya + (t - xa) * (yb - ya) / (xb - xa)

    def alpha(self, t):
        return torch.exp(self.log_alpha_at(t))

    def sigma(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha_at(t)))

    def lam(self, t):
        la = self.log_alpha_at(t)
        return la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))

    def model_time(self, t):
        return (t - 1.0 / self.n) * self.n


def _rows(n: int) -> np.ndarray:
    return np.zeros((n, COEF_STRIDE), dtype=np.float32)


def dpm_solver_pp_program(betas: torch.Tensor, steps: int) -> Tuple[torch.Tensor, np.ndarray]:
    """DPM-Solver++(2M), time_uniform.  Returns (denoiser timesteps [steps], coefficient rows [steps+1, 12])."""
    if steps < 2:
        raise ValueError("DPM-Solver++ order 2 needs at least 2 steps")
    ns = DiscreteVPSchedule(betas, clip_lambda=-5.1)
    ts = torch.linspace(1.0, 1.0 / ns.n, steps + 1)
    rows = _rows(steps + 1)
    f = lambda v: float(v.reshape(-1)[0])
    for k in range(steps + 1):
        t = ts[k]
        rows[k, 0], rows[k, 1] = f(ns.sigma(t)), f(ns.alpha(t))
        if k == 0:
            continue
        s0 = ts[k - 1]
        h = ns.lam(t) - ns.lam(s0)
        alpha_t = torch.exp(ns.log_alpha_at(t))
        phi_1 = torch.expm1(-h)
        order = 1 if k == 1 else (min(2, steps + 1 - k) if steps < 10 else 2)
        rows[k, 2] = f(ns.sigma(t) / ns.sigma(s0))
        rows[k, 3] = f(alpha_t * phi_1)
        rows[k, 6] = order
        if order == 2:
            h0 = ns.lam(s0) - ns.lam(ts[k - 2])
            rows[k, 4] = f(0.5 * (alpha_t * phi_1))
            rows[k, 5] = f(1.0 / (h0 / h))
    return ns.model_time(ts[:steps]), rows


def unipc_bh2_program(betas: torch.Tensor, steps: int) -> Tuple[torch.Tensor, np.ndarray]:
    """UniPC-bh2 (data prediction, order 2, lower_order_final).  Same return convention."""
    if steps < 2:
        raise ValueError("UniPC order 2 needs at least 2 steps")
    ns = DiscreteVPSchedule(betas, clip_lambda=None)
    ts = torch.linspace(1.0, 1.0 / ns.n, steps + 1)
    rows = _rows(steps + 1)
    f = lambda v: float(v.reshape(-1)[0])
    for k in range(steps + 1):
        t = ts[k].view(-1)
        rows[k, 0], rows[k, 1] = f(ns.sigma(t)), f(ns.alpha(t))
        if k == 0:
            continue
        order = 1 if k == 1 else min(2, steps + 1 - k)
        corrector = k < steps
        s0 = ts[k - 1]
        lam0 = ns.lam(s0)
        h = ns.lam(t) - lam0
        alpha_t = torch.exp(ns.log_alpha_at(t))
        rks = []
        if order == 2:
            rk = (ns.lam(ts[k - 2]) - lam0) / h
            rks.append(rk)
            rows[k, 5] = f(rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        b_h = torch.expm1(hh)
        r_mat, b_vec, fact = [], [], 1
        for i in range(1, order + 1):
            r_mat.append(torch.pow(rks, i - 1))
            b_vec.append(h_phi_k * fact / b_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        r_mat, b_vec = torch.stack(r_mat), torch.cat(b_vec)
        rows[k, 2] = f(ns.sigma(t) / ns.sigma(s0))
        rows[k, 3] = f(alpha_t * h_phi_1)
        rows[k, 4] = f(alpha_t * b_h)
        rows[k, 6], rows[k, 7], rows[k, 10] = order, float(corrector), 0.5
        if corrector:
            rho_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(r_mat, b_vec)
            rows[k, 8], rows[k, 9] = f(rho_c[0]), f(rho_c[-1])
    return ns.model_time(ts[:steps]), rows


def ddpm_program(buffers: Dict[str, torch.Tensor], k_step: int) -> Tuple[torch.Tensor, np.ndarray]:
    """Ancestral sampling t = k_step-1 ... 0.  Returns (integer timesteps as float [k_step], rows [k_step, 12])."""
    rows = _rows(k_step)
    b = {k: v.detach().to("cpu", torch.float32) for k, v in buffers.items()}
    ts = torch.arange(k_step - 1, -1, -1)
    for j, i in enumerate(ts.tolist()):
        rows[j, 0] = float(b["sqrt_recip_alphas_cumprod"][i])
        rows[j, 1] = float(b["sqrt_recipm1_alphas_cumprod"][i])
        rows[j, 2] = float(b["posterior_mean_coef1"][i])
        rows[j, 3] = float(b["posterior_mean_coef2"][i])
        mask = torch.tensor(0.0 if i == 0 else 1.0)
        rows[j, 4] = float(mask * (0.5 * b["posterior_log_variance_clipped"][i]).exp())
    return ts.float(), rows


def _strided_timesteps(t_total: int, interval: int):
    return list(reversed(range(0, t_total, interval)))


def ddim_program(alphas_cumprod: torch.Tensor, t_total: int, interval: int) -> Tuple[torch.Tensor, np.ndarray]:
    """DDIM, eta = 0 (diffusion.py:123-132,317-332).  Returns (integer timesteps as float [steps], rows [steps, 12])."""
    acp = alphas_cumprod.detach().to("cpu", torch.float32)
    ts = _strided_timesteps(t_total, interval)
    rows = _rows(len(ts))
    for j, i in enumerate(ts):
        a_t, a_prev = acp[i], acp[max(i - interval, 0)]
        rows[j, 0] = float(a_t.sqrt())
        rows[j, 1] = float(((1 - a_prev) / a_prev).sqrt() - ((1 - a_t) / a_t).sqrt())
        rows[j, 2] = float(a_prev.sqrt())
    return torch.tensor(ts, dtype=torch.float32), rows


def pndm_program(alphas_cumprod: torch.Tensor, t_total: int, interval: int) -> Tuple[torch.Tensor, np.ndarray]:
    """PLMS / PNDM (diffusion.py:134-167,300-316).  Denoiser timesteps [steps+1]: t_0, max(t_0-interval,0), t_1, t_2, ...
    (the first step is bootstrapped with a second evaluation); rows [steps, 12]."""
    acp = alphas_cumprod.detach().to("cpu", torch.float32)
    ts = _strided_timesteps(t_total, interval)
    rows = _rows(len(ts))
    for j, i in enumerate(ts):
        a_t, a_prev = acp[i], acp[max(i - interval, 0)]
        a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
        rows[j, 0] = float(a_prev - a_t)
        rows[j, 1] = float(1 / (a_t_sq * (a_t_sq + a_prev_sq)))
        rows[j, 2] = float(1 / (a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())))
    evals = [ts[0], max(ts[0] - interval, 0)] + ts[1:]
    return torch.tensor(evals, dtype=torch.float32), rows
