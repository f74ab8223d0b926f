"""CPU: the host-computed sampler programs + the update formulas of csrc/solver.cu (emulated with
torch fp32 ops in the kernels' evaluation order) reproduce the oracle's samplers bit-for-bit when
both are driven by the same toy eps-model."""
import numpy as np
import pytest
import torch

from latent_diffusion_speech_b200 import sampler_tables as st
from oracle import unit2mel_oracle as O


def toy_eps(x, t_in):
    # deterministic, nonlinear in x and t (t_in is the model-time the wrapper feeds the denoiser)
    return torch.sin(x * 0.37 + t_in.reshape(-1, 1, 1, 1) * 0.011) * 0.8 + 0.05 * x


def emulate(kind, t_in, rows, x, noises=None):
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    m = [None, None]       # m0 (latest), m1 (previous)
    S = t_in.shape[0]
    B = x.shape[0]

    def x0_pred(xx, k):
        eps = toy_eps(xx, t_in[k].expand(B))
        return (xx - f(rows[k, 0]) * eps) / f(rows[k, 1])

    if kind == st.SAMPLER_DDPM:
        for j in range(S):
            r = rows[j]
            eps = toy_eps(x, t_in[j].expand(B))
            x0 = (f(r[0]) * x - f(r[1]) * eps).clamp(-1.0, 1.0)
            x = (f(r[2]) * x0 + f(r[3]) * x) + f(r[4]) * noises[j]
        return x
    m[0] = x0_pred(x, 0)
    for k in range(1, S + 1):
        r = rows[k]
        order = int(r[6])
        if kind == st.SAMPLER_DPMPP_2M:
            if order == 1:
                x = f(r[2]) * x - f(r[3]) * m[0]
            else:
                d1 = f(r[5]) * (m[0] - m[1])
                x = f(r[2]) * x - f(r[3]) * m[0] - f(r[4]) * d1
            if k < S:
                m = [x0_pred(x, k), m[0]]
        else:
            xb = f(r[2]) * x - f(r[3]) * m[0]
            xp = xb if order == 1 else xb - f(r[4]) * (f(r[10]) * ((m[1] - m[0]) / f(r[5])))
            if int(r[7]):
                mt = x0_pred(xp, k)
                if order == 1:
                    x = xb - f(r[4]) * (f(r[9]) * (mt - m[0]))
                else:
                    x = xb - f(r[4]) * (f(r[8]) * ((m[1] - m[0]) / f(r[5])) + f(r[9]) * (mt - m[0]))
                m = [mt, m[0]]
            else:
                x = xp
    return x


@pytest.mark.parametrize("method,steps,n", [("dpm-solver", 20, 1000), ("dpm-solver", 8, 1000), ("dpm-solver", 10, 1000),
                                            ("dpm-solver", 20, 100), ("unipc", 10, 1000), ("unipc", 20, 1000),
                                            ("unipc", 10, 100), ("unipc", 2, 1000), ("dpm-solver", 2, 1000)])
def test_solver_program_matches_oracle(method, steps, n):
    torch.manual_seed(0)
    x = torch.randn(2, 1, 8, 12)
    betas = O.diffusion_buffers()["betas"][:n]
    if method == "dpm-solver":
        t_in, rows = st.dpm_solver_pp_program(betas, steps)
        want = O.sample_dpm_solver_pp(toy_eps, betas, x.clone(), steps)
        kind = st.SAMPLER_DPMPP_2M
    else:
        t_in, rows = st.unipc_bh2_program(betas, steps)
        want = O.sample_unipc_bh2(toy_eps, betas, x.clone(), steps)
        kind = st.SAMPLER_UNIPC_BH2
    assert rows.shape == (steps + 1, st.COEF_STRIDE) and t_in.shape == (steps,)
    got = emulate(kind, t_in, rows, x.clone())
    assert torch.equal(got, want), float((got - want).abs().max())


def test_ddpm_program_matches_oracle():
    torch.manual_seed(1)
    buf = O.diffusion_buffers()
    k_step = 25
    x = torch.randn(2, 1, 8, 12)
    noises = [torch.randn(2, 1, 8, 12) for _ in range(k_step)]
    t_in, rows = st.ddpm_program(buf, k_step)
    want = O.sample_ddpm(lambda xx, t: toy_eps(xx, t.float()), buf, x.clone(), k_step, noises)
    got = emulate(st.SAMPLER_DDPM, t_in, rows, x.clone(), noises)
    assert torch.equal(got, want)
    assert rows[-1, 4] == 0.0 and t_in[-1] == 0.0


def emulate_ddim(t_in, rows, x):
    """csrc/solver.cu ddim_step_kernel: x = c2 * (x / c0 + c1 * eps)."""
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    B = x.shape[0]
    for j in range(rows.shape[0]):
        eps = toy_eps(x, t_in[j].expand(B))
        x = f(rows[j, 2]) * (x / f(rows[j, 0]) + f(rows[j, 1]) * eps)
    return x


def emulate_pndm(t_in, rows, x):
    """lds_api.cu PNDM branch + csrc/solver.cu pndm_update_kernel (modes 0..4, ring of noise predictions)."""
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    B = x.shape[0]
    upd = lambda xx, ep, r: xx + f(r[0]) * (f(r[1]) * xx - f(r[2]) * ep)
    hist = []
    for k in range(rows.shape[0]):
        r = rows[k]
        e = toy_eps(x, t_in[0 if k == 0 else k + 1].expand(B))
        if k == 0:
            xp = upd(x, e, r)
            e2 = toy_eps(xp, t_in[1].expand(B))
            ep = (e + e2) / 2.0
        elif k == 1:
            ep = (3.0 * e - hist[-1]) / 2.0
        elif k == 2:
            ep = ((23.0 * e - 16.0 * hist[-1]) + 5.0 * hist[-2]) / 12.0
        else:
            ep = (((55.0 * e - 59.0 * hist[-1]) + 37.0 * hist[-2]) - 9.0 * hist[-3]) / 24.0
        x = upd(x, ep, r)
        hist = (hist + [e])[-3:]
    return x


@pytest.mark.parametrize("method,t_total,interval", [("ddim", 1000, 50), ("ddim", 1000, 30), ("ddim", 100, 10), ("ddim", 1000, 500),
                                                     ("pndm", 1000, 50), ("pndm", 1000, 30), ("pndm", 100, 10), ("pndm", 1000, 500),
                                                     ("pndm", 100, 100)])
def test_ddim_pndm_program_matches_oracle(method, t_total, interval):
    torch.manual_seed(2)
    buf = O.diffusion_buffers()
    x = torch.randn(2, 1, 8, 12)
    steps = len(range(0, t_total, interval))
    eps_int = lambda xx, t: toy_eps(xx, t.float())
    if method == "ddim":
        t_in, rows = st.ddim_program(buf["alphas_cumprod"], t_total, interval)
        assert rows.shape == (steps, st.COEF_STRIDE) and t_in.shape == (steps,)
        got, want = emulate_ddim(t_in, rows, x.clone()), O.sample_ddim(eps_int, buf, x.clone(), t_total, interval)
    else:
        t_in, rows = st.pndm_program(buf["alphas_cumprod"], t_total, interval)
        assert rows.shape == (steps, st.COEF_STRIDE) and t_in.shape == (steps + 1,)
        got, want = emulate_pndm(t_in, rows, x.clone()), O.sample_pndm(eps_int, buf, x.clone(), t_total, interval)
    assert torch.equal(got, want), float((got - want).abs().max())


def test_timestep_sinusoid_matches_oracle():
    t = torch.tensor([0.0, 1.0, 417.25, 999.0, 998.001])
    assert torch.equal(st.timestep_sinusoid(t, 256), O.timestep_sinusoid(t, 256))


def test_piecewise_linear_extrapolates_like_reference_semantics():
    ns = st.DiscreteVPSchedule(O.diffusion_buffers()["betas"][:100], clip_lambda=-5.1)
    ora = O.VPSchedule(O.diffusion_buffers()["betas"][:100], clip=True)
    t = torch.tensor([1e-4, 0.01, 0.0100001, 0.5, 0.995, 1.0, 1.2])   # below first knot, on knots, above last knot
    assert torch.equal(ns.log_alpha_at(t), ora.log_mean(t))
    assert ns.n == 100
