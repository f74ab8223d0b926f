"""GPU: the CUDA vocoder (csrc/vocoder.cu through the C ABI) against the goldens of the executed reference Generator and the
fp64 oracle.  The 256- / 128-channel levels run their dilated convolutions as split-f16 tensor-core GEMMs (fp32-accurate: 22-bit
operands, fp32 accumulation in TMEM), the narrow levels as fp32 FFMA: |wav| <= 1 (tanh), tolerance max-abs 2e-6 and relative L2 1e-5
against fp64 — the tolerance of the all-FFMA form; measured 3-5e-8 / 2.9e-7 on |wav| ~ 0.04 (the reference's own fp32: 2e-8 / 1.7e-7)."""
import pytest
import torch

import gpu_util as G
from conftest import load_golden
from oracle import vocoder_oracle as V

pytestmark = pytest.mark.gpu


def _gen(seed, kind):
    from latent_diffusion_speech_b200.vocoder import Generator
    h = dict(V.DEFAULT_H, resblock=kind)
    torch.manual_seed(seed)
    return Generator(h).eval(), h


@pytest.mark.parametrize("name", ["vocoder_rb1_b2_t9", "vocoder_rb1_b1_t40", "vocoder_rb2_b2_t13"])
def test_vocoder_vs_reference_golden(name):
    g = load_golden(name)
    gen, h = _gen(int(g["seed"]), str(g["resblock"]))
    sd64 = {k: v.detach().double() for k, v in gen.state_dict().items()}
    z = V.synthetic_latents(int(g["B"]), int(g["T"]), h["inter_channels"])
    wav = gen.cuda().decode_frames(z.cuda()).cpu()
    want = torch.from_numpy(g["wav"])
    with torch.no_grad():
        ref64 = V.vocoder_infer(sd64, h, z.double())
    e, e64, floor = G.errs(wav, want), G.errs(wav, ref64), G.errs(want, ref64)
    G.report(test="vocoder_golden", name=name, vs_ref=e, vs_fp64=e64, ref_fp32_vs_fp64=floor)
    assert wav.shape == want.shape
    assert e64["max_abs"] <= 2e-6 and e64["rel_l2"] <= 1e-5, (e64, floor)
    assert e["max_abs"] <= 2e-6, (e, floor)


def test_vocoder_10s_utterance_vs_fp64_gpu_oracle_and_batch_invariance():
    """T=864 frames (10 s -> 442 368 samples), B=2, against the oracle in fp64 on the GPU; utterance 1 alone is bit-identical."""
    gen, h = _gen(1234, "1")
    z = V.synthetic_latents(2, 864, h["inter_channels"], seed=33)
    sd64 = {k: v.detach().double().cuda() for k, v in gen.state_dict().items()}
    gen = gen.cuda()
    wav = gen.decode_frames(z.cuda())
    with torch.no_grad():
        ref64 = V.vocoder_infer(sd64, h, z.cuda().double())
    e = G.errs(wav.cpu(), ref64.cpu())
    G.report(test="vocoder_b2_t864_vs_fp64_gpu_oracle", **e)
    assert wav.shape == (2, 1, 864 * 512) and torch.isfinite(wav).all()
    assert e["max_abs"] <= 2e-6 and e["rel_l2"] <= 1e-5, e
    alone = gen.decode_frames(z[1:2].cuda())
    assert torch.equal(alone, wav[1:2])


def test_vocoder_interface_mirrors_reference_wrapper():
    from latent_diffusion_speech_b200.vocoder import Vocoder
    torch.manual_seed(1)
    voc = Vocoder("hifi-vaegan", None, device="cuda")
    assert (voc.dimension, voc.vocoder_hop_size, voc.vocoder_sample_rate) == (128, 512, 44100)
    wav = voc.infer(V.synthetic_latents(1, 16).cuda())
    assert wav.shape == (1, 1, 16 * 512) and float(wav.abs().max()) <= 1.0
    with pytest.raises(RuntimeError):
        voc.infer(V.synthetic_latents(1, 4))           # CPU tensor: no fallback
    with pytest.raises(ValueError):
        Vocoder("nsf-hifigan", None, device="cuda")


@pytest.mark.parametrize("taps,dil,cin,N,rows,batches,epi", [(3, 1, 128, 128, 200, 2, 0), (7, 3, 256, 256, 333, 2, 4), (11, 5, 128, 128, 160, 3, 0),
                                                          (11, 1, 256, 256, 97, 1, 4), (5, 2, 128, 256, 128, 2, 0), (7, 3, 64, 64, 301, 2, 0), (11, 5, 64, 64, 256, 2, 4),
                                                          (3, 1, 64, 320, 100, 1, 0)])
def test_dilated_conv1d_tc_vs_fp64(taps, dil, cin, N, rows, batches, epi):
    """lds_op_conv1d_tc (the generator's dilated ResBlock convolutions as implicit GEMMs, models.py:166-184) against F.conv1d in fp64:
    'same' padding dil*(k-1)/2, bias, leaky_relu(0.1) epilogue or fp32 residual; ragged rows, several utterances."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(taps * 100 + dil)
    x = torch.randn(batches, rows, cin, generator=g)
    w = torch.randn(N, cin, taps, generator=g) * (cin * taps) ** -0.5
    bias = torch.randn(N, generator=g)
    R = torch.randn(batches * rows, N, generator=g) if epi == 0 else None
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=dil * (taps - 1) // 2, dilation=dil).transpose(1, 2).reshape(-1, N)
    ref = F.leaky_relu(ref, 0.1) if epi == 4 else ref + R.double()
    a = G.op_split_cast(x.reshape(-1, cin).cuda(), 2)
    wp = G.pack_w_parts(w.permute(0, 2, 1).reshape(N, taps * cin).contiguous().cuda(), taps, 2)
    out = G.op_conv1d_tc(a, batches, rows, cin, 2, wp, N, taps, dil, bias=bias.cuda(), R=None if R is None else R.cuda(), epilogue=epi, act_slope=0.1)
    e = G.errs(out.cpu(), ref)
    G.report(test="dilated_conv1d_tc", taps=taps, dil=dil, cin=cin, N=N, **e)
    assert e["rel_l2"] <= 5e-6 and e["max_abs"] <= 4e-5, e      # measured 0.3-2.5e-6 (K = 384 ... 2816: truncating fp32 accumulate in TMEM)


@pytest.mark.parametrize("c0,rates,ksizes", [(256, [4, 3], [8, 7]),        # GEMM conv_pre + up 0, 128-ch level on gemm_tc, FFMA up 1 (odd stride), 64-ch level on gemm_tc (N = 64)
                                            (192, [4, 3], [8, 7]),        # GEMM conv_pre + up 0, 96-ch and 48-ch levels on the CUDA cores
                                            (128, [2, 2, 2], [4, 4, 4])]) # 64-, 32- (time-folded) and 16-channel (fold 4) levels, every layer but conv_post a GEMM
def test_vocoder_mixed_layer_forms_vs_fp64(c0, rates, ksizes):
    """Generator layouts other than HiFi-GAN V1 exercise every transition between the tensor-core (channels-last) and CUDA-core
    (channels-first) forms of vocoder.cu, the FFMA fallback of the transposed convolution, and time folding by 2 and by 4."""
    from latent_diffusion_speech_b200.vocoder import Generator
    h = dict(V.DEFAULT_H, upsample_initial_channel=c0, upsample_rates=rates, upsample_kernel_sizes=ksizes,
             resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]])
    torch.manual_seed(99)
    gen = Generator(h).eval()
    sd64 = {k: v.detach().double() for k, v in gen.state_dict().items()}
    z = V.synthetic_latents(2, 21, h["inter_channels"], seed=3)
    wav = gen.cuda().decode_frames(z.cuda()).cpu()
    with torch.no_grad():
        ref64 = V.vocoder_infer(sd64, h, z.double())
    e = G.errs(wav, ref64)
    G.report(test="vocoder_mixed_forms", c0=c0, rates=str(rates), **e)
    hop = 1
    for u in rates:
        hop *= u
    assert wav.shape == (2, 1, 21 * hop)
    assert e["max_abs"] <= 2e-6 and e["rel_l2"] <= 1e-5, e
