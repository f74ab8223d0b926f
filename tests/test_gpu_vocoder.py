"""GPU: the CUDA vocoder (csrc/vocoder.cu through the C ABI) against the goldens of the executed reference Generator and the
fp64 oracle.  fp32 FFMA vs the reference's fp32 convolutions differ only in summation order: |wav| <= 1 (tanh), tolerance
max-abs 2e-6 and relative L2 1e-5 against fp64."""
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
