"""GPU: the units front-end (csrc/units.cu through the C ABI: log-mel, Whisper AudioEncoder on the tcgen05 kernels, alignment gather)
against the goldens of the executed reference and the fp64 oracle.

Tolerances (units are LayerNorm outputs, |u| ~ 3.4): fp32-accurate mode max-abs <= 2e-5 and relative L2 <= 4e-6 against fp64 (measured
2-6e-6 / 0.6-1.5e-6; the reference's own fp32 evaluation sits at 1.3e-6 / 2.6e-7 on the goldens); bf16 mode relative L2 <= 1e-2
(measured 2e-3).  Log-mel (values in [-1, 1.5]): max-abs <= 2e-5 against fp64 (measured 5-8e-6), <= 1e-4 against the reference's fp32
output (whose own distance to fp64 is 2-3.5e-5: fp32 FFT, where this library accumulates the DFT in fp64)."""
import numpy as np
import pytest
import torch

import gpu_util as G
from conftest import load_golden
from oracle import units_oracle as U

pytestmark = pytest.mark.gpu
NAMES = ["units_small_l4800", "units_small_l9733", "units_h8_l16000"]


def _encoder(seed, dims, precision="fp32"):
    from latent_diffusion_speech_b200.units import AudioEncoder
    torch.manual_seed(seed)
    enc = AudioEncoder(dims["n_mels"], dims["n_state"], dims["n_head"], dims["n_layer"]).eval()
    sd64 = {k: v.detach().double() for k, v in enc.state_dict().items()}
    return enc.cuda().set_precision(precision), sd64


def _dims(g):
    return dict(zip(("n_mels", "n_state", "n_head", "n_layer"), (int(v) for v in g["dims"])))


@pytest.mark.parametrize("name", NAMES)
def test_log_mel_vs_reference_golden_and_fp64(name):
    from latent_diffusion_speech_b200.units import log_mel_spectrogram, slaney_mel_filterbank
    g = load_golden(name)
    audio = torch.from_numpy(g["audio"])
    mel = log_mel_spectrogram(audio.view(1, -1), n_mels=128, device="cuda").cpu()
    want = torch.from_numpy(g["mel"])
    ref64 = U.log_mel_spectrogram(audio.double().view(1, -1), torch.from_numpy(slaney_mel_filterbank(128)).double())
    e, e64, floor = G.errs(mel, want), G.errs(mel, ref64), G.errs(want, ref64)
    G.report(test="units_log_mel", name=name, vs_ref=e, vs_fp64=e64, ref_fp32_vs_fp64=floor)
    assert mel.shape == want.shape
    assert e64["max_abs"] <= 2e-5, (e64, floor)
    assert e["max_abs"] <= 1e-4, (e, floor)
    assert mel.dim() == 3 and log_mel_spectrogram(audio, device="cuda").dim() == 2      # [L] -> [n_mels, frames] like the reference


def test_log_mel_30s_batch_vs_fp64_gpu_oracle():
    """Two 30 s rows (480 000 samples -> 3 000 frames each), global max over the batch as in the reference."""
    from latent_diffusion_speech_b200.units import log_mel_spectrogram, slaney_mel_filterbank
    audio = U.synthetic_audio(480000, seed=21, batch=2)
    audio[1] *= 0.05                                   # a quiet row: most of it sits on the (global max - 8) floor
    mel = log_mel_spectrogram(audio, device="cuda")
    ref64 = U.log_mel_spectrogram(audio.cuda().double(), torch.from_numpy(slaney_mel_filterbank(128)).double().cuda())
    e = G.errs(mel.cpu(), ref64.cpu())
    G.report(test="units_log_mel_30s_b2_vs_fp64_gpu_oracle", **e)
    assert mel.shape == (2, 128, 3000) and e["max_abs"] <= 2e-5, e


@pytest.mark.parametrize("name", NAMES)
def test_encoder_vs_reference_golden_and_fp64(name):
    g = load_golden(name)
    dims = _dims(g)
    enc, sd64 = _encoder(int(g["seed"]), dims)
    mel = torch.from_numpy(g["mel"])
    units = enc(mel.cuda()).cpu()
    want = torch.from_numpy(g["units"])
    with torch.no_grad():
        ref64 = U.audio_encoder(sd64, dims["n_head"], mel.double())
    e, e64, floor = G.errs(units, want), G.errs(units, ref64), G.errs(want, ref64)
    G.report(test="units_encoder_golden", name=name, vs_ref=e, vs_fp64=e64, ref_fp32_vs_fp64=floor)
    assert units.shape == want.shape
    assert e64["max_abs"] <= 2e-5 and e64["rel_l2"] <= 4e-6, (e64, floor)
    assert e["max_abs"] <= 2e-5, (e, floor)


@pytest.mark.parametrize("name", NAMES[:2])
def test_encoder_bf16_mode(name):
    g = load_golden(name)
    dims = _dims(g)
    enc, sd64 = _encoder(int(g["seed"]), dims, "bf16")
    mel = torch.from_numpy(g["mel"])
    units = enc(mel.cuda()).cpu()
    with torch.no_grad():
        ref64 = U.audio_encoder(sd64, dims["n_head"], mel.double())
    e = G.errs(units, ref64)
    G.report(test="units_encoder_bf16", name=name, **e)
    assert e["rel_l2"] <= 1e-2, e


@pytest.mark.parametrize("L,B", [(37, 1), (61, 3), (256, 2)])
def test_encoder_ragged_lengths_vs_fp64(L, B):
    enc, sd64 = _encoder(3, U.SMALL_DIMS)
    mel = U.synthetic_mel(B, L, seed=L)
    units = enc(mel.cuda()).cpu()
    with torch.no_grad():
        ref64 = U.audio_encoder(sd64, U.SMALL_DIMS["n_head"], mel.double())
    e = G.errs(units, ref64)
    G.report(test="units_encoder_ragged", L=L, B=B, **e)
    assert units.shape == (B, (L - 1) // 2 + 1, 256)
    assert e["max_abs"] <= 2e-5 and e["rel_l2"] <= 4e-6, e


def test_encoder_large_v3_width_30s_vs_fp64_gpu_oracle_and_batch_invariance():
    """whisper-large-v3 width (1280 channels, 20 heads of 64) at its full context (L = 3000 mel frames -> 1500 units), two layers,
    B = 2, against the oracle in fp64 on the GPU; utterance 1 alone is bit-identical.  The K = 1280 ... 5120 contractions of this width
    carry the tensor core's truncating accumulation over 80 ... 320 steps (measured 2.3e-5 / 4.2e-6): max-abs <= 5e-5, rel-L2 <= 1e-5."""
    dims = dict(n_mels=128, n_state=1280, n_head=20, n_layer=2)
    enc, sd64 = _encoder(1234, dims)
    mel = U.synthetic_mel(2, 3000, seed=8)
    units = enc(mel.cuda())
    with torch.no_grad():
        ref64 = U.audio_encoder({k: v.cuda() for k, v in sd64.items()}, 20, mel.cuda().double())
    e = G.errs(units.cpu(), ref64.cpu())
    G.report(test="units_encoder_w1280_l3000_b2_vs_fp64_gpu_oracle", **e)
    assert units.shape == (2, 1500, 1280) and torch.isfinite(units).all()
    assert e["max_abs"] <= 5e-5 and e["rel_l2"] <= 1e-5, e
    alone = enc(mel[1:2].cuda())
    assert torch.equal(alone, units[1:2])
    enc.set_precision("bf16")
    e16 = G.errs(enc(mel.cuda()).cpu(), ref64.cpu())
    G.report(test="units_encoder_w1280_l3000_b2_bf16", **e16)
    assert e16["rel_l2"] <= 1e-2, e16


@pytest.mark.parametrize("name", NAMES)
def test_units_forced_alignment_equals_reference(name):
    from latent_diffusion_speech_b200.units import units_forced_alignment
    g = load_golden(name)
    scale = float(g["align_scale"])
    units = torch.from_numpy(g["units"])
    out = units_forced_alignment(units.cuda(), n_frames=int(g["align_frames"]), scale_factor=None if scale < 0 else scale,
                                 units_forced_mode=str(g["align_mode"]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["aligned"]))
    out2 = units_forced_alignment(units[0].cuda(), n_frames=int(g["align_frames"]), scale_factor=None if scale < 0 else scale,
                                  units_forced_mode=str(g["align_mode"]))
    assert out2.dim() == 2 and torch.equal(out2.cpu(), torch.from_numpy(g["aligned"])[0])


def test_alignment_to_mel_rate_and_codebook_decode():
    """1500 units of 30 s -> 2584 mel frames (44.1 kHz / 512), B = 3; and F.embedding through the same gather kernel."""
    from latent_diffusion_speech_b200.units import EuclideanCodebook, units_forced_alignment
    gen = torch.Generator().manual_seed(4)
    units = torch.randn(3, 1500, 1280, generator=gen)
    out = units_forced_alignment(units.cuda(), n_frames=2584)
    assert torch.equal(out.cpu(), U.units_forced_alignment(units, 2584))
    book = torch.randn(2048, 1280, generator=gen)
    ind = torch.randint(0, 2048, (2, 77), generator=gen)
    dec = EuclideanCodebook(book.numpy()).cuda().decode(ind.cuda())
    assert torch.equal(dec.cpu(), U.codebook_decode(book, ind))


def test_audio_to_units_wrapper_vs_golden():
    """WhisperLargeV3.__call__ (tools/tools.py:112-126): audio -> log-mel -> encoder -> CPU float units, on the golden's weights."""
    from latent_diffusion_speech_b200.units import ModelDimensions, Units_Encoder
    g = load_golden("units_small_l9733")
    d = _dims(g)
    torch.manual_seed(int(g["seed"]))
    ue = Units_Encoder("whisper_large_v3", 16000, 320, device="cuda", checkpoint=None,
                       dims=ModelDimensions(n_mels=d["n_mels"], n_audio_ctx=1500, n_audio_state=d["n_state"], n_audio_head=d["n_head"],
                                            n_audio_layer=d["n_layer"]))
    units = ue.encode(torch.from_numpy(g["audio"]).cuda(), 16000)
    want = torch.from_numpy(g["units"])[0]
    e = G.errs(units, want)
    G.report(test="units_audio_to_units_wrapper", **e)
    assert units.shape == want.shape and not units.is_cuda
    assert e["max_abs"] <= 5e-4, e
    with pytest.raises(NotImplementedError):
        ue.encode(torch.zeros(22050).cuda(), 22050)


def test_codebook_quantize_and_roundtrip():
    """EuclideanCodebook.encode / quantize (quantize/kmeans_codebook.py:15-23,37-46): nearest codeword on the GPU (fp32 FFMA GEMM + row
    argmax).  Perturbed codewords must come back as themselves; on unstructured inputs the chosen codeword is the fp64 optimum up to
    near-ties (its distance within 1e-5 relative of the best)."""
    from latent_diffusion_speech_b200.units import EuclideanCodebook
    gen = torch.Generator().manual_seed(12)
    V, Cc = 2048, 1280
    book = torch.randn(V, Cc, generator=gen)
    cb = EuclideanCodebook(book.numpy()).cuda()
    true = torch.randint(0, V, (3, 257), generator=gen)
    x = book[true] + 0.05 * torch.randn(3, 257, Cc, generator=gen)
    got = cb.encode(x.cuda()).cpu()
    assert got.shape == true.shape and torch.equal(got, true)
    assert torch.equal(cb(x.cuda()).cpu(), book[true])                    # forward = decode(encode(x))
    y = torch.randn(500, Cc, generator=gen)
    idx = cb.quantize(y.cuda()).cpu()
    d = torch.cdist(y.double(), book.double()) ** 2
    best, chosen = d.min(dim=1).values, d[torch.arange(500), idx]
    agree = float((idx == d.argmin(dim=1)).float().mean())
    G.report(test="units_codebook_quantize", agree=agree, worst_rel_excess=float(((chosen - best) / best).max()))
    assert agree >= 0.99 and float(((chosen - best) / best).max()) <= 1e-5
