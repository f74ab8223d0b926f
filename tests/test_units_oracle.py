"""CPU: the units front-end oracle (oracle/units_oracle.py) and the host-side mirror (latent_diffusion_speech_b200/units.py) against
the executed reference — goldens from oracle/make_golden_units.py everywhere, live against /root/reference in the authoring
container (`needs_reference`)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, state_dict_checksum
from oracle import units_oracle as U

NAMES = ["units_small_l4800", "units_small_l9733", "units_h8_l16000"]
needs_reference = pytest.mark.skipif(not U.reference_available(), reason="reference tree not present")


def _host_encoder(g):
    from latent_diffusion_speech_b200.units import AudioEncoder
    n_mels, n_state, n_head, n_layer = (int(v) for v in g["dims"])
    torch.manual_seed(int(g["seed"]))
    return AudioEncoder(n_mels, n_state, n_head, n_layer).eval(), n_head


@pytest.mark.parametrize("name", NAMES)
def test_host_encoder_reproduces_reference_init_and_oracle_matches_golden(name):
    from latent_diffusion_speech_b200.units import slaney_mel_filterbank
    g = load_golden(name)
    enc, n_head = _host_encoder(g)
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    assert state_dict_checksum(sd) == str(g["weights_sha256"]), "default init differs from the reference AudioEncoder's"
    audio = U.synthetic_audio(int(g["n_samples"]), seed=int(g["seed"]))
    assert torch.equal(audio, torch.from_numpy(g["audio"]))
    with torch.no_grad():
        mel = U.log_mel_spectrogram(audio.view(1, -1), torch.from_numpy(slaney_mel_filterbank(128)))
        # the golden mel used the reference's shipped filterbank asset, ours is recomputed: equal to fp32 round-off
        assert float((mel - torch.from_numpy(g["mel"])).abs().max()) <= 2e-6
        units = U.audio_encoder(sd, n_head, torch.from_numpy(g["mel"]))
    assert float((units - torch.from_numpy(g["units"])).abs().max()) <= 2e-5        # same ops; thread-count dependent summation order only
    scale = float(g["align_scale"])
    aligned = U.units_forced_alignment(torch.from_numpy(g["units"]), int(g["align_frames"]), None if scale < 0 else scale, str(g["align_mode"]))
    assert torch.equal(aligned, torch.from_numpy(g["aligned"]))


def test_sinusoids_equal_oracle():
    from latent_diffusion_speech_b200.units import sinusoids
    for n, c in ((15, 256), (1500, 1280), (7, 384)):
        assert torch.equal(sinusoids(n, c), U.sinusoids(n, c))


@pytest.mark.parametrize("n_in,n_frames", [(15, 26), (30, 53), (50, 87), (1500, 2584), (1500, 2585), (750, 1500), (100, 100), (431, 864),
                                           (37, 11), (3, 1000), (1, 5), (999, 1000)])
def test_alignment_index_is_aten_nearest(n_in, n_frames):
    """units_forced_alignment 'nearest' = F.interpolate(size=n_frames): the host-computed gather index must pick the same rows."""
    from latent_diffusion_speech_b200.units import alignment_index
    u = torch.arange(n_in, dtype=torch.float32).view(1, n_in, 1).repeat(1, 1, 2)
    ref = U.units_forced_alignment(u, n_frames, None, "nearest")[0, :, 0].long()
    assert torch.equal(alignment_index(n_in, n_frames, None, "nearest"), ref)


@pytest.mark.parametrize("n_in,n_frames,scale", [(30, 53, 0.5742), (1500, 2584, 0.58049886), (10, 40, 0.25), (10, 40, 0.3)])
def test_alignment_index_left(n_in, n_frames, scale):
    from latent_diffusion_speech_b200.units import alignment_index
    u = torch.arange(n_in, dtype=torch.float32).view(1, n_in, 1).repeat(1, 1, 2)
    ref = U.units_forced_alignment(u, n_frames, scale, "left")[0, :, 0].long()
    assert torch.equal(alignment_index(n_in, n_frames, scale, "left"), ref)


def test_no_cpu_fallback():
    from latent_diffusion_speech_b200.units import AudioEncoder, log_mel_spectrogram, units_forced_alignment, EuclideanCodebook
    enc = AudioEncoder(128, 256, 4, 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(1, 128, 30))
    with pytest.raises(RuntimeError, match="CUDA"):
        log_mel_spectrogram(torch.zeros(4800))
    with pytest.raises(RuntimeError, match="CUDA"):
        units_forced_alignment(torch.zeros(1, 15, 256), n_frames=26)
    with pytest.raises(RuntimeError, match="CUDA"):
        EuclideanCodebook(np.zeros((8, 4), np.float32)).decode(torch.zeros(3, dtype=torch.long))


# ---- live against the unmodified reference (authoring container) ----------------------------------------------------------
@needs_reference
def test_filterbank_equals_reference_asset():
    import os
    from latent_diffusion_speech_b200.units import slaney_mel_filterbank
    with np.load(os.path.join(U.REFERENCE_ROOT, "encoder", "whisper", "assets", "mel_filters.npz")) as f:
        for n_mels in (80, 128):
            ref = f[f"mel_{n_mels}"]
            ours = slaney_mel_filterbank(n_mels)
            assert ours.shape == ref.shape and ours.dtype == ref.dtype
            assert float(np.abs(ours - ref).max()) <= 1e-8 and np.array_equal(ours != 0, ref != 0)


@needs_reference
@pytest.mark.parametrize("n_samples,batch", [(4800, 0), (9733, 0), (16000, 2)])
def test_oracle_log_mel_equals_live_reference(n_samples, batch):
    _, audio_mod = U.import_reference_whisper()
    audio = U.synthetic_audio(n_samples, seed=3, batch=batch)
    a = audio if batch else audio.view(1, -1)
    with torch.no_grad():
        ref = audio_mod.log_mel_spectrogram(a, n_mels=128)
        ours = U.log_mel_spectrogram(a, audio_mod.mel_filters(a.device, 128))
    assert torch.equal(ours, ref)


@needs_reference
@pytest.mark.parametrize("dims,L,B", [(U.SMALL_DIMS, 30, 1), (U.SMALL_DIMS, 61, 2), (dict(n_mels=128, n_state=384, n_head=6, n_layer=1), 100, 1)])
def test_oracle_encoder_equals_live_reference(dims, L, B):
    from latent_diffusion_speech_b200.units import AudioEncoder
    model_mod, _ = U.import_reference_whisper()
    torch.manual_seed(9)
    ref_enc = model_mod.AudioEncoder(dims["n_mels"], dims["n_state"], dims["n_head"], dims["n_layer"]).eval()
    torch.manual_seed(9)
    host = AudioEncoder(dims["n_mels"], dims["n_state"], dims["n_head"], dims["n_layer"]).eval()
    ref_sd, host_sd = ref_enc.state_dict(), host.state_dict()
    assert list(ref_sd.keys()) == list(host_sd.keys())
    assert all(torch.equal(ref_sd[k], host_sd[k]) for k in ref_sd)
    host.load_state_dict(ref_sd, strict=True)                    # a reference checkpoint loads strictly
    mel = U.synthetic_mel(B, L)
    with torch.no_grad(), U.cuda_moves_ignored():
        ref = ref_enc(mel)
        ours = U.audio_encoder({k: v.detach() for k, v in ref_sd.items()}, dims["n_head"], mel)
    assert torch.equal(ours, ref)


@needs_reference
@pytest.mark.parametrize("mode,n_frames,scale", [("nearest", 26, None), ("nearest", 2584, None), ("left", 53, 0.5742), ("rfa441to512", 40, None)])
def test_oracle_alignment_equals_live_reference(mode, n_frames, scale):
    align = U.reference_function("tools/tools.py", "units_forced_alignment", {"torch": torch, "np": np})
    units = torch.randn(1, 1500 if n_frames > 1000 else 15, 16, generator=torch.Generator().manual_seed(1))
    ref = align(units, n_frames=n_frames, scale_factor=scale, units_forced_mode=mode)
    assert torch.equal(U.units_forced_alignment(units, n_frames, scale, mode), ref)
