"""TEST INFRASTRUCTURE ONLY — CPU restatement of the units front-end in plain PyTorch.

Nothing under ``latent_diffusion_speech_b200/`` imports this module; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s baseline legs do, as the checker.  Every function cites the reference code it follows (file:line relative to
the reference tree) and is pinned to the executed reference by ``tests/test_units_oracle.py`` (live, in the authoring
container) and by the goldens of ``oracle/make_golden_units.py`` under ``tests/golden/``.

All functions are dtype-generic: called with float64 inputs / weights they give the fp64 reference the GPU tests compare to.
"""
import ast
import contextlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

N_FFT, HOP_LENGTH = 400, 160                       # encoder/whisper/audio.py:10-11
SMALL_DIMS = dict(n_mels=128, n_state=256, n_head=4, n_layer=2)      # head dim 64 like every Whisper size


def log_mel_spectrogram(audio: torch.Tensor, filters: torch.Tensor) -> torch.Tensor:
    """encoder/whisper/audio.py:60-80 with the filterbank passed in (audio.py:53-58 loads it from an asset)."""
    window = torch.hann_window(N_FFT, dtype=audio.dtype).to(audio.device)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = filters.to(audio.dtype) @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec


def sinusoids(length, channels, max_timescale=10000):
    """encoder/whisper/model.py:32-38 without the hard-coded ``.to(device="cuda")``."""
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


def _attention(sd, p, n_head, x):
    """MultiHeadAttention.forward / qkv_attention (model.py:52-87), self-attention, no mask, no cache."""
    q = F.linear(x, sd[p + "query.weight"], sd[p + "query.bias"])
    k = F.linear(x, sd[p + "key.weight"])
    v = F.linear(x, sd[p + "value.weight"], sd[p + "value.bias"])
    n_batch, n_ctx, n_state = q.shape
    scale = (n_state // n_head) ** -0.25
    q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3) * scale
    k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 3, 1) * scale
    v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    qk = q @ k
    w = F.softmax(qk, dim=-1).to(q.dtype)
    wv = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
    return F.linear(wv, sd[p + "out.weight"], sd[p + "out.bias"])


def audio_encoder(sd, n_head: int, x: torch.Tensor) -> torch.Tensor:
    """AudioEncoder.forward (model.py:119-131) + ResidualAttentionBlock.forward (:104-110); sd = AudioEncoder.state_dict().
    x [B, n_mels, L] -> [B, (L - 1) // 2 + 1, n_state]."""
    n_state = sd["ln_post.weight"].shape[0]
    x = F.gelu(F.conv1d(x, sd["conv1.weight"], sd["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd["conv2.weight"], sd["conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1)
    x = (x + sinusoids(x.size(1), n_state).to(device=x.device, dtype=x.dtype)).to(x.dtype)
    n_layer = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    for i in range(n_layer):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (n_state,), sd[p + "attn_ln.weight"], sd[p + "attn_ln.bias"])
        x = x + _attention(sd, p + "attn.", n_head, h)
        h = F.layer_norm(x, (n_state,), sd[p + "mlp_ln.weight"], sd[p + "mlp_ln.bias"])
        h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])), sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
        x = x + h
    return F.layer_norm(x, (n_state,), sd["ln_post.weight"], sd["ln_post.bias"])


def units_forced_alignment(units: torch.Tensor, n_frames: int, scale_factor=None, units_forced_mode="nearest") -> torch.Tensor:
    """tools/tools.py:193-223 for tensor input [B, T, C] (the reference's own callers pass n_frames for 'nearest',
    diffusion/data_loaders.py:200-203, and scale_factor for 'left')."""
    if units_forced_mode == "left":
        index = torch.clamp(torch.round(scale_factor * torch.arange(n_frames)).long(), max=units.size(1) - 1)
        return torch.gather(units, 1, index.unsqueeze(0).unsqueeze(-1).repeat([units.size(0), 1, units.size(-1)]))
    u = units.transpose(1, 2)
    return F.interpolate(u, size=n_frames, scale_factor=scale_factor, mode="nearest").transpose(-1, -2)


def codebook_decode(embed: torch.Tensor, embed_ind: torch.Tensor) -> torch.Tensor:
    """EuclideanCodebook.dequantize (quantize/kmeans_codebook.py:29-31)."""
    return F.embedding(embed_ind, embed)


def synthetic_audio(n_samples: int, seed: int = 11, batch: int = 0) -> torch.Tensor:
    """Speech-like test signal in [-1, 1]: a few decaying harmonics with vibrato, an amplitude envelope with a silent stretch
    (so that the (max - 8) floor of the log-mel is active) and a little noise."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(max(1, batch)):
        t = torch.arange(n_samples, dtype=torch.float64) / 16000.0
        f0 = 110.0 + 40.0 * float(torch.rand(1, generator=g)) + 8.0 * torch.sin(2 * np.pi * 5.0 * t)
        phase = 2 * np.pi * torch.cumsum(f0, 0) / 16000.0
        x = sum((0.6 ** h) * torch.sin((h + 1) * phase) for h in range(8))
        env = 0.5 * (1 + torch.sin(2 * np.pi * 1.3 * t + float(torch.rand(1, generator=g)) * 6.28)) ** 2
        env[n_samples // 3: n_samples // 3 + n_samples // 8] = 0.0
        x = 0.3 * x * env + 1e-3 * torch.randn(n_samples, generator=g, dtype=torch.float64)
        rows.append(x.clamp(-1, 1).float())
    return torch.stack(rows) if batch else rows[0]


def synthetic_mel(B: int, L: int, n_mels: int = 128, seed: int = 5) -> torch.Tensor:
    """Log-mel-like input for encoder-only tests: values in about [-1, 1.5] with structure along time."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, n_mels, (L + 3) // 4 + 1, generator=g)
    x = F.interpolate(base, size=L, mode="linear", align_corners=True) * 0.5 + 0.1 * torch.randn(B, n_mels, L, generator=g)
    return x.clamp(-1.0, 1.5)


# ---- the reference itself (authoring container only) ---------------------------------------------------------------------
REFERENCE_ROOT = os.environ.get("LDS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "encoder", "whisper", "model.py"))


@contextlib.contextmanager
def cuda_moves_ignored():
    """The reference's ``sinusoids`` ends in ``.to(device="cuda")`` (model.py:38), which cannot run in the CPU-only authoring
    container.  While this context is active ``Tensor.to(device="cuda")`` returns the tensor unchanged, so the UNMODIFIED
    reference forward runs on the CPU; nothing else of torch is touched."""
    orig = torch.Tensor.to

    def to(self, *args, **kwargs):
        if kwargs.get("device", None) == "cuda" and not torch.cuda.is_available():
            kwargs = {k: v for k, v in kwargs.items() if k != "device"}
            if not args and not kwargs:
                return self
        return orig(self, *args, **kwargs)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def import_reference_whisper():
    """(encoder.whisper.model, encoder.whisper.audio) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("encoder.whisper.model"), importlib.import_module("encoder.whisper.audio")


def reference_function(rel_path: str, name: str, namespace: dict):
    """Compiles ONE top-level function of a reference file, verbatim, into `namespace` — for modules whose other imports
    (librosa, fairseq: tools/tools.py:1-11) are absent here."""
    src = open(os.path.join(REFERENCE_ROOT, rel_path)).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REFERENCE_ROOT, rel_path), "exec")
    exec(code, namespace)
    return namespace[name]
