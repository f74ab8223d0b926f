"""Units front-end on the B200 library: audio -> ``units[B, T, 1280]``, the step right BEFORE
``Unit2Mel.forward``.  Mirrors, name for name (file:line relative to the reference tree):

* ``log_mel_spectrogram`` / ``mel_filters``      encoder/whisper/audio.py:53-80
* ``ModelDimensions`` / ``sinusoids`` / ``AudioEncoder``   encoder/whisper/model.py:10-21,32-38,112-131
* ``WhisperLargeV3`` / ``Units_Encoder.encode``   tools/tools.py:43-126
* ``units_forced_alignment``                       tools/tools.py:193-223
* ``EuclideanCodebook`` (encode / decode)          quantize/kmeans_codebook.py:6-52

``AudioEncoder`` keeps the reference's constructor, ``state_dict()`` keys / shapes and default random init (so
``torch.manual_seed(s); AudioEncoder(...)`` holds bit-identical parameters and a reference checkpoint loads strictly); its
``forward`` runs on the CUDA library (csrc/units.cu on the sampler's tcgen05 GEMM / attention / LayerNorm kernels).  There
is no CPU fallback.  Resampling to 16 kHz (torchaudio ``Resample`` / librosa, tools/tools.py:78-95), the w2v-bert / xlsr
encoders stay with the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional, Union

import numpy as np
import torch
from torch import nn

from .capi import LdsError, load_library

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160


@dataclass
class ModelDimensions:          # encoder/whisper/model.py:10-21
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int = 0
    n_text_ctx: int = 0
    n_text_state: int = 0
    n_text_head: int = 0
    n_text_layer: int = 0


LARGE_V3 = ModelDimensions(n_mels=128, n_audio_ctx=1500, n_audio_state=1280, n_audio_head=20, n_audio_layer=32,
                           n_vocab=51866, n_text_ctx=448, n_text_state=1280, n_text_head=20, n_text_layer=32)


class UnitsConfig(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("n_state", C.c_int32), ("n_head", C.c_int32), ("n_layer", C.c_int32),
                ("precision", C.c_int32)]


def sinusoids(length: int, channels: int, max_timescale: int = 10000) -> torch.Tensor:
    """model.py:32-38, evaluated on the host exactly as the reference does before its ``.to("cuda")``."""
    assert channels % 2 == 0
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


def slaney_mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """The filterbank the reference ships as encoder/whisper/assets/mel_filters.npz (``librosa.filters.mel(sr=16000,
    n_fft=400, n_mels=n_mels)``: Slaney mel scale, area-normalised triangles), recomputed instead of shipped."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, np.newaxis]).astype(np.float32)


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int) -> torch.Tensor:
    """audio.py:53-58."""
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(slaney_mel_filterbank(n_mels)).to(device)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ucheck(lib, rc: int, what: str) -> None:
    if rc != 0:
        raise LdsError(f"{what} failed (status {rc}): {lib.lds_units_last_error().decode(errors='replace')}")


@torch.no_grad()
def log_mel_spectrogram(audio: Union[np.ndarray, torch.Tensor], n_mels: int = 128, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """audio.py:60-80: audio [L] or [B, L] (16 kHz) -> log-mel [n_mels, L // 160] / [B, n_mels, L // 160] on the CUDA device.
    (Decoding a file path with ffmpeg, audio.py:17-36, is not part of this package.)"""
    if isinstance(audio, str):
        raise NotImplementedError("load the waveform first; file decoding (ffmpeg) stays with the reference")
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(audio)
    if device is not None:
        audio = audio.to(device)
    if not audio.is_cuda:
        raise RuntimeError("log_mel_spectrogram runs on a CUDA device only (no CPU fallback): pass device='cuda'")
    if padding > 0:
        audio = torch.nn.functional.pad(audio, (0, padding))
    squeeze = audio.dim() == 1
    a = audio.reshape(-1, audio.shape[-1]).float().contiguous()
    B, L = a.shape
    lib = load_library()
    filt = mel_filters(audio.device, n_mels)
    out = torch.empty(B, n_mels, L // HOP_LENGTH, device=a.device, dtype=torch.float32)
    scratch = torch.empty(1, device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _ucheck(lib, lib.lds_units_log_mel(C.c_void_p(a.data_ptr()), B, L, C.c_void_p(filt.data_ptr()), n_mels,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()), _stream(a.device)),
                "lds_units_log_mel")
    return out[0] if squeeze else out.reshape(*audio.shape[:-1], n_mels, L // HOP_LENGTH)


@torch.no_grad()
def gather_rows(table: torch.Tensor, index: torch.Tensor, batched: bool) -> torch.Tensor:
    """batched: table [B, Tin, C], index [n] -> [B, n, C] (units_forced_alignment); else table [V, C], index [...] -> [..., C]
    (F.embedding)."""
    if not table.is_cuda:
        raise RuntimeError("gather_rows runs on a CUDA device only (no CPU fallback)")
    lib = load_library()
    t = table.float().contiguous()
    idx = index.to(device=t.device, dtype=torch.int64).contiguous()
    Cc = t.shape[-1]
    if batched:
        B, t_in = t.shape[0], t.shape[1]
        out = torch.empty(B, idx.numel(), Cc, device=t.device, dtype=torch.float32)
        args = (B, idx.numel(), t_in)
    else:
        out = torch.empty(*idx.shape, Cc, device=t.device, dtype=torch.float32)
        args = (1, idx.numel(), 0)
    with torch.cuda.device(t.device):
        _ucheck(lib, lib.lds_units_gather_rows(C.c_void_p(t.data_ptr()), C.c_void_p(idx.data_ptr()), *args, Cc,
                                               C.c_void_p(out.data_ptr()), _stream(t.device)), "lds_units_gather_rows")
    return out


def alignment_index(n_in: int, n_frames: int, scale_factor=None, units_forced_mode: str = "nearest") -> torch.Tensor:
    """Source frame of every output frame (host, int64 [n_frames]).
    'left' (tools/tools.py:204-207): clamp(round(scale_factor * arange(n_frames)), max = n_in - 1), the reference's expression.
    'nearest' (tools/tools.py:214-217 -> F.interpolate(mode='nearest', size=n_frames)): min(floor(i * (n_in / n_frames)), n_in - 1) with
    the ratio and the product in fp32, ATen's nearest_idx (UpSample.h); size == n_in and size == 2 n_in take its exact paths."""
    if units_forced_mode == "left":
        assert scale_factor is not None
        return torch.clamp(torch.round(scale_factor * torch.arange(n_frames)).long(), max=n_in - 1)
    if units_forced_mode in ("nearest", "rfa441to512", "rfa512to441"):
        i = torch.arange(n_frames)
        if n_frames == n_in:
            return i
        if n_frames == 2 * n_in:
            return i >> 1
        scale = torch.tensor(n_in, dtype=torch.float32) / torch.tensor(n_frames, dtype=torch.float32)
        return torch.clamp(torch.floor(i.to(torch.float32) * scale).long(), max=n_in - 1)
    raise NotImplementedError(f"units_forced_mode={units_forced_mode!r}: only the gathering modes ('nearest', 'left', 'rfa*') run on "
                              "the B200 path; interpolating modes stay with the reference")


def units_forced_alignment(units, audio=None, sample_rate=None, hop_size=None, n_frames=None, scale_factor=None,
                           units_forced_mode="nearest", device="cpu"):
    """tools/tools.py:193-223 for device-resident units [T, C] / [B, T, C]: a row gather along time on the GPU.  As in the
    reference, `size=n_frames` decides the output length (F.interpolate rejects size together with scale_factor, so the
    'nearest' modes are called with scale_factor=None by the reference's own callers, diffusion/data_loaders.py:200-203)."""
    assert (audio is not None and sample_rate is not None and hop_size is not None) or n_frames is not None or scale_factor is not None
    n_frames = int(audio.size(-1) // hop_size + 1) if n_frames is None else n_frames
    unit_is_tensor = torch.is_tensor(units)
    if not unit_is_tensor:
        units = torch.from_numpy(units)
    units_dim = units.dim()
    if units_dim == 2:
        units = units.unsqueeze(0)
    if units_forced_mode != "left" and scale_factor is not None:
        raise ValueError("only one of size or scale_factor should be defined")       # F.interpolate's own error
    idx = alignment_index(units.size(1), n_frames, scale_factor, units_forced_mode)
    units_aligned = gather_rows(units, idx, batched=True)
    if units_dim == 2:
        units_aligned = units_aligned.squeeze(0)
    return units_aligned if unit_is_tensor else units_aligned.cpu().numpy()


class UnitsEngine:
    """One ``lds_units`` handle (C ABI, include/lds_b200.h)."""

    def __init__(self, n_mels: int, n_state: int, n_head: int, n_layer: int, device: torch.device, precision: str = "fp32"):
        device = torch.device(device)
        if device.type != "cuda":
            raise LdsError("the units encoder runs on a CUDA device (B200, sm_100a) only; got %s" % device)
        self.lib = load_library()
        self.device, self.precision = device, precision
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.cfg = UnitsConfig(n_mels, n_state, n_head, n_layer, {"fp32": 0, "bf16": 1}[precision])
        self.handle = C.c_void_p()
        _ucheck(self.lib, self.lib.lds_units_create(C.byref(self.cfg), self.index, C.byref(self.handle)), "lds_units_create")
        self._pos = {}

    def load_state_dict(self, sd) -> None:
        for key, t in sd.items():
            t = t.detach()
            dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}.get(t.dtype)
            if dt is None:
                t, dt = t.float(), 0
            t = t.contiguous()
            shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
            _ucheck(self.lib, self.lib.lds_units_load_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt),
                    f"lds_units_load_weight({key})")
        _ucheck(self.lib, self.lib.lds_units_finalize(self.handle), "lds_units_finalize")

    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        mel = mel.to(device=self.device, dtype=torch.float32).contiguous()
        B, n_mels, L = mel.shape
        if n_mels != self.cfg.n_mels:
            raise ValueError(f"mel has {n_mels} bins, the encoder expects {self.cfg.n_mels}")
        T = int(self.lib.lds_units_out_frames(L))
        if T not in self._pos:
            self._pos = {T: sinusoids(T, self.cfg.n_state).to(self.device).contiguous()}
        out = torch.empty(B, T, self.cfg.n_state, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            _ucheck(self.lib, self.lib.lds_units_encode(self.handle, C.c_void_p(mel.data_ptr()), B, L, C.c_void_p(self._pos[T].data_ptr()),
                                                        C.c_void_p(out.data_ptr()), _stream(self.device)), "lds_units_encode")
        self._keep = mel
        return out

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.lds_units_launches(self.handle))

    @property
    def last_flops(self) -> float:
        return float(self.lib.lds_units_last_flops(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.lds_units_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _MultiHeadAttention(nn.Module):          # parameter container of model.py:43-50 (key has no bias)
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = nn.Linear(n_state, n_state)
        self.key = nn.Linear(n_state, n_state, bias=False)
        self.value = nn.Linear(n_state, n_state)
        self.out = nn.Linear(n_state, n_state)


class _ResidualAttentionBlock(nn.Module):      # model.py:89-102 (the encoder has no cross attention)
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.attn = _MultiHeadAttention(n_state, n_head)
        self.attn_ln = nn.LayerNorm(n_state)
        self.mlp = nn.Sequential(nn.Linear(n_state, n_state * 4), nn.GELU(), nn.Linear(n_state * 4, n_state))
        self.mlp_ln = nn.LayerNorm(n_state)


class AudioEncoder(nn.Module):
    """Parameter container + CUDA forward with the reference AudioEncoder's interface (model.py:112-131)."""

    def __init__(self, n_mels: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.conv1 = nn.Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.blocks = nn.ModuleList([_ResidualAttentionBlock(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = nn.LayerNorm(n_state)
        self.n_audio_state = n_state
        self._hp = dict(n_mels=n_mels, n_state=n_state, n_head=n_head, n_layer=n_layer)
        self.precision = "fp32"          # "fp32": split-f16 tcgen05 (fp32-accurate); "bf16": bf16 operands
        self._engine: Optional[UnitsEngine] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str) -> "AudioEncoder":
        if precision not in ("fp32", "bf16"):
            raise ValueError(precision)
        if precision != self.precision:
            self.precision = precision
            self._invalidate()
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, n_mels, L] -> [B, (L - 1) // 2 + 1, n_state] (model.py:120-131)."""
        if not x.is_cuda:
            raise RuntimeError("the units encoder runs on a CUDA device only (no CPU fallback): move the input to cuda")
        if self._engine is None or self._engine.device != x.device or self._engine.precision != self.precision:
            self._invalidate()
            self._engine = UnitsEngine(device=x.device, precision=self.precision, **self._hp)
            self._engine.load_state_dict(self.state_dict())
        return self._engine.encode(x)


class Whisper(nn.Module):
    """The audio half of model.py:133-139 (the reference only ever calls ``model.encoder``, tools/tools.py:124)."""

    def __init__(self, dims: ModelDimensions):
        super().__init__()
        self.dims = dims
        self.encoder = AudioEncoder(dims.n_mels, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer)


class WhisperLargeV3(torch.nn.Module):
    """tools/tools.py:112-126.  ``checkpoint``: path of 'pretrain/large-v3_encoder.pt' ({"dims", "model_state_dict"}), or None with
    ``dims`` (+ optional ``state_dict``) given directly (the reference tree ships no checkpoint)."""

    def __init__(self, device="cuda", checkpoint: Optional[str] = "pretrain/large-v3_encoder.pt", dims: Optional[ModelDimensions] = None,
                 state_dict=None):
        super().__init__()
        self.device = device
        if checkpoint is not None and dims is None:
            blob = torch.load(checkpoint, map_location="cpu")
            dims, state_dict = ModelDimensions(**blob["dims"]), blob["model_state_dict"]
        model = Whisper(dims)
        if state_dict is not None:
            enc = {k[len("encoder."):]: v for k, v in state_dict.items() if k.startswith("encoder.")}
            model.encoder.load_state_dict(enc if enc else state_dict)
        self.hidden_dim = dims
        self.model = model.to(device)
        self.model.eval()

    @torch.inference_mode()
    def __call__(self, audio, padding_mask=None):
        audio = audio.view(1, -1)
        mel = log_mel_spectrogram(audio, n_mels=self.hidden_dim.n_mels, device=self.device)
        if len(mel.shape) == 2:
            mel = mel.unsqueeze(0)
        units = self.model.encoder(mel).squeeze().data.cpu().float()
        return units


class Units_Encoder:
    """tools/tools.py:43-110 for encoder == 'whisper_large_v3' at the encoder's own sample rate."""

    def __init__(self, encoder, encoder_sample_rate=16000, encoder_hop_size=320, device=None, units_forced_mode="nearest", **whisper_kwargs):
        if device is None:
            device = "cuda"
        self.device = device
        self.encoder = encoder
        if units_forced_mode is None:
            units_forced_mode = "left"
        self.units_forced_mode = units_forced_mode
        if encoder != "whisper_large_v3":
            raise ValueError(f"[x] Unknown units encoder: {encoder}")       # w2v-bert / xlsr stay with the reference
        self.model = WhisperLargeV3(device=device, **whisper_kwargs)
        self.encoder_sample_rate = encoder_sample_rate
        self.encoder_hop_size = encoder_hop_size

    def encode(self, audio, sample_rate, padding_mask=None):
        if sample_rate != self.encoder_sample_rate:
            raise NotImplementedError("resampling (torchaudio Resample / librosa, tools/tools.py:78-95) stays with the reference: "
                                      f"pass {self.encoder_sample_rate} Hz audio")
        audio_res = audio
        if isinstance(audio_res, torch.Tensor) and audio_res.size(-1) < 400:
            audio_res = torch.nn.functional.pad(audio, (0, 400 - audio_res.size(-1)))
        units = self.model(audio_res, padding_mask=padding_mask)
        if units.shape[0] == 1:
            units = units.squeeze(0)
        return units


class EuclideanCodebook(nn.Module):
    """quantize/kmeans_codebook.py:6-46, decode direction (indices -> embeddings, F.embedding) on the GPU gather kernel."""

    def __init__(self, codebook_weight):
        super().__init__()
        self.register_buffer("embed", torch.as_tensor(codebook_weight).clone())

    def dequantize(self, embed_ind):
        return gather_rows(self.embed, embed_ind, batched=False)

    def decode(self, embed_ind):
        return self.dequantize(embed_ind)

    @torch.no_grad()
    def quantize(self, x):
        """x [M, C] -> nearest codeword index [M] (kmeans_codebook.py:15-23) on the GPU: fp32 FFMA GEMM x E^T - |e|^2/2, row argmax."""
        if not x.is_cuda:
            raise RuntimeError("EuclideanCodebook.quantize runs on a CUDA device only (no CPU fallback)")
        lib = load_library()
        xx = x.float().contiguous()
        emb = self.embed.to(device=xx.device, dtype=torch.float32).contiguous()
        M, Cc = xx.shape
        V = emb.shape[0]
        scratch = torch.empty((V + 63) // 64 * 64 + M * V, device=xx.device, dtype=torch.float32)
        idx = torch.empty(M, device=xx.device, dtype=torch.int64)
        with torch.cuda.device(xx.device):
            _ucheck(lib, lib.lds_units_quantize(C.c_void_p(xx.data_ptr()), C.c_void_p(emb.data_ptr()), M, V, Cc, C.c_void_p(scratch.data_ptr()),
                                                C.c_void_p(idx.data_ptr()), _stream(xx.device)), "lds_units_quantize")
        return idx

    def encode(self, x):
        """kmeans_codebook.py:37-46: flatten the leading dims, quantize, restore them."""
        shape = x.shape
        return self.quantize(x.reshape(-1, shape[-1])).view(*shape[:-1])

    def forward(self, x):
        return self.decode(self.encode(x))
