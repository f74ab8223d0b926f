"""ctypes binding of liblds_b200.so (include/lds_b200.h) — the drop-in boundary.

There is deliberately no fallback: if the shared library is missing, cannot be loaded, or no
sm_100 device is present, the product path raises.  (The only CPU implementation in this repo
is the test oracle under ``oracle/``, which this package never imports.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblds_b200.so")
MAX_BLOCKS = 8
COEF_STRIDE = 12
PREC_FP32, PREC_BF16, PREC_FP32_FFMA = 0, 1, 2
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2

EXPORTS = [
    "lds_version", "lds_last_error", "lds_create", "lds_destroy", "lds_load_weight", "lds_finalize_weights",
    "lds_plan", "lds_cond", "lds_denoise", "lds_sample_begin", "lds_sample_begin_shallow", "lds_sample_steps", "lds_sample_end",
    "lds_sample", "lds_train_loss",
    "lds_num_steps", "lds_workspace_bytes", "lds_kernel_launches", "lds_set_profiling", "lds_profile_num_classes",
    "lds_profile_class_name", "lds_profile_class_ms", "lds_profile_class_launches", "lds_profile_class_flops",
    "lds_profile_class_bytes", "lds_op_gemm", "lds_op_attention", "lds_op_groupnorm", "lds_op_groupnorm_fused", "lds_op_groupnorm_cluster", "lds_op_layernorm",
    "lds_op_split_cast", "lds_op_gemm_tc", "lds_op_conv1d_tc", "lds_op_qkv_attention_tc",
    "lds_op_x0_pred", "lds_op_dpm_update", "lds_op_unipc_predict", "lds_op_unipc_correct", "lds_op_ddpm_step", "lds_op_ddim_step",
    "lds_op_pndm_update", "lds_op_q_sample", "lds_op_cast_gather", "lds_op_transpose", "lds_op_div_copy",
    "lds_vocoder_last_error", "lds_vocoder_create", "lds_vocoder_destroy", "lds_vocoder_load_weight", "lds_vocoder_finalize",
    "lds_vocode", "lds_vocoder_hop", "lds_vocoder_launches", "lds_vocoder_last_flops", "lds_vocoder_workspace_bytes",
    "lds_units_last_error", "lds_units_create", "lds_units_destroy", "lds_units_load_weight", "lds_units_finalize", "lds_units_encode",
    "lds_units_out_frames", "lds_units_launches", "lds_units_last_flops", "lds_units_workspace_bytes", "lds_units_log_mel",
    "lds_units_gather_rows", "lds_units_quantize",
]


class LdsConfig(C.Structure):
    _fields_ = [
        ("input_channel", C.c_int32), ("n_spk", C.c_int32), ("out_dims", C.c_int32), ("n_layers", C.c_int32),
        ("n_blocks", C.c_int32), ("block_out_channels", C.c_int32 * MAX_BLOCKS), ("n_heads", C.c_int32),
        ("n_hidden", C.c_int32), ("norm_groups", C.c_int32), ("acoustic_scale", C.c_float), ("precision", C.c_int32),
    ]


class LdsError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads liblds_b200.so and declares every prototype of include/lds_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("LDS_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise LdsError(
            f"{p} not found: build it with `python -m latent_diffusion_speech_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the diffusion sampling path.")
    lib = C.CDLL(p)
    vp, i32, i64p, fp = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_float)
    f32, i64 = C.c_float, C.c_int64
    sig = {
        "lds_version": (i32, []),
        "lds_last_error": (C.c_char_p, []),
        "lds_create": (i32, [C.POINTER(LdsConfig), i32, C.POINTER(vp)]),
        "lds_destroy": (None, [vp]),
        "lds_load_weight": (i32, [vp, C.c_char_p, vp, i64p, i32, i32]),
        "lds_finalize_weights": (i32, [vp]),
        "lds_plan": (i32, [vp, i32, i32, i32, i32, fp, i32, fp]),
        "lds_cond": (i32, [vp, vp, vp, vp, vp]),
        "lds_denoise": (i32, [vp, vp, vp, fp, vp, vp]),
        "lds_sample_begin": (i32, [vp, vp, vp, vp]),
        "lds_sample_begin_shallow": (i32, [vp, vp, vp, vp, C.c_float, C.c_float, vp]),
        "lds_sample_steps": (i32, [vp, i32, i32, vp, vp]),
        "lds_sample_end": (i32, [vp, vp, vp]),
        "lds_sample": (i32, [vp, vp, vp, vp, vp, vp]),
        "lds_train_loss": (i32, [vp, vp, vp, vp, fp, fp, fp, i32, vp, vp, vp]),
        "lds_num_steps": (i32, [vp]),
        "lds_workspace_bytes": (C.c_int64, [vp]),
        "lds_kernel_launches": (C.c_int64, [vp]),
        "lds_set_profiling": (i32, [vp, i32]),
        "lds_profile_num_classes": (i32, []),
        "lds_profile_class_name": (C.c_char_p, [i32]),
        "lds_profile_class_ms": (C.c_double, [vp, i32]),
        "lds_profile_class_launches": (C.c_int64, [vp, i32]),
        "lds_profile_class_flops": (C.c_double, [vp, i32]),
        "lds_profile_class_bytes": (C.c_double, [vp, i32]),
        "lds_op_gemm": (i32, [vp, i32, vp, vp, vp, i32, i32, vp, i32] + [i32] * 10 + [C.c_float, i32, vp]),
        "lds_op_attention": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "lds_op_groupnorm": (i32, [vp, i32, vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, i32, vp, vp, vp]),
        "lds_op_groupnorm_fused": (i32, [vp, i32, vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, i32, vp, vp]),
        "lds_op_groupnorm_cluster": (i32, [vp, i32, vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, i32, vp, vp]),
        "lds_op_layernorm": (i32, [vp, vp, vp, C.c_float, i32, i32, vp, vp]),
        "lds_op_split_cast": (i32, [vp, vp, C.c_int64, i32, i32, vp]),
        "lds_op_gemm_tc": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, vp]),
        "lds_op_conv1d_tc": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, i32, f32, vp]),
        "lds_op_qkv_attention_tc": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "lds_op_x0_pred": (i32, [vp, vp, f32, f32, vp, i64, vp]),
        "lds_op_dpm_update": (i32, [vp, vp, vp, f32, f32, f32, f32, i32, i64, vp]),
        "lds_op_unipc_predict": (i32, [vp, vp, vp, f32, f32, f32, f32, f32, i32, vp, vp, i64, vp]),
        "lds_op_unipc_correct": (i32, [vp, vp, vp, vp, f32, f32, f32, f32, i32, vp, i64, vp]),
        "lds_op_ddpm_step": (i32, [vp, vp, vp, f32, f32, f32, f32, f32, i32, i32, i32, vp]),
        "lds_op_ddim_step": (i32, [vp, vp, f32, f32, f32, i64, vp]),
        "lds_op_pndm_update": (i32, [vp, vp, vp, vp, vp, f32, f32, f32, i32, vp, i64, vp]),
        "lds_op_q_sample": (i32, [vp, vp, vp, f32, f32, f32, i32, i32, i32, vp]),
        "lds_op_cast_gather": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
        "lds_op_transpose": (i32, [vp, vp, i32, i32, i32, f32, i32, vp]),
        "lds_op_div_copy": (i32, [vp, vp, i64, f32, vp]),
        "lds_vocoder_last_error": (C.c_char_p, []),
        "lds_vocoder_create": (i32, [vp, i32, C.POINTER(vp)]),
        "lds_vocoder_destroy": (None, [vp]),
        "lds_vocoder_load_weight": (i32, [vp, C.c_char_p, vp, i64p, i32, i32]),
        "lds_vocoder_finalize": (i32, [vp]),
        "lds_vocode": (i32, [vp, vp, i32, i32, vp, vp]),
        "lds_vocoder_hop": (i32, [vp]),
        "lds_vocoder_launches": (C.c_int64, [vp]),
        "lds_vocoder_last_flops": (C.c_double, [vp]),
        "lds_vocoder_workspace_bytes": (C.c_int64, [vp]),
        "lds_units_last_error": (C.c_char_p, []),
        "lds_units_create": (i32, [vp, i32, C.POINTER(vp)]),
        "lds_units_destroy": (None, [vp]),
        "lds_units_load_weight": (i32, [vp, C.c_char_p, vp, i64p, i32, i32]),
        "lds_units_finalize": (i32, [vp]),
        "lds_units_encode": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "lds_units_out_frames": (i32, [i32]),
        "lds_units_launches": (C.c_int64, [vp]),
        "lds_units_last_flops": (C.c_double, [vp]),
        "lds_units_workspace_bytes": (C.c_int64, [vp]),
        "lds_units_log_mel": (i32, [vp, i32, i32, vp, i32, vp, vp, vp]),
        "lds_units_gather_rows": (i32, [vp, vp, i64, i64, i64, i32, vp, vp]),
        "lds_units_quantize": (i32, [vp, vp, i64, i32, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.lds_last_error().decode(errors="replace")
        raise LdsError(f"{what or 'lds call'} failed (status {rc}): {msg}")


def _fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Engine:
    """One lds_handle: replicated weights + workspace on one B200, driven from one Python thread."""

    def __init__(self, *, input_channel: int, n_spk: Optional[int], out_dims: int, n_layers: int,
                 block_out_channels, n_heads: int, n_hidden: int, acoustic_scale: float, device: torch.device,
                 precision: str = "fp32", norm_groups: int = 8):
        if device.type != "cuda":
            raise LdsError("the diffusion sampling path runs on a CUDA device (B200, sm_100a) only; got %s" % device)
        self.lib = load_library()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        cfg = LdsConfig()
        cfg.input_channel, cfg.n_spk = int(input_channel), int(n_spk or 0)
        cfg.out_dims, cfg.n_layers = int(out_dims), int(n_layers)
        ch = list(block_out_channels)
        cfg.n_blocks = len(ch)
        for i, c in enumerate(ch):
            cfg.block_out_channels[i] = int(c)
        cfg.n_heads, cfg.n_hidden, cfg.norm_groups = int(n_heads), int(n_hidden), int(norm_groups)
        cfg.acoustic_scale = float(acoustic_scale)
        cfg.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp32_ffma": PREC_FP32_FFMA}[precision]
        self.cfg, self.precision = cfg, precision
        self.handle = C.c_void_p()
        check(self.lib, self.lib.lds_create(C.byref(cfg), self.index, C.byref(self.handle)), "lds_create")
        self.plan_key = None
        self.B = self.T = None
        self._keep = []

    # ---- weights -------------------------------------------------------------------------
    def load_state_dict(self, sd) -> None:
        skip = ("decoder.betas", "decoder.alphas", "decoder.sqrt_", "decoder.log_one", "decoder.posterior_",
                "decoder.spec_m")
        for key, t in sd.items():
            if key.startswith(skip):
                continue
            t = t.detach()
            code = {torch.float32: DTYPE_F32, torch.bfloat16: DTYPE_BF16, torch.float16: DTYPE_F16}.get(t.dtype)
            if code is None:
                t, code = t.float(), DTYPE_F32
            t = t.contiguous()
            shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
            check(self.lib, self.lib.lds_load_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), shape,
                                                     t.dim(), code), f"lds_load_weight({key})")
        check(self.lib, self.lib.lds_finalize_weights(self.handle), "lds_finalize_weights")

    # ---- planning ------------------------------------------------------------------------
    def plan(self, B: int, T: int, sampler: int, t_sin: Optional[np.ndarray], coefs: Optional[np.ndarray], key) -> None:
        if key is not None and key == self.plan_key:
            return
        n_nfe = 0 if t_sin is None else int(t_sin.shape[0])
        n_rows = 0 if coefs is None else int(coefs.shape[0])
        ts = None if t_sin is None else np.ascontiguousarray(t_sin, dtype=np.float32)
        cf = None if coefs is None else np.ascontiguousarray(coefs, dtype=np.float32)
        # lds_plan synchronises the device itself, and only when it has to grow the workspace or replace the
        # time-conditioning table (include/lds_b200.h); there is no synchronisation here.
        with torch.cuda.device(self.index):
            check(self.lib, self.lib.lds_plan(self.handle, B, T, sampler, n_nfe, None if ts is None else _fptr(ts), n_rows,
                                              None if cf is None else _fptr(cf)), "lds_plan")
        self.plan_key, self.B, self.T = key, B, T

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype).contiguous()
        return t

    # ---- compute -------------------------------------------------------------------------
    def cond(self, units: torch.Tensor, spk_id: Optional[torch.Tensor]) -> torch.Tensor:
        B, T, _ = units.shape
        units = self._dev(units)
        out = torch.empty(B, T, self.cfg.n_hidden, device=self.device, dtype=torch.float32)
        sp = None
        if self.cfg.n_spk > 1:
            if spk_id is None:
                raise ValueError("spk_id is required when n_spk > 1")
            sp = self._dev(spk_id.reshape(-1), torch.int64)
            if sp.numel() != B:
                raise ValueError("spk_id must have one entry per utterance")
        check(self.lib, self.lib.lds_cond(self.handle, C.c_void_p(units.data_ptr()),
                                          C.c_void_p(sp.data_ptr()) if sp is not None else None,
                                          C.c_void_p(out.data_ptr()), self._stream()), "lds_cond")
        self._keep = [units, sp]
        return out

    def denoise(self, x_bmt: torch.Tensor, cond_bth: torch.Tensor, t_sin_row: np.ndarray) -> torch.Tensor:
        x, cond = self._dev(x_bmt), self._dev(cond_bth)
        eps = torch.empty_like(x)
        row = np.ascontiguousarray(t_sin_row, dtype=np.float32).reshape(-1)
        check(self.lib, self.lib.lds_denoise(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(cond.data_ptr()), _fptr(row),
                                             C.c_void_p(eps.data_ptr()), self._stream()), "lds_denoise")
        self._keep = [x, cond]
        return eps

    def train_loss(self, cond_bth: torch.Tensor, gt_spec_btm: torch.Tensor, noise_bmt: torch.Tensor, t_sin: np.ndarray,
                   sqrt_acp: np.ndarray, sqrt_1m_acp: np.ndarray, loss_type: str = "l2", return_eps: bool = False):
        """p_losses forward (diffusion.py:173-187): per-utterance q_sample, one denoiser evaluation with per-utterance timesteps,
        l1 / l2 loss against the noise.  Returns the 0-dim loss tensor (and the prediction [B, M, T] when asked)."""
        cond, gt, nz = self._dev(cond_bth), self._dev(gt_spec_btm), self._dev(noise_bmt)
        ts = np.ascontiguousarray(t_sin, dtype=np.float32)
        sa, sb = np.ascontiguousarray(sqrt_acp, dtype=np.float32), np.ascontiguousarray(sqrt_1m_acp, dtype=np.float32)
        if ts.shape[0] != self.B or sa.shape[0] != self.B or sb.shape[0] != self.B:
            raise ValueError("one timestep / coefficient pair per utterance is required")
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        eps = torch.empty_like(nz) if return_eps else None
        check(self.lib, self.lib.lds_train_loss(self.handle, C.c_void_p(cond.data_ptr()), C.c_void_p(gt.data_ptr()), C.c_void_p(nz.data_ptr()),
                                                _fptr(ts), _fptr(sa), _fptr(sb), {"l1": 1, "l2": 2}[loss_type], C.c_void_p(loss.data_ptr()),
                                                C.c_void_p(eps.data_ptr()) if eps is not None else None, self._stream()), "lds_train_loss")
        self._keep = [cond, gt, nz]
        return (loss, eps) if return_eps else loss

    def sample_begin(self, cond_bth: torch.Tensor, x_init_bmt: torch.Tensor) -> None:
        cond, x = self._dev(cond_bth), self._dev(x_init_bmt)
        check(self.lib, self.lib.lds_sample_begin(self.handle, C.c_void_p(cond.data_ptr()), C.c_void_p(x.data_ptr()),
                                                  self._stream()), "lds_sample_begin")
        self._keep = [cond, x]

    def sample_begin_shallow(self, cond_bth: torch.Tensor, gt_spec_btm: torch.Tensor, noise_bmt: torch.Tensor,
                             sqrt_acp: float, sqrt_1m_acp: float) -> None:
        """x <- sqrt_acp * norm_spec(gt_spec)^T + sqrt_1m_acp * noise  (q_sample at t = k_step - 1) in one kernel."""
        cond, gt, nz = self._dev(cond_bth), self._dev(gt_spec_btm), self._dev(noise_bmt)
        check(self.lib, self.lib.lds_sample_begin_shallow(self.handle, C.c_void_p(cond.data_ptr()), C.c_void_p(gt.data_ptr()),
                                                          C.c_void_p(nz.data_ptr()), float(sqrt_acp), float(sqrt_1m_acp),
                                                          self._stream()), "lds_sample_begin_shallow")
        self._keep = [cond, gt, nz]

    def sample_steps(self, k0: int, k1: int, step_noise: Optional[torch.Tensor] = None) -> None:
        nz = None if step_noise is None else self._dev(step_noise)
        check(self.lib, self.lib.lds_sample_steps(self.handle, k0, k1, C.c_void_p(nz.data_ptr()) if nz is not None else None,
                                                  self._stream()), "lds_sample_steps")
        # No keep-alive for the noise chunk: it is allocated on the stream the library enqueues on, so torch's caching
        # allocator already orders its reuse behind the kernels that read it (a 1000-step DDPM run would otherwise pin
        # every chunk: ~14 GB at B=32, T=864).

    def sample_end(self) -> torch.Tensor:
        mel = torch.empty(self.B, self.T, self.cfg.out_dims, device=self.device, dtype=torch.float32)
        check(self.lib, self.lib.lds_sample_end(self.handle, C.c_void_p(mel.data_ptr()), self._stream()), "lds_sample_end")
        return mel

    @property
    def num_steps(self) -> int:
        return int(self.lib.lds_num_steps(self.handle))

    # ---- introspection -------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        check(self.lib, self.lib.lds_set_profiling(self.handle, int(on)), "lds_set_profiling")

    def profile(self) -> dict:
        out = {}
        for i in range(self.lib.lds_profile_num_classes()):
            name = self.lib.lds_profile_class_name(i).decode()
            out[name] = dict(ms=self.lib.lds_profile_class_ms(self.handle, i),
                             launches=int(self.lib.lds_profile_class_launches(self.handle, i)),
                             flops=self.lib.lds_profile_class_flops(self.handle, i),
                             bytes=self.lib.lds_profile_class_bytes(self.handle, i))
        return out

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.lds_kernel_launches(self.handle))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.lds_workspace_bytes(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.lds_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
