#!/usr/bin/env python
"""bench.py — mel frames/s of the Unit2Mel diffusion-sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full ``Unit2Mel.forward(infer=True)`` over one batch of synthetic units (cond prep, every
denoiser evaluation, every solver update, final layout).  Workload at N=1 (default `dpm20_b64_t864_fp32`) is
BASELINE.json configs[1]: batch 64 x 10 s (T=864 frames), 20-step DPM-Solver++, fp32-accurate mode.  With N>1
(launched under torchrun, one rank per GPU) every rank processes its own 64 utterances (weak scaling, weights
replicated, no collective in the step loop) and the mels are all-gathered over NCCL inside the timed region.

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM; `e2e` = the same metric
through the public API with pinned-host inputs (H2D of units/spk_id and D2H of the mel inside the timed region).
`--impl reference` times the reference's own CPU sampler on the host cores on a bounded sample of the same workload: the
UNMODIFIED `diffusion.unit2mel.Unit2Mel` staged under `baseline/_ref/` by `oracle/build_ref.py` (`kind: "reference"`), or —
only when that directory is absent — the oracle port of it (`kind: "port"`).

Besides the headline the default N-GPU run also reports (same JSON line):
  `strong`             BASELINE configs[2] as BASELINE names it: global batch 512 x T=864, bf16, UniPC 10 NFE, split 512/N over the
                       ranks through `distributed.sharded_infer`, per-utterance seeded inputs and noise, SHA-256 of utterances 0 and
                       511 of the gathered mel (must be identical at every N)
  `gpu_eager_baseline` (N=1) the reference PyTorch eager sampler on the same B200 (fp32 with TF32 off; autocast bf16), B=64 x T=864
  `tf32_tflops_measured` (N=1) torch.matmul 8192^3 with TF32 on (the fp32-mode tensor peak SURVEY.md section 6 asks for)
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, method, infer_speedup, k_step, precision)
    "dpm20_b64_t864_fp32": (64, 864, "dpm-solver", 50, None, "fp32"),      # BASELINE configs[1] (headline)
    "dpm20_b8_t864_fp32": (8, 864, "dpm-solver", 50, None, "fp32"),
    "unipc10_b64_t864_fp32": (64, 864, "unipc", 100, None, "fp32"),
    "dpm20_b2_t216_fp32": (2, 216, "dpm-solver", 50, None, "fp32"),       # tiny, for plumbing checks
    "dpm20_b1_t432_fp32": (1, 432, "dpm-solver", 50, None, "fp32"),       # BASELINE configs[0] shape (single 5 s utterance, latency)
    "dpm20_b64_t864_ffma": (64, 864, "dpm-solver", 50, None, "fp32_ffma"),  # CUDA-core fp32 implementation (A/B)
    "dpm20_b64_t864_bf16": (64, 864, "dpm-solver", 50, None, "bf16"),
    "unipc10_b64_t864_bf16": (64, 864, "unipc", 100, None, "bf16"),        # per-GPU shard of BASELINE configs[2]
    "shallow_dpm20_b32_t2584_bf16": (32, 2584, "dpm-solver", 5, 100, "bf16"),   # per-GPU shard of BASELINE configs[3]
    "shallow_dpm20_b32_t2584_fp32": (32, 2584, "dpm-solver", 5, 100, "fp32"),
    "ddpm1000_b32_t864_fp32": (32, 864, None, 1, None, "fp32"),            # BASELINE configs[4]: 1000-step ancestral sampling
    "ddpm1000_b32_t864_bf16": (32, 864, None, 1, None, "bf16"),
    "ddim20_b64_t864_fp32": (64, 864, "ddim", 50, None, "fp32"),
    "pndm20_b64_t864_fp32": (64, 864, "pndm", 50, None, "fp32"),
}
HEADLINE = "dpm20_b64_t864_fp32"
CPU_SAMPLE_B = 4                   # utterances of the batch in the bounded CPU sample (same T, sampler and NFE as the workload)
FRAME_RATE = 44100 / 512
STRONG = dict(B=512, T=864, method="unipc", speedup=100, precision="bf16")    # BASELINE configs[2]


def flops_per_utt_nfe(T: int) -> float:
    """dense FLOPs of one denoiser evaluation of one utterance (8 | T)."""
    return 72.21e6 * T + 15424.0 * T * T + 0.041e9


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thr = threading.Thread(target=pump, daemon=True)
        self.thr.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def workload_config(name: str, world: int) -> dict:
    """The `config` object of the JSON line — identical for both arms (the reference arm times a bounded sample of it)."""
    B, T, method, speedup, k_step, precision = WORKLOADS[name]
    t_total = k_step if k_step is not None else 1000
    nfe = t_total // speedup + (1 if method == "pndm" else 0)
    return {"workload": name, "sampler": method, "nfe": nfe, "T": T, "batch_per_gpu": B, "global_batch": B * world,
            "precision_mode": precision, "l2_policy": "inputs_exceed_l2 (units 283 MB/step, activations >1 GB)",
            "parallelism": f"batch-shard x{world}, final NCCL all_gather" if world > 1 else "single GPU"}


class HostReference:
    """The reference sampler on a device of PyTorch's choosing: the UNMODIFIED reference `Unit2Mel` staged under baseline/_ref
    (kind "reference"), else the oracle port (kind "port").  Checker / baseline only — never on the product path."""

    def __init__(self, device="cpu"):
        import torch
        from oracle import ref_import
        self.torch, self.device = torch, torch.device(device)
        self.model = ref_import.build_reference_model(1234)
        if self.model is not None:
            self.kind, self._ri, self.root = "reference", ref_import, ref_import.REFERENCE_ROOT
            self.model = self.model.to(self.device)
        else:
            from oracle import unit2mel_oracle as O
            from latent_diffusion_speech_b200.unit2mel import Unit2Mel
            self.kind, self._O = "port", O
            torch.manual_seed(1234)
            self.sd = {k: v.detach().to(self.device) for k, v in Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).state_dict().items()}

    def inputs(self, B, T, k_step=None):
        torch = self.torch
        g = torch.Generator().manual_seed(7)
        units = torch.randn(B, T, 1280, generator=g).to(self.device)
        spk = torch.randint(1, 324, (B, 1), generator=g).to(self.device)
        gt = (torch.rand(B, T, 128, generator=g) * 14 - 12).to(self.device) if k_step is not None else None
        return units, spk, gt

    def run(self, units, spk, gt, method, speedup, k_step):
        torch = self.torch
        if self.kind == "reference":
            return self._ri.reference_infer(self.model, units, spk, method, speedup, gt_spec=gt, k_step=k_step)
        B, T = units.shape[:2]
        noise = torch.randn(B, 1, 128, T, device=self.device)
        n_steps = (k_step or 1000) if (method is None or speedup == 1) else 0
        steps = [torch.randn(B, 1, 128, T, device=self.device) for _ in range(n_steps)]
        with torch.no_grad():
            return self._O.unit2mel_infer(self.sd, self._O.DEFAULT_CFG, units, spk, noise, method, speedup, gt_spec=gt, k_step=k_step,
                                          step_noises=steps)

    def timer(self, B, T, method, speedup, k_step):
        units, spk, gt = self.inputs(B, T, k_step)

        def once():
            t0 = time.perf_counter()
            self.run(units, spk, gt, method, speedup, k_step)
            if self.device.type == "cuda":
                self.torch.cuda.synchronize()
            return time.perf_counter() - t0
        return once


def cpu_sample_batch(name: str) -> int:
    """Utterances in one bounded CPU step: 4 of the batch for the 20/10-NFE 10 s workloads (about 7 s on 16 cores), one for the
    1000-step and the 30 s ones."""
    B, T, method, speedup, k_step, _ = WORKLOADS[name]
    heavy = (method is None or speedup == 1) or T > 1000
    return 1 if heavy else min(CPU_SAMPLE_B, B)


def run_reference(args):
    """--impl reference: the reference's CPU sampler on all host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    B, T, method, speedup, k_step, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = HostReference("cpu")
    sb = cpu_sample_batch(args.workload)
    once = ref.timer(sb, T, method, speedup, k_step)
    for _ in range(args.warmup):
        once()
    times = [once() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = sb * T / sec
    cfg = workload_config(args.workload, int(os.environ.get("WORLD_SIZE", "1")))
    what = (f"unmodified reference diffusion.unit2mel.Unit2Mel ({os.path.relpath(ref.root, ROOT)})" if ref.kind == "reference"
            else "oracle port of the reference sampler")
    sample = (f"{sb} of the {B} utterances x T={T}, {method or 'ddpm'} {cfg['nfe']} NFE, fp32, {what}, torch CPU {cores} threads, "
              f"{args.warmup} warm-up + {args.steps} timed passes")
    line = {
        "impl": "reference", "metric": "mel_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": sec / (sb * T / FRAME_RATE),
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(name: str) -> dict:
    """Bounded CPU sample inside the GPU arm (rank 0, N=1): BASELINE configs[0] exactly (B=1, T=432, DPM-Solver 20 NFE; median of
    3 after 1 warm-up) and one bounded sample of the benched workload itself."""
    import torch
    B, T, method, speedup, k_step, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = HostReference("cpu")
    c1 = ref.timer(1, 432, "dpm-solver", 50, None)
    c1()
    sec1 = sorted(c1() for _ in range(3))[1]
    sb = cpu_sample_batch(name)
    once = ref.timer(sb, T, method, speedup, k_step)
    once()
    sec = once()
    what = (f"unmodified reference Unit2Mel ({os.path.relpath(ref.root, ROOT)})" if ref.kind == "reference"
            else "oracle port of the reference sampler")
    return {"value": sb * T / sec, "unit": "frames/s", "cores": cores, "kind": ref.kind,
            "sample": f"{sb} x T={T}, {method or 'ddpm'} fp32, {what}; one timed pass after one warm-up",
            "config1_b1_t432_dpm20": {"value": 432 / sec1, "unit": "frames/s", "seconds": sec1, "rtf": sec1 / (432 / FRAME_RATE),
                                      "sample": "BASELINE configs[0] exactly: B=1 x T=432, DPM-Solver 20 NFE; median of 3 after 1 warm-up"}}


def gpu_eager_leg(dev, B: int, T: int, method: str, speedup: int) -> dict:
    """SURVEY.md §0.1 'the bar to beat': the reference sampler as PyTorch eager dispatches it on this B200 (cuDNN / cuBLAS / SDPA),
    in IEEE fp32 (TF32 off — the only setting that meets the fp32 parity bar) and under autocast(bf16)."""
    import torch
    out = {"B": B, "T": T, "sampler": method, "nfe": 1000 // speedup}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        ref = HostReference(dev)
        out["kind"] = ref.kind
        units, spk, gt = ref.inputs(B, T)
        for mode in ("fp32_ieee", "autocast_bf16", "fp32_tf32"):
            on = mode == "fp32_tf32"
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = on
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "autocast_bf16" else contextlib.nullcontext()
            try:
                with ctx:
                    ref.run(units, spk, gt, method, speedup, None)        # warm-up (cuDNN autotune, allocator)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(2):
                        ref.run(units, spk, gt, method, speedup, None)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                out[mode] = {"frames_per_sec": B * T / (ms * 1e-3), "ms_per_step": ms}
            except Exception as ex:   # e.g. the dtype clash the survey notes for non-autocast bf16
                out[mode] = {"error": repr(ex)[:200]}
        del ref
    except Exception as ex:
        out["error"] = repr(ex)[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()
    return out


def tf32_peak(dev) -> dict:
    """torch.matmul fp32 8192^3 with TF32 allowed: best of 10 (burst) and back to back for 2 s (sustained), like MEASURED_PEAKS.json."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(2000 / best))
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"burst": fl / (best * 1e-3) / 1e12, "sustained": fl / (sus * 1e-3) / 1e12, "how": "torch.matmul fp32 8192^3, allow_tf32=True"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()


def vocoder_leg(dev, B: int = 16, T: int = 864) -> dict:
    """HiFi-VAEGAN Generator decode of B x T latent frames to 44.1 kHz audio through the library
    (csrc/vocoder.cu: the transposed convolutions and every ResBlock level as split-f16 tensor-core implicit GEMMs — the 32-channel level
    time-folded onto 64-wide K blocks — conv_post fp32 FFMA) — the step that turns
    the headline's mel frames into waveforms; random-init weights of the HiFi-GAN V1 layout (latent_diffusion_speech_b200/vocoder.py
    DEFAULT_H)."""
    import torch
    from latent_diffusion_speech_b200.vocoder import Vocoder
    torch.manual_seed(1234)
    voc = Vocoder("hifi-vaegan", None, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    mel = torch.randn(B, T, voc.dimension, generator=g, device=dev)
    for _ in range(2):
        voc.infer(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        wav = voc.infer(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    eng = voc.generator._engine
    out = {"workload": f"hifi-vaegan generator decode, B={B} x T={T} frames -> {wav.shape[-1]} samples each, fp32", "ms_per_step": ms,
           "value": B * T / (ms * 1e-3), "unit": "frames/s", "rtf": (ms * 1e-3) / (B * T / FRAME_RATE),
           "tflops": eng.last_flops / (ms * 1e-3) / 1e12, "flops_per_frame": eng.last_flops / (B * T),
           "arithmetic": "transposed convolutions + all ResBlock levels (32-channel level time-folded): split-f16 tcgen05 implicit GEMMs (fp32-accurate), conv_pre too; conv_post: fp32 FFMA",
           "finite": bool(torch.isfinite(wav).all())}
    voc.generator._invalidate()
    del voc, mel, wav
    torch.cuda.empty_cache()
    return out


def units_leg(dev, B: int = 8, L: int = 3000) -> dict:
    """the units front-end at whisper-large-v3 size (32 layers x 1280 channels x 20 heads, random init) —
    B x 30 s of 16 kHz audio -> log-mel (csrc/units.cu) -> AudioEncoder on the tcgen05 kernels -> units [B, 1500, 1280] -> nearest
    alignment to the mel frame grid (2584 frames).  Context: the oracle port of the reference encoder as PyTorch eager dispatches it
    on the same GPU (fp32 with TF32 off; autocast bf16)."""
    import torch
    from latent_diffusion_speech_b200 import units as UN
    d = UN.LARGE_V3
    torch.manual_seed(1234)
    enc = UN.AudioEncoder(d.n_mels, d.n_audio_state, d.n_audio_head, d.n_audio_layer).eval().to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    tt = torch.arange(L * UN.HOP_LENGTH, device=dev, dtype=torch.float32) / 16000.0
    f0 = 100.0 + 80.0 * torch.rand(B, 1, generator=g, device=dev)                 # synthetic "speech": harmonics under a slow envelope + noise
    audio = sum((0.6 ** h) * torch.sin(2 * 3.14159265 * (h + 1) * f0 * tt) for h in range(6)) * (0.5 + 0.5 * torch.sin(2 * 3.14159265 * 1.3 * tt)) ** 2
    audio = (0.3 * audio + 1e-3 * torch.randn(B, tt.numel(), generator=g, device=dev)).clamp(-1, 1).contiguous()
    T = (L - 1) // 2 + 1
    n_align = int(T * 320 / 16000 * FRAME_RATE)

    spread = {}

    def timed(fn, n=3, warm=2, key=None):
        """Median of n individually timed calls (a 25-60 ms call right after a power-capped leg sees clock dips that a short
        mean would fold in); min / max of the samples are kept under `key`."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            r = fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
        if key:
            spread[key] = {"ms_min": ms[0], "ms_max": ms[-1], "samples": n}
        return ms[n // 2], r

    out = {"workload": f"whisper-large-v3 AudioEncoder (32 x 1280 x 20 heads, random init), B={B} x 30 s audio -> {T} units each -> "
                       f"{n_align} aligned frames", "unit": "units/s (encoder frames, 50 per second of audio)"}
    ms_mel, mel = timed(lambda: UN.log_mel_spectrogram(audio, n_mels=d.n_mels))
    out["log_mel_ms"] = ms_mel
    for prec in ("fp32", "bf16"):
        enc.set_precision(prec)
        ms, units = timed(lambda: enc(mel), n=9, warm=3, key=prec)
        fl = enc._engine.last_flops
        out[prec] = {"ms_per_step": ms, "value": B * T / (ms * 1e-3), "tflops": fl / (ms * 1e-3) / 1e12, "launches_per_step": None,
                     "rtf": (ms * 1e-3) / (B * L * UN.HOP_LENGTH / 16000.0), "finite": bool(torch.isfinite(units).all()),
                     "arithmetic": "split-f16 tcgen05 (fp32-accurate, 3 MMAs per product)" if prec == "fp32" else "bf16 tcgen05"}
        out[prec].update(spread[prec])
        n0 = enc._engine.kernel_launches
        enc(mel)
        out[prec]["launches_per_step"] = enc._engine.kernel_launches - n0
    ms_al, _ = timed(lambda: UN.units_forced_alignment(units, n_frames=n_align))
    out["align_ms"] = ms_al
    out["flops_per_unit"] = fl / (B * T)
    enc._invalidate()
    from oracle import units_oracle as UO           # baseline leg only: the reference encoder restated, as PyTorch eager runs it
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    eager = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            ms, _ = timed(lambda: UO.audio_encoder(sd, d.n_audio_head, mel), n=1)
            eager["fp32_ieee"] = {"ms_per_step": ms, "value": B * T / (ms * 1e-3)}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms, _ = timed(lambda: UO.audio_encoder(sd, d.n_audio_head, mel), n=1)
            eager["autocast_bf16"] = {"ms_per_step": ms, "value": B * T / (ms * 1e-3)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["gpu_eager_baseline"] = dict(eager, kind="port (oracle/units_oracle.py: the reference AudioEncoder restated; the reference "
                                                 "module itself hard-codes .to('cuda') and needs no port on a GPU box, but is not staged)")
    del enc, sd, mel, units, audio
    torch.cuda.empty_cache()
    return out


def train_loss_leg(model, units, spk, dev, steps: int = 3) -> dict:
    """SURVEY.md §8(f) rank 4, forward half: Unit2Mel.forward(infer=False) — per-utterance q_sample, ONE denoiser evaluation with
    per-utterance timesteps and the l2 loss (diffusion.py:173-201) — on the headline batch; no backward pass exists in this library."""
    import torch
    B, T = units.shape[0], units.shape[1]
    g = torch.Generator(device=dev).manual_seed(11)
    gt = torch.rand(B, T, 128, generator=g, device=dev) * 14 - 12
    noise = torch.randn(B, 1, 128, T, generator=g, device=dev)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(5))
    for _ in range(2):
        loss = model(units, None, spk_id=spk, gt_spec=gt, infer=False, t=t, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = model(units, None, spk_id=spk, gt_spec=gt, infer=False, t=t, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": f"Unit2Mel.forward(infer=False): p_losses forward, B={B} x T={T}, per-utterance timesteps, l2", "ms_per_step": ms,
            "value": B * T / (ms * 1e-3), "unit": "frames/s", "loss": float(loss), "model_tflops_per_s": B * flops_per_utt_nfe(T) / (ms * 1e-3) / 1e12,
            "note": "forward only (validation loss, diffusion/solver.py:56-62); backward / optimiser / DP all-reduce are not built"}


def strong_leg(dev, world: int, rank: int, steps: int) -> dict:
    """BASELINE configs[2]: global batch 512 x T=864, bf16 mode, UniPC 10 NFE, split 512/N across the ranks (strong scaling) through
    distributed.sharded_infer; inputs and noise are seeded PER UTTERANCE (global index), so the gathered mel is bit-identical at
    every N — the SHA-256 of utterances 0 and 511 in the line proves it across the driver's N = 1, 2, 4, 8 runs."""
    import hashlib
    import torch
    import torch.distributed as dist
    from latent_diffusion_speech_b200.distributed import shard_bounds, gather_mels
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel
    Bg, T = STRONG["B"], STRONG["T"]
    lo, hi = shard_bounds(Bg, world, rank)
    torch.manual_seed(1234)
    model = Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval().to(dev).set_precision(STRONG["precision"])
    units = torch.empty(hi - lo, T, 1280, device=dev)
    noise = torch.empty(hi - lo, 1, 128, T, device=dev)
    spk = torch.empty(hi - lo, 1, dtype=torch.int64, device=dev)
    g = torch.Generator(device=dev)
    for b in range(lo, hi):                                  # global utterance index -> seed
        g.manual_seed(100000 + b)
        units[b - lo] = torch.randn(T, 1280, generator=g, device=dev)
        noise[b - lo] = torch.randn(1, 128, T, generator=g, device=dev)
        spk[b - lo, 0] = 1 + (b * 7919) % 323

    def step():
        mel = model(units, None, spk_id=spk, infer=True, infer_speedup=STRONG["speedup"], method=STRONG["method"], noise=noise)
        return gather_mels(mel, Bg) if world > 1 else mel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(2):
            full = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            full = step()
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    out = {"workload": "unipc10_b512_t864_bf16 (BASELINE configs[2])", "scaling": "strong", "global_batch": Bg, "batch_per_gpu": hi - lo,
           "T": T, "nfe": 10, "dtype": "bf16", "ms_per_step": ms, "value": Bg * T / (ms * 1e-3), "unit": "frames/s", "steps": steps, "warmup": 2,
           "model_tflops_per_s": Bg * 10 * flops_per_utt_nfe(T) / (ms * 1e-3) / 1e12}
    if rank == 0:
        sha = lambda t: hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]
        out["sha256_utt0"], out["sha256_utt_last"] = sha(full[0]), sha(full[Bg - 1])
    model.invalidate_engine()
    del model, units, noise, full
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from latent_diffusion_speech_b200.distributed import gather_mels
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, method, speedup, k_step, precision = WORKLOADS[args.workload]
    t_total = k_step if k_step is not None else 1000
    nfe = t_total // speedup + (1 if method == "pndm" else 0)

    torch.manual_seed(1234)
    model = Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval().to(dev)
    model.set_precision(precision)
    # this rank's shard of the global batch (global utterance index = rank*B + b -> shard-invariant noise)
    g = torch.Generator().manual_seed(7 + rank)
    units_h = torch.randn(B, T, 1280, generator=g).pin_memory()
    spk_h = torch.randint(1, 324, (B, 1), generator=g).pin_memory()
    gn = torch.Generator(device=dev).manual_seed(1000 + rank)
    noise = torch.randn(B, 1, 128, T, generator=gn, device=dev)
    units_d, spk_d = units_h.to(dev), spk_h.to(dev)
    eng = model._get_engine(dev)
    gt_h = (torch.rand(B, T, 128, generator=g) * 14 - 12).pin_memory() if k_step is not None else None   # shallow diffusion start
    gt_d = None if gt_h is None else gt_h.to(dev)

    def step_resident():
        mel = model(units_d, None, spk_id=spk_d, gt_spec=gt_d, k_step=k_step, infer=True, infer_speedup=speedup, method=method,
                    noise=noise)
        return gather_mels(mel, B * world) if world > 1 else mel

    # result landing zone in pinned host memory: every rank (= process) reads its OWN shard back over its own PCIe link, in
    # parallel; the NVLink all-gather still runs, so the whole batch is also resident on every GPU (north star: "a final NCCL
    # gather of mels"), but no single rank serialises a D2H of the global mel behind it (r01: 1.5 % of the N=8 step).
    mel_h = torch.empty(B, T, 128, dtype=torch.float32).pin_memory()

    def step_e2e():
        u = units_h.to(dev, non_blocking=True)
        s = spk_h.to(dev, non_blocking=True)
        gt = None if gt_h is None else gt_h.to(dev, non_blocking=True)
        # noise: torch.randn on the device, as the reference draws it
        mel = model(u, None, spk_id=s, gt_spec=gt, k_step=k_step, infer=True, infer_speedup=speedup, method=method)
        mel_h.copy_(mel, non_blocking=True)
        if world > 1:
            gather_mels(mel, B * world)
        torch.cuda.current_stream().synchronize()
        return mel_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        l0 = eng.kernel_launches
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms_total = timed(step_resident, args.steps)
        clk = clocks.stop() if rank == 0 else {}
        launches = eng.kernel_launches - l0
        # per-kernel-class device time over K more steps (CUDA events between launches on the launch stream)
        eng.set_profiling(True)
        prof = {}
        for _ in range(args.steps):
            step_resident()
            torch.cuda.synchronize()
            for k, v in eng.profile().items():
                a = prof.setdefault(k, dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
                for f in a:
                    a[f] += v[f]
        eng.set_profiling(False)
        for _ in range(1):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

    frames = B * T * world
    ms_step = ms_total / args.steps
    value = frames / (ms_step * 1e-3)
    e2e_val = frames / (ms_e2e / args.steps * 1e-3)
    peaks = measured_peaks()

    # dominant kernel: the implicit-GEMM (conv k3 + linear/1x1 share one kernel)
    gm = {f: prof["conv_k3_gemm"][f] + prof["linear_gemm"][f] for f in ("ms", "launches", "flops", "bytes")}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    achieved_tf = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    kern = {"fp32": "gemm_tc_kernel (tcgen05/TMEM/TMA implicit-GEMM conv k3 / 1x1 / linear, split-f16: two fp16 planes per operand, 3 tensor-core products per logical product, fp32-accurate)",
            "bf16": "gemm_tc_kernel (tcgen05/TMEM/TMA implicit-GEMM conv k3 / 1x1 / linear, bf16 operands)",
            "fp32_ffma": "gemm_f32_kernel (implicit-GEMM conv k3 / 1x1 / linear, fp32 FFMA)"}[precision]
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")     # ncu dram bytes per gemm_tc launch (one capture per change)
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tp) and precision in ("fp32", "bf16") and (B, T) == (64, 864):
        tj = json.load(open(tp))
        traffic, traffic_src = tj[precision]["traffic_bytes_per_launch"], tj["source"]
    mma_per_product = 3.0 if precision == "fp32" else 1.0
    roofline = {
        "kernel": kern, "bound": "tensor",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": f"{peaks['source']} bf16 dense sustained (MEASURED_PEAKS.json); achieved counts ALGORITHMIC (logical fp32) FLOPs - "
                       "the split-f16 mode issues 3 fp16 MMAs per logical product, so its tensor-pipe occupancy is 3x this fraction",
        "flops_per_launch": gm["flops"] / max(1, gm["launches"]), "avg_launch_ms": gm["ms"] / max(1, gm["launches"]),
        "share_of_step": gm["ms"] / tot_ms, "traffic": traffic, "traffic_source": traffic_src,
        "tensor_pipe_frac": achieved_tf * mma_per_product / peak_tf if precision != "fp32_ffma" else None,
    }
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    classes = {}
    for k, v in prof.items():
        tf = (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None
        gbs = (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 and v["flops"] == 0 else None
        classes[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps, "tflops": tf, "gbs": gbs,
                      # algorithmic bytes / time against the measured copy bandwidth (HBM-bound classes)
                      "hbm_frac": None if gbs is None else gbs / hbm_peak}

    # ---- context legs (outside the timed regions above) ----
    workspace_bytes = eng.workspace_bytes
    train_loss = None
    if world == 1 and not args.no_train_loss and precision != "fp32_ffma":
        try:
            train_loss = train_loss_leg(model, units_d, spk_d, dev)
        except Exception as ex:
            train_loss = {"error": repr(ex)[:300]}
    model.invalidate_engine()
    del model, eng, units_d, noise
    torch.cuda.empty_cache()
    strong = None
    if not args.no_strong:
        try:
            strong = strong_leg(dev, world, rank, max(2, min(args.steps, 5)))
        except Exception as ex:      # reported, never fatal for the headline
            strong = {"error": repr(ex)[:300]}
    line = None
    if rank == 0:
        cpu_baseline = gpu_eager = tf32 = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_baseline_leg(args.workload)
        if world == 1 and not args.no_gpu_eager and method is not None and k_step is None:
            gpu_eager = gpu_eager_leg(dev, B, T, method, speedup)
            tf32 = tf32_peak(dev)
        vocoder = None
        if world == 1 and not args.no_vocoder:
            try:
                vocoder = vocoder_leg(dev)
            except Exception as ex:
                vocoder = {"error": repr(ex)[:300]}
        units_fe = None
        if world == 1 and not args.no_units:
            try:
                units_fe = units_leg(dev)
            except Exception as ex:
                units_fe = {"error": repr(ex)[:300]}
        pipeline = None
        try:   # seconds of GPU time per second of audio through the three stages that now run on the library (resident inputs)
            parts = {"units_frontend": units_fe["fp32"]["rtf"] + (units_fe["log_mel_ms"] + units_fe["align_ms"]) * 1e-3 / (8 * 30.0),
                     "sampler": (ms_step * 1e-3) / (frames / FRAME_RATE), "vocoder": vocoder["rtf"]}
            pipeline = dict(parts, rtf_total=sum(parts.values()),
                            note="audio -> log-mel -> Whisper encoder -> alignment | Unit2Mel sampler (this workload) | HiFi-VAEGAN decode; "
                                 "each stage at its own batch size, fp32-accurate mode")
        except Exception:
            pass
        line = {
            "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "rtf": (ms_step * 1e-3) / (frames / FRAME_RATE),
            "e2e": {"value": e2e_val, "unit": "frames/s",
                    "h2d_bytes_per_step": int(world * (units_h.numel() * 4 + spk_h.numel() * 8 + (0 if gt_h is None else gt_h.numel() * 4))),
                    "d2h_bytes_per_step": int(B * world * T * 128 * 4), "ms_per_step": ms_e2e / args.steps,
                    "note": "bytes are whole-job totals; each rank copies its own shard in (pinned host -> HBM) and out, in parallel"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager,
            "tf32_tflops_measured": tf32,
            "strong": strong,
            "vocoder": vocoder,
            "units_frontend": units_fe,
            "train_loss_forward": train_loss,
            "pipeline_rtf": pipeline,
            "kernel_classes": classes,
            "model_tflops_per_s": B * world * nfe * flops_per_utt_nfe(T) / (ms_step * 1e-3) / 1e12,
            "workspace_bytes": workspace_bytes,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-vocoder", action="store_true")
    ap.add_argument("--no-units", action="store_true")
    ap.add_argument("--no-train-loss", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:   # convenience: self-launch one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
