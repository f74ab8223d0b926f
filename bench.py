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
`--impl reference` times the reference algorithm on the host CPU cores (the oracle port of the reference's
PyTorch sampler — the reference is pure Python and cannot travel to the GPU box) on a bounded sample.
"""
from __future__ import annotations

import argparse
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
CPU_SAMPLE = dict(B=1, T=864)      # bounded CPU sample of the same workload (one utterance of the batch)
FRAME_RATE = 44100 / 512


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


def cpu_oracle_run(B: int, T: int, method: str, speedup: int, threads: int):
    """One pass of the reference algorithm (oracle port) on the host CPU; returns seconds."""
    import torch
    from oracle import unit2mel_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel
    sd = {k: v.detach() for k, v in Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).state_dict().items()}
    units, spk, noise, _, _ = O.synthetic_inputs(B, T)

    def once():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.unit2mel_infer(sd, O.DEFAULT_CFG, units, spk, noise, method, speedup)
        return time.perf_counter() - t0
    return once


def run_reference(args):
    """--impl reference: the reference's CPU sampler (oracle port) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    B, T, method, speedup, _, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    once = cpu_oracle_run(CPU_SAMPLE["B"], T, method, speedup, cores)
    for _ in range(min(args.warmup, 1)):       # one CPU warm-up is enough to page the weights in
        once()
    times = [once() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = CPU_SAMPLE["B"] * T / sec
    sample = f"B={CPU_SAMPLE['B']} utterance of the batch x T={T}, {method} {1000 // speedup} NFE, fp32, torch CPU {cores} threads"
    line = {
        "impl": "reference", "metric": "mel_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sampler": method, "nfe": 1000 // speedup, "T": T, "batch_per_gpu": B,
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": sec / (CPU_SAMPLE["B"] * T / FRAME_RATE),
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from latent_diffusion_speech_b200.distributed import gather_mels
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel
    from oracle import unit2mel_oracle as O   # synthetic input generator + CPU baseline leg only

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

    # result landing zone in pinned host memory: rank 0 reads the gathered global mel, the other ranks their own shard
    mel_h = torch.empty((B * world if rank == 0 else B), T, 128, dtype=torch.float32).pin_memory()

    def step_e2e():
        u = units_h.to(dev, non_blocking=True)
        s = spk_h.to(dev, non_blocking=True)
        gt = None if gt_h is None else gt_h.to(dev, non_blocking=True)
        # noise: torch.randn on the device, as the reference draws it
        mel = model(u, None, spk_id=s, gt_spec=gt, k_step=k_step, infer=True, infer_speedup=speedup, method=method)
        full = gather_mels(mel, B * world) if world > 1 else mel
        mel_h.copy_(full if rank == 0 else mel, non_blocking=True)
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
    kern = {"fp32": "gemm_tc_kernel (tcgen05/TMEM/TMA implicit-GEMM conv k3 / 1x1 / linear, split-bf16 x6 products, fp32-accurate)",
            "bf16": "gemm_tc_kernel (tcgen05/TMEM/TMA implicit-GEMM conv k3 / 1x1 / linear, bf16 operands)",
            "fp32_ffma": "gemm_f32_kernel (implicit-GEMM conv k3 / 1x1 / linear, fp32 FFMA)"}[precision]
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")     # ncu dram bytes per gemm_tc launch (one capture per change)
    if os.path.exists(tp) and precision in ("fp32", "bf16") and (B, T) == (64, 864):
        tj = json.load(open(tp))
        traffic, traffic_src = tj[precision]["traffic_bytes_per_launch"], tj["source"]
    mma_per_product = 6.0 if precision == "fp32" else 1.0
    roofline = {
        "kernel": kern, "bound": "tensor",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": f"{peaks['source']} bf16 dense sustained (MEASURED_PEAKS.json); achieved counts ALGORITHMIC (logical fp32) FLOPs - "
                       "the split mode issues 6 bf16 MMAs per logical product, so its tensor-pipe occupancy is 6x this fraction",
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

    line = None
    if rank == 0:
        cpu_baseline = None
        # bounded CPU sample of the same workload; the 1000-step and shallow workloads would take minutes per utterance
        if world == 1 and not args.no_cpu_baseline and method is not None and k_step is None:
            cores = os.cpu_count() or 1
            once = cpu_oracle_run(CPU_SAMPLE["B"], T, method, speedup, cores)
            once()                                             # warm-up pass (pages the weights in, sizes the allocator)
            sec = sorted(once() for _ in range(3))[1]          # median of 3 (median of 3 after 1 warm-up)
            cpu_baseline = {"value": CPU_SAMPLE["B"] * T / sec, "unit": "frames/s", "cores": cores, "kind": "port",
                            "sample": f"B={CPU_SAMPLE['B']} x T={T}, {method} {nfe} NFE fp32, oracle port of the reference sampler, "
                                      "median of 3 passes after 1 warm-up"}
        line = {
            "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sampler": method, "nfe": nfe, "T": T, "batch_per_gpu": B,
                       "global_batch": B * world, "precision_mode": precision, "l2_policy": "inputs_exceed_l2 (units 283 MB/step, activations >1 GB)",
                       "parallelism": f"batch-shard x{world}, final NCCL all_gather" if world > 1 else "single GPU"},
            "rtf": (ms_step * 1e-3) / (frames / FRAME_RATE),
            "e2e": {"value": e2e_val, "unit": "frames/s",
                    "h2d_bytes_per_step": int(units_h.numel() * 4 + spk_h.numel() * 8 + (0 if gt_h is None else gt_h.numel() * 4)),
                    "d2h_bytes_per_step": int(B * world * T * 128 * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_classes": classes,
            "model_tflops_per_s": B * world * nfe * flops_per_utt_nfe(T) / (ms_step * 1e-3) / 1e12,
            "workspace_bytes": eng.workspace_bytes,
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
