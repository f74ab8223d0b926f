"""Micro-benchmark of the fused QKV + tcgen05 attention op (not a test)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402

B = int(os.environ.get("AB", 8))
for (T, C) in [(864, 256), (432, 384), (216, 512)]:
    for parts in (2, 1):
        heads, d = 8, C // 8
        dpad = 32 if d <= 32 else 64
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B * T, C, generator=g).cuda()
        w = (torch.randn(3 * heads * dpad, C, generator=g) * C ** -0.5).cuda()
        wp, xp = G.pack_w_parts(w, 1, parts), G.op_split_cast(x, parts)
        t_pad = (T + 7) // 8 * 8
        ap = parts
        q = torch.zeros(B * T * ap * heads * dpad, device="cuda", dtype=torch.bfloat16)
        k = torch.zeros_like(q)
        vt = torch.zeros(B * ap * heads * dpad * t_pad, device="cuda", dtype=torch.bfloat16)
        out = torch.zeros(B * T, parts * C, device="cuda", dtype=torch.bfloat16)

        def run():
            G.check(G.lib().lds_op_qkv_attention_tc(G.ptr(xp), G.ptr(wp), B, T, C, heads, dpad, parts, G.ptr(q), G.ptr(k), G.ptr(vt),
                                                    G.ptr(out), G.stream()), "op")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * B * T * T * C
        print(f"T={T} C={C} parts={parts} B={B}: qkv+attention {ms:.3f} ms  attention-logical {fl / ms / 1e9:.1f} TFLOP/s (incl. qkv gemm time)", flush=True)
