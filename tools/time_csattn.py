#!/usr/bin/env python
"""CUDA-event timing of one tensor-core cross-scale attention call (module API) per shape:
    python tools/time_csattn.py            # C=64 192x192, C=180 192x192, C=64 B=16 48x48
Prints one JSON line per shape: ms (best of 5 after 2 warm-ups), algorithmic TFLOP/s (20.25 C (HW)^2 per image,
SURVEY.md 8d) and its fraction of the measured sustained dense bf16 peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ciaosr_b200 import synth  # noqa: E402
from ciaosr_b200.cross_scale_attention import CrossScaleAttention  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 1400.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("bf16_tflops_sustained", peak)
    for c, b, n in [(64, 1, 192), (64, 2, 192), (180, 1, 192), (64, 16, 48), (180, 4, 128)]:
        holder = torch.nn.Module()
        holder.cs_attn = CrossScaleAttention(channel=c, scale=[2])
        synth.fill_module(holder, 3)
        holder = holder.to(dev)
        feat = synth.synth_feature(b, c, n, n, 9).to(dev)
        for _ in range(2):
            holder.cs_attn(feat)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            holder.cs_attn(feat)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        flops = b * 20.25 * c * float(n * n) ** 2
        print(json.dumps(dict(C=c, B=b, HW=n, ms=round(best, 3), ms_per_image=round(best / b, 3),
                              algorithmic_tflops=round(flops / best / 1e9, 1),
                              frac_of_sustained_bf16=round(flops / best / 1e9 / peak, 3))), flush=True)
        del holder, feat
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
