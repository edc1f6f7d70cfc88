#!/usr/bin/env python
"""CUDA-event timing of the native SwinIR-trunk pieces at the size of one config-4 call (2 tiles of 192x192 =
73 728 tokens, C = 180): the four Linear shapes, window attention, LayerNorm, the NHWC 3x3 convolution."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ciaosr_b200 import native  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def main():
    dev = torch.device("cuda:0")
    b, h, w, c, heads, ws = 2, 192, 192, 180, 6, 8
    rows = b * h * w
    g = torch.Generator().manual_seed(0)
    out = {}
    for name, k, n, gelu, res in [("qkv", 180, 540, False, False), ("proj+res", 180, 180, False, True),
                                  ("fc1+gelu", 180, 360, True, False), ("fc2+res", 360, 180, False, True)]:
        x = torch.randn(rows, k, generator=g).to(dev)
        plan = native.LinearPlan((torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.zeros(n).to(dev))
        r = torch.randn(rows, n, generator=g).to(dev) if res else None
        us = timed(lambda: plan.forward(x, gelu=gelu, residual=r))
        out[name] = dict(us=round(us, 1), tflops_exec=round(3 * 2 * rows * k * n / us / 1e6, 1))
    # the same four shapes with the operand already in split form (TMA-fed A; what the block fast path runs)
    for name, k, n, gelu, res, so in [("qkv split-in", 180, 540, False, False, False), ("proj+res split-in", 180, 180, False, True, False),
                                      ("fc1+gelu split-in/out", 180, 360, True, False, True), ("fc2+res split-in", 360, 180, False, True, False)]:
        a = native.SplitTensor(rows, k, dev)
        a.hi.copy_(torch.randn(rows, a.ld, generator=g).half()); a.lo.zero_()
        plan = native.LinearPlan((torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.zeros(n).to(dev))
        r = torch.randn(rows, n, generator=g).to(dev) if res else None
        us = timed(lambda: plan.forward_split(a, gelu=gelu, residual=r, split_out=so))
        out[name] = dict(us=round(us, 1), tflops_exec=round(3 * 2 * rows * k * n / us / 1e6, 1))
    qkv = torch.randn(b, h * w, 3 * c, generator=g).to(dev)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g).to(dev)
    for shift in (0, 4):
        out[f"window_attention shift={shift}"] = dict(us=round(timed(
            lambda: native.window_attention(qkv, table, h, w, heads, ws, shift, 30 ** -0.5)), 1))
    ln = torch.nn.LayerNorm(c).to(dev)
    x = torch.randn(rows, c, generator=g).to(dev)
    out["layernorm"] = dict(us=round(timed(lambda: native.layernorm(x, ln)), 1))
    out["layernorm split-out"] = dict(us=round(timed(lambda: native.layernorm_split(x, ln)), 1))
    out["window_attention split-out"] = dict(us=round(timed(
        lambda: native.window_attention(qkv, table, h, w, heads, ws, 4, 30 ** -0.5, split_out=True)), 1))
    conv = torch.nn.Conv2d(c, c, 3, 1, 1).to(dev)
    cp = native.Conv3x3Plan(conv.weight, conv.bias)
    xm = x.view(b, h, w, c)
    us = timed(lambda: cp.forward(xm, residual=xm))
    out["conv3x3+res"] = dict(us=round(us, 1), tflops_exec=round(3 * 2 * rows * 9 * c * c / us / 1e6, 1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
