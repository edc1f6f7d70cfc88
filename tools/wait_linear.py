"""Diagnostic (GPU box, timing build): where the trunk's TMA-fed Linear kernels wait, per mbarrier class.
    bash tools/build_timing.sh && CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_timing.so python tools/wait_linear.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ciaosr_b200 import _lib, native

NAMES = {10: "producer: W_EMPTY", 12: "producer: A_FREE (TMA A)", 20: "issuer: D_FREE", 21: "issuer: A_READY",
         22: "issuer: W_FULL", 30: "rows: A_FREE", 31: "rows: D_READY"}
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)
rows = 2 * 192 * 192
cyc = (ctypes.c_ulonglong * 64)(); cnt = (ctypes.c_ulonglong * 64)()
for name, k, n, split_in in [("proj split-in", 180, 180, True), ("qkv split-in", 180, 540, True), ("proj fp32-in", 180, 180, False)]:
    plan = native.LinearPlan((torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.zeros(n).to(dev))
    if split_in:
        a = native.SplitTensor(rows, k, dev)
        a.hi.copy_(torch.randn(rows, a.ld, generator=g).half()); a.lo.zero_()
        fn = lambda: plan.forward_split(a)
    else:
        x = torch.randn(rows, k, generator=g).to(dev)
        fn = lambda: plan.forward(x)
    for _ in range(3):
        fn()
    lib.ciaosr_debug_wait_read_linear(cyc, cnt, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.ciaosr_debug_wait_read_linear(cyc, cnt, 1)
    print(f"{name}: {e0.elapsed_time(e1) * 1e3:.1f} us (timing build)")
    for i in range(64):
        if cnt[i]:
            print(f"  class {i:2d} {NAMES.get(i, '?'):26s} waits {cnt[i]:8d}  per-CTA Mcyc {cyc[i] / 148 / 1e6:7.3f}  avg {cyc[i] / cnt[i]:7.0f} cyc")
