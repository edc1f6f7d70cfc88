"""Diagnostic (GPU box): where the tcgen05 kernels wait.  Needs the timing build of the library:
    nvcc ... -DCIAOSR_TC_TIMING -o ciaosr_b200/csrc/libciaosr_b200_timing.so   (tools/build_timing.sh)
    CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_timing.so python tools/wait_breakdown.py
Prints, per mbarrier tag class (tag / 10), the warp-cycles spent waiting during ONE bench step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ciaosr_b200 import _lib

NAMES = {10: "producer: W_EMPTY", 20: "issuer: D_FREE", 21: "issuer: A_READY", 22: "issuer: W_FULL",
         30: "rows: A_FREE", 31: "rows: D_READY", 40: "conv producer: A_FREE", 41: "conv producer: W_EMPTY",
         42: "conv issuer: D_FREE", 43: "conv issuer: A_READY", 44: "conv issuer: W_FULL", 45: "conv rows: D_READY",
         46: "conv stagers: W_EMPTY"}
dev = torch.device("cuda:0")
model = bench.build_model("auto").to(dev)
lq, coord, cell = bench.make_inputs(bench.B, 100)
lq = (lq - torch.tensor(bench.RGB_MEAN).view(1, 3, 1, 1)).to(dev)
coord, cell = coord.to(dev), cell.to(dev)
lib = _lib.load()
cyc = (ctypes.c_ulonglong * 64)(); cnt = (ctypes.c_ulonglong * 64)()
with torch.no_grad():
    for _ in range(2):
        model.generator(lq, coord, cell, test_mode=True)
    for fn in ("ciaosr_debug_wait_read_head", "ciaosr_debug_wait_read_rdn"):
        getattr(lib, fn)(cyc, cnt, 1)
    model.generator(lq, coord, cell, test_mode=True)
    torch.cuda.synchronize()
for fn in ("ciaosr_debug_wait_read_head", "ciaosr_debug_wait_read_rdn"):
    getattr(lib, fn)(cyc, cnt, 1)
    print(fn)
    for i in range(64):
        if cnt[i]:
            print(f"  class {i:2d} {NAMES.get(i, '?'):28s} waits {cnt[i]:9d}  warp-cycles {cyc[i]:14d}  "
                  f"per SM {cyc[i] / 148 / 1e6:8.3f} Mcyc  avg {cyc[i] / cnt[i]:8.0f}")
