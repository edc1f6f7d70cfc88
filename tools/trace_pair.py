"""Diagnostic (GPU box, timing build): timeline of CTA 0 of pair_mlp_kernel -- issuer warp and first row warp.
    CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_timing.so python tools/trace_pair.py > gpurun_out/trace.txt"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ciaosr_b200 import _lib
dev = torch.device("cuda:0")
model = bench.build_model("auto").to(dev)
lq, coord, cell = bench.make_inputs(bench.B, 100)
lq = (lq - torch.tensor(bench.RGB_MEAN).view(1, 3, 1, 1)).to(dev)
coord, cell = coord.to(dev), cell.to(dev)
lib = _lib.load()
if os.environ.get("CIAOSR_DBG_FLAGS"):
    lib.ciaosr_debug_flags(int(os.environ["CIAOSR_DBG_FLAGS"]))
buf = (ctypes.c_ulonglong * (2 * 8192))(); n = ctypes.c_uint(0)
with torch.no_grad():
    for _ in range(2):
        model.generator(lq, coord, cell, test_mode=True)
    lib.ciaosr_debug_trace(1, None, None)
    model.generator(lq, coord, cell, test_mode=True)
    torch.cuda.synchronize()
lib.ciaosr_debug_trace(0, buf, ctypes.byref(n))
raw = [(buf[2 * i + 1], buf[2 * i]) for i in range(min(n.value, 8192))]
ns = [t for t, tag in raw if tag == 9000]                       # %globaltimer at the tile starts
ev = sorted((t, tag) for t, tag in raw if tag != 9000)
cyc = [t for t, tag in ev if tag == 3000]
if len(ns) > 2 and len(cyc) == len(ns):
    print(f"# SM clock during the kernel: {(cyc[-1] - cyc[0]) / (ns[-1] - ns[0]) * 1e3:.0f} MHz "
          f"({cyc[-1] - cyc[0]} cycles in {ns[-1] - ns[0]} ns over {len(ns) - 1} tiles)")
t0 = ev[0][0]
NAMES = {1010: "ISSUER job issued", 2010: "rows  D drained", 3000: "rows  TILE START", 3001: "rows  k.L1 written",
         3002: "rows  k.L4 complete", 3003: "rows  v.L1 written", 3004: "rows  softmax done"}
for t, tag in ev:
    name = NAMES.get(tag) or (f"ISSUER slab {tag - 1000} ready" if 1000 <= tag < 1010 else
                              f"rows  slab {tag - 2020} written" if 2020 <= tag < 2030 else f"rows  D half {tag - 2000} ready")
    print(f"{t - t0:10d}  {name}")
