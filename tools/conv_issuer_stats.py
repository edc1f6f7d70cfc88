"""Diagnostic (GPU box): where the UMMA issuer of convw_tc_kernel spends its time over one RDN forward.  Needs a build of the
library with -DCIAOSR_CONV_STATS (register-accumulated clock64 deltas around the issuer's three mbarrier waits, summed over
all CTAs / layers; csrc/rdn_tc.cu):
    (cd ciaosr_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared \
        -DCIAOSR_CONV_STATS -o libciaosr_b200_stats.so *.cu)
    CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_stats.so python tools/conv_issuer_stats.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ciaosr_b200 import _lib
dev = torch.device("cuda:0")
model = bench.build_model("auto").to(dev)
model.generator.cuda_graph = False
lq, coord, cell = bench.make_inputs(bench.B, 100)
lq = (lq - torch.tensor(bench.RGB_MEAN).view(1, 3, 1, 1)).to(dev)
lib = _lib.load()
st = (ctypes.c_ulonglong * 8)()
with torch.no_grad():
    for _ in range(3):
        model.generator.gen_feature(lq)
    lib.ciaosr_debug_conv_stats(st, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.generator.gen_feature(lq); e1.record()
    torch.cuda.synchronize()
lib.ciaosr_debug_conv_stats(st, 1)
tw, ta, td, tot, taps, tiles, ctas = [st[i] for i in range(7)]
print(f"RDN forward {e0.elapsed_time(e1):.3f} ms; issuer warps {ctas}, tiles {tiles}, taps {taps}")
print(f"issuer cycles: total {tot}  wait W_FULL {tw} ({100 * tw / tot:.1f} %)  wait A_READY {ta} ({100 * ta / tot:.1f} %)  "
      f"wait D_FREE {td} ({100 * td / tot:.1f} %)  issuing {tot - tw - ta - td} ({100 * (tot - tw - ta - td) / tot:.1f} %)")
print(f"per tap: total {tot / taps:.0f} cycles, issuing {(tot - tw - ta - td) / taps:.0f}, waiting for weights {tw / taps:.0f}")
