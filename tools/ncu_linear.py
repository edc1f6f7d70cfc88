"""Profiling target: the SwinIR trunk's proj Linear (K = 180, N = 180) and qkv Linear (N = 540) on 73 728 rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ciaosr_b200 import native
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rows = 2 * 192 * 192
plans = []
for k, n in [(180, 180), (180, 540)]:
    x = torch.randn(rows, k, generator=g).to(dev)
    plans.append((native.LinearPlan((torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.zeros(n).to(dev)), x))
for p, x in plans:
    for _ in range(3):
        p.forward(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for p, x in plans:
    p.forward(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
