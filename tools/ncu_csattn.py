"""Profiling target: one tensor-core cross-scale attention call on a 192x192 tile (run under ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ciaosr_b200 import synth
from ciaosr_b200.cross_scale_attention import CrossScaleAttention
c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
dev = torch.device("cuda:0")
holder = torch.nn.Module()
holder.cs_attn = CrossScaleAttention(channel=c, scale=[2])
synth.fill_module(holder, 3)
holder = holder.to(dev)
feat = synth.synth_feature(1, c, n, n, 9).to(dev)
for _ in range(2):
    holder.cs_attn(feat)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
holder.cs_attn(feat)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
