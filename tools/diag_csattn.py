"""Diagnostic: tensor-core cross-scale attention vs the fp32 CUDA-core one (error and time) at tile sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ciaosr_b200 import synth
from ciaosr_b200.cross_scale_attention import CrossScaleAttention

dev = torch.device("cuda:0")
for c, b, h, w in [(64, 16, 48, 48), (64, 1, 192, 192), (180, 1, 192, 192), (64, 1, 256, 256)]:
    holder = torch.nn.Module()
    holder.cs_attn = CrossScaleAttention(channel=c, scale=[2])
    synth.fill_module(holder, 3)
    holder = holder.to(dev)
    feat = synth.synth_feature(b, c, h, w, 9).to(dev)
    holder.cs_attn(feat)
    plan = holder.cs_attn._plan[1]
    res = {}
    for eng in ("tcgen05", "simt"):
        plan.cross_scale_attention(feat, engine=eng)
        torch.cuda.synchronize()
        t0 = time.time()
        res[eng] = plan.cross_scale_attention(feat, engine=eng)
        torch.cuda.synchronize()
        res[eng + "_ms"] = (time.time() - t0) * 1e3
    d = (res["tcgen05"] - res["simt"]).abs()
    print(f"C={c} B={b} {h}x{w}: tc {res['tcgen05_ms']:.2f} ms, simt {res['simt_ms']:.1f} ms, "
          f"max-abs {float(d.max()):.2e}, mean-abs {float(d.mean()):.2e}, |out| mean {float(res['simt'].abs().mean()):.3f} "
          f"signed mean diff {float((res['tcgen05'] - res['simt']).mean()):.2e}", flush=True)
