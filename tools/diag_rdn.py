"""Diagnostic (GPU box): native RDN encoder vs PyTorch fp32, max-abs error and time per shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
from ciaosr_b200 import synth
from ciaosr_b200.builder import build
from ciaosr_b200.generators import LocalImplicitSRRDN
dev = torch.device("cuda:0")
mlp = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=[256, 256, 256, 256])
nb, nl = int(os.environ.get("NB", 3)), int(os.environ.get("NL", 4))
g = build(dict(type=LocalImplicitSRRDN,
               encoder=dict(type="RDN", in_channels=3, out_channels=3, mid_channels=64, num_blocks=nb,
                            upscale_factor=4, num_layers=nl, channel_growth=64),
               imnet_q=mlp(), imnet_k=mlp(), imnet_v=mlp(), eval_bsize=30000))
synth.fill_module(g, 11)
g = g.eval().to(dev)
for b, h, w in [(2, 20, 17), (1, 48, 48), (16, 48, 48), (1, 12, 70)]:
    x = synth.synth_lr_image(b, h, w, 11).to(dev)
    with torch.no_grad():
        g.native_encoder = False
        ref = g.gen_feature(x)[0]
        g.native_encoder = True
        out = g.gen_feature(x)[0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            out = g.gen_feature(x)[0]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
    print(f"shape {(b, h, w)} max-abs {float((out - ref).abs().max()):.3e} ref-scale {float(ref.abs().mean()):.3f} "
          f"time {dt * 1e3:.3f} ms", flush=True)
