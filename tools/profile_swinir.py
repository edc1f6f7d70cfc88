#!/usr/bin/env python
"""Where the SwinIR trunk's time goes (BASELINE.json configs 4 / 5): torch.profiler CUDA-time table of ONE
gen_feature call on a batch of two 192x192 tiles (config 4's tile batch), eager, native Linear path on.
    python tools/profile_swinir.py [tile] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import run_configs as rc  # noqa: E402
from ciaosr_b200 import synth  # noqa: E402
from ciaosr_b200.builder import build  # noqa: E402


def main():
    tile = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, test_cfg = rc.swinir_model(dict(scale=3, tile=192, tile_overlap=32), real=False)
    m = build(cfg, test_cfg=test_cfg)
    synth.fill_module(m.generator, 0)
    g = m.eval().to(dev).generator
    x = synth.synth_lr_image(batch, tile, tile, 5).to(dev)
    with torch.no_grad():
        for _ in range(2):
            g.gen_feature(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.gen_feature(x)
        e1.record()
        torch.cuda.synchronize()
        print(f"gen_feature (eager) on {batch} x {tile}x{tile}: {e0.elapsed_time(e1):.2f} ms")
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            g.gen_feature(x)
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))


if __name__ == "__main__":
    main()
