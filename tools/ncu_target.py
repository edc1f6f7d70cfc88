"""Profiling target: warm up, then one profiled step of bench.py's workload (cudaProfilerStart/Stop
delimit the captured region; run under `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
model = bench.build_model(os.environ.get("CIAOSR_ENGINE", "auto")).to(dev)
lq, coord, cell = bench.make_inputs(bench.B, 100)
lq = (lq - torch.tensor(bench.RGB_MEAN).view(1, 3, 1, 1)).to(dev)
coord, cell = coord.to(dev), cell.to(dev)
with torch.no_grad():
    for _ in range(3):
        model.generator(lq, coord, cell, test_mode=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.generator(lq, coord, cell, test_mode=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
