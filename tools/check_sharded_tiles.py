"""torchrun --nproc-per-node N tools/check_sharded_tiles.py : tiled inference with tiles sharded over the
ranks (one all-gather of the predictions) must reproduce the single-rank result bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ciaosr_b200 import synth
from ciaosr_b200.builder import build
from ciaosr_b200.generators import LocalImplicitSRRDN
from ciaosr_b200.restorers import CiaoSR

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
mlp = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=[256, 256, 256, 256])
cfg = dict(type=CiaoSR,
           generator=dict(type=LocalImplicitSRRDN,
                          encoder=dict(type="RDN", in_channels=3, out_channels=3, mid_channels=64, num_blocks=4,
                                       upscale_factor=4, num_layers=4, channel_growth=64),
                          imnet_q=mlp(), imnet_k=mlp(), imnet_v=mlp(), eval_bsize=30000),
           rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1., 1., 1.), pixel_loss=dict(type="L1Loss"))
m = build(cfg, test_cfg=dict(scale=3, tile=64, tile_overlap=16))
synth.fill_module(m.generator, 5)
m = m.eval().to(dev)
lq = (synth.synth_lr_image(1, 160, 208, 5) + torch.tensor((0.4488, 0.4371, 0.4040)).view(1, 3, 1, 1)).to(dev)
for _ in range(2):
    out = m(lq=lq, gt=None, test_mode=True)["output"]
torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
out = m(lq=lq, gt=None, test_mode=True)["output"]
torch.cuda.synchronize(); t_sh = time.time() - t0
m.test_cfg["shard_tiles"] = False
for _ in range(2):
    ref = m(lq=lq, gt=None, test_mode=True)["output"]
torch.cuda.synchronize(); t0 = time.time()
ref = m(lq=lq, gt=None, test_mode=True)["output"]
torch.cuda.synchronize(); t_one = time.time() - t0
err = float((out - ref).abs().max())
if rank == 0:
    print(f"ranks={dist.get_world_size()} out={tuple(out.shape)} sharded {t_sh*1e3:.1f} ms vs single-rank {t_one*1e3:.1f} ms, "
          f"max-abs difference {err:.3e}")
assert err == 0.0
dist.destroy_process_group()
