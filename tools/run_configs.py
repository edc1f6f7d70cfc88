#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 at their full sizes on the GPU (synthetic weights and images).

    python tools/run_configs.py [--configs 3,4,5] [--check 1] [--out gpurun_out/configs.jsonl]
    python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py --configs 4,5      # tiles sharded

They are the parity-test cases beside the bench line (bench.py times configs[1]); this tool runs them once at
BASELINE size, times the restorer call a user makes (`model(lq=..., test_mode=True)`: encoder + head + tile
blend, result resident on the device; CUDA events, 1 warm-up + `--reps` timed runs) and -- with `--check` --
compares the product path (tcgen05 engine, native RDN encoder) against the all-fp32 CUDA-core engine on the
same inputs: max-abs and PSNR between the two outputs (the fp32 engine is pinned to the reference goldens at
3.5e-6, tests/test_gpu_parity.py).  One JSON line per case.

  3  RDN-CiaoSR, 256x256 LR, x2/x3/x4 through clip_test (tile 192 / overlap 32, configs/001_*rdn*.py:48),
     x6/x8 un-tiled (:50)
  4  SwinIR-CiaoSR config 001 (C = 180, cross-scale attention on), 1280x720 -> x3 = 3840x2160, tile 192/32
  5  real-world config 002 w/o GAN (SwinIR, no cross-scale attention, no residual, EMA generator),
     1920x1080 -> x4 = 7680x4320, tile 128/32
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RGB_MEAN = (0.4488, 0.4371, 0.4040)


def mlp():
    return dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=[256, 256, 256, 256])


def rdn_model(test_cfg):
    from ciaosr_b200.generators import LocalImplicitSRRDN
    from ciaosr_b200.restorers import CiaoSR
    return dict(type=CiaoSR,
                generator=dict(type=LocalImplicitSRRDN,
                               encoder=dict(type="RDN", in_channels=3, out_channels=3, mid_channels=64,
                                            num_blocks=16, upscale_factor=4, num_layers=8, channel_growth=64),
                               imnet_q=mlp(), imnet_k=mlp(), imnet_v=mlp(), feat_unfold=True, eval_bsize=30000),
                rgb_mean=RGB_MEAN, rgb_std=(1., 1., 1.), pixel_loss=dict(type="L1Loss")), test_cfg


def swinir_encoder():
    from ciaosr_b200.swinir import SwinIR
    return dict(type=SwinIR, upscale=4, in_chans=3, img_size=48, window_size=8, img_range=1.,
                depths=[6] * 6, embed_dim=180, num_heads=[6] * 6, mlp_ratio=2, upsampler="pixelshuffle",
                resi_connection="1conv")


def swinir_model(test_cfg, real):
    from ciaosr_b200.generators import LocalImplicitSRSWINIR
    from ciaosr_b200.restorers import CiaoSR, RealCiaoSR
    gen = dict(type=LocalImplicitSRSWINIR, window_size=8, encoder=swinir_encoder(), imnet_q=mlp(), imnet_k=mlp(),
               imnet_v=mlp(), feat_unfold=True, eval_bsize=30000)
    if real:        # configs/002_real_wogan_*.py:54-59
        gen.update(local_ensemble_coord=True, imnet_k_type="mul_w", imnet_v_type="mul_w", res=False,
                   non_local_attn=False, cat_nla_v=False)
        return dict(type=RealCiaoSR, generator=gen, rgb_mean=RGB_MEAN, rgb_std=(1., 1., 1.),
                    pixel_loss=dict(type="L1Loss"), is_use_sharpened_gt_in_pixel=True, is_use_ema=True), test_cfg
    return dict(type=CiaoSR, generator=gen, rgb_mean=RGB_MEAN, rgb_std=(1., 1., 1.),
                pixel_loss=dict(type="L1Loss")), test_cfg


def cases(which):
    out = []
    if 3 in which:
        for s in (2, 3, 4):
            out.append(dict(config=3, name=f"RDN-CiaoSR 256x256 x{s} (tile 192/32)", h=256, w=256, scale=s,
                            build=lambda s=s: rdn_model(dict(scale=s, tile=192, tile_overlap=32))))
        for s in (6, 8):
            out.append(dict(config=3, name=f"RDN-CiaoSR 256x256 x{s} (un-tiled)", h=256, w=256, scale=s,
                            build=lambda s=s: rdn_model(dict(scale=s))))
    if 4 in which:
        out.append(dict(config=4, name="SwinIR-CiaoSR 001, 1280x720 -> x3 (tile 192/32)", h=720, w=1280, scale=3,
                        build=lambda: swinir_model(dict(scale=3, tile=192, tile_overlap=32), real=False)))
    if 5 in which:
        out.append(dict(config=5, name="real-world 002 w/o GAN, 1920x1080 -> x4 (tile 128/32)", h=1080, w=1920,
                        scale=4, build=lambda: swinir_model(dict(scale=4, tile=128, tile_overlap=32), real=True)))
    return out


def calibrate_features(gen, dev, target_std=0.5):
    """Rescale the synthetic SwinIR trunk so that its features have the O(1) spread the synthetic RDN has
    (std ~0.5): with random weights the transformer trunk otherwise feeds the head features whose products
    push the pre-clamp RGB to +-4, far outside the image range the 1e-4 parity tolerance is stated for."""
    from ciaosr_b200 import synth
    x = synth.synth_lr_image(1, 64, 64, 3).to(dev)
    with torch.no_grad():
        for _ in range(3):
            std = float(gen.gen_feature(x)[0].std())
            k = target_std / std
            for conv in (gen.conv_first, gen.conv_after_body):
                conv.weight.mul_(k)
                conv.bias.mul_(k)
        return float(gen.gen_feature(x)[0].std())


def prepare(case, dev, cuda_graph=True):
    """Build one case's restorer with synthetic weights on `dev`: (model, test generator, raw LR frame, coord/cell
    kwargs for un-tiled cases, test_cfg, synthetic feature std for the calibrated SwinIR trunks)."""
    from ciaosr_b200 import synth
    from ciaosr_b200.builder import build
    from ciaosr_b200.coords import make_cell, make_coord
    cfg, test_cfg = case["build"]()
    m = build(cfg, test_cfg=test_cfg)
    synth.fill_module(m.generator, 0)
    if getattr(m, "generator_ema", None) is not None:
        m.generator_ema.load_state_dict(m.generator.state_dict())
    m = m.eval().to(dev)
    gen = m._test_generator()
    feat_std = calibrate_features(gen, dev) if case["config"] in (4, 5) else None
    gen.cuda_graph = cuda_graph
    h, w, s = case["h"], case["w"], case["scale"]
    lq = (synth.synth_lr_image(1, h, w, 7) + torch.tensor(RGB_MEAN).view(1, 3, 1, 1)).to(dev)
    kw = {}
    if not test_cfg.get("tile"):
        kw = dict(coord=make_coord((h * s, w * s)).unsqueeze(0).to(dev),
                  cell=make_cell((h * s, w * s), h * s * w * s).unsqueeze(0).to(dev))
    return m, gen, lq, kw, test_cfg, feat_std


def run_frame(model, x, kw, world=1):
    """forward_test minus its final .cpu(): the blended, de-normalised, clamped frame [1, Ho*Wo, 3] on the device.
    With a process group, tiles (tiled cases) or bands of the coordinate list (un-tiled) are sharded over the ranks
    and assembled by one all-gather (ciaosr_b200/dist.py)."""
    x = (x - model.lq_mean.to(x)) / model.lq_std.to(x)
    model.gt_mean, model.gt_std = model.gt_mean.to(x), model.gt_std.to(x)
    g = model._test_generator()
    with torch.no_grad():
        if model.test_cfg.get("tile"):
            return model.clip_test(x, g, denorm=True)
        if world > 1:
            from ciaosr_b200 import dist as cdist
            p = cdist.sharded_query_forward(g, x, kw["coord"], kw["cell"], g.eval_bsize)
        else:
            p = g(x, kw["coord"], kw["cell"], test_mode=True)
        return (p * model.gt_std + model.gt_mean).clamp_(0, 1)


def run_alone(model, x, kw):
    """The same frame computed by this rank alone (no sharding): the reference for the bit-equality check."""
    gen = model._test_generator()
    x = (x - model.lq_mean.to(x)) / model.lq_std.to(x)
    with torch.no_grad():
        if model.test_cfg.get("tile"):
            model.test_cfg["shard_tiles"] = False
            try:
                return model.clip_test(x, gen, denorm=True)
            finally:
                model.test_cfg["shard_tiles"] = True
        return (gen(x, kw["coord"], kw["cell"], test_mode=True) * model.gt_std + model.gt_mean).clamp_(0, 1)


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--stages", type=int, default=1, help="add per-stage times of one eager pass (1 GPU)")
    ap.add_argument("--check-crop", type=int, default=0,
                    help="tiled cases: run the fp32-engine comparison on a crop of this many LR rows/cols "
                         "(0 = 2x2 tiles worth); the timed run is always the full frame")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    args = ap.parse_args()
    from ciaosr_b200 import synth
    from ciaosr_b200.builder import build
    from ciaosr_b200.coords import make_cell, make_coord
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    lines = []
    for case in cases({int(c) for c in args.configs.split(",")}):
        m, gen, lq, kw, test_cfg, feat_std = prepare(case, dev)
        h, w, s = case["h"], case["w"], case["scale"]

        def run(model=m, x=lq, kw=kw):
            return run_frame(model, x, kw, world)

        torch.cuda.reset_peak_memory_stats(dev)
        out = run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t))
        ms = min(times)
        npx = h * s * w * s
        assert out.shape == (1, npx, 3), out.shape
        line = dict(config=case["config"], case=case["name"], n_gpus=world, hr_px=npx, ms=ms,
                    mpix_s=npx / ms / 1e3, finite=bool(torch.isfinite(out).all()),
                    min=float(out.min()), max=float(out.max()),
                    peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                    synthetic_feature_std=feat_std, engine="tcgen05" if gen.head_plan().engine_supported("tcgen05") else "simt",
                    tiles=(len(m.tile_origins(h, min(test_cfg["tile"], h, w), test_cfg["tile_overlap"])) *
                           len(m.tile_origins(w, min(test_cfg["tile"], h, w), test_cfg["tile_overlap"]))
                           if test_cfg.get("tile") else 1))
        if world > 1 and args.check:
            # the sharded frame must equal the frame this rank computes alone (no collective inside)
            line["check_sharded_vs_single_max_abs"] = float((run_alone(m, lq, kw) - out).abs().max())
        if args.stages and world == 1:
            # one eager pass with the library's per-stage CUDA events; "encoder+glue" is the remainder
            from ciaosr_b200 import native
            gen.cuda_graph = False
            run()
            torch.cuda.synchronize()
            native.profile_read()
            native.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            native.profile_enable(False)
            st = {k: round(v[0], 3) for k, v in native.profile_read().items() if v[1]}
            st["encoder+glue (eager)"] = round(e0.elapsed_time(e1) - sum(st.values()), 3)
            line["stage_ms_per_frame"] = st
            gen.cuda_graph = True
        if args.check and rank == 0 and world == 1:
            # product path vs the fp32 CUDA-core engine (+ PyTorch fp32 encoder) on the same input
            if test_cfg.get("tile"):
                t = test_cfg["tile"]
                n = args.check_crop or (2 * t - test_cfg["tile_overlap"])
                x = lq[..., :min(n, h), :min(n, w)].contiguous()
                kc = {}
            else:
                n = args.check_crop or 96             # un-tiled: a crop keeps the fp32 engine's run short
                x = lq[..., :n, :n].contiguous()
                kc = dict(coord=make_coord((n * s, n * s)).unsqueeze(0).to(dev),
                          cell=make_cell((n * s, n * s), n * s * n * s).unsqueeze(0).to(dev))
            a = run(m, x, kc)
            gen.engine, gen.cuda_graph = "simt", False
            if hasattr(gen, "native_encoder"):
                gen.native_encoder = False
            t0 = time.time()
            b = run(m, x, kc)
            torch.cuda.synchronize()
            line.update(check_lr=list(x.shape[-2:]), check_max_abs=float((a - b).abs().max()),
                        check_psnr_db=psnr(a, b), check_fp32_engine_s=time.time() - t0)
        if rank == 0:
            print(json.dumps(line), flush=True)
            lines.append(line)
        del m, gen, out
        torch.cuda.empty_cache()
    if rank == 0:
        with open(args.out, "a") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
