"""Restorer (mmedit "model") level of the boundary: ``CiaoSR`` as the reference's
configs build it (mmedited/models/restorers/ciaosr.py:17-258 on top of
basic_restorer.py:16-124).  Inference side only: ``forward(test_mode=True)``,
``forward_test``, ``clip_test``, ``evaluate``.  The generator call inside is the
drop-in boundary; the tile blend + de-normalise + clamp epilogue runs in two small
native kernels instead of the reference's Python double loop over full-frame masks.
"""
import copy
import math
import numbers
import os.path as osp

import torch
import torch.nn as nn

from . import dist as cdist
from . import native
from .builder import build_backbone, build_loss
from .coords import make_coord
from .metrics import psnr, psnr_device, ssim, ssim_device, tensor2img


def _to_host(tensors):
    """``{k: v.cpu()}`` (ciaosr.py:181-183) with ONE synchronisation: every CUDA tensor is copied into a fresh
    page-locked host tensor (PyTorch's caching host allocator recycles the blocks) with non-blocking copies on
    the current stream.  A pageable ``.cpu()`` stages through a bounce buffer at ~1/4 of the PCIe rate and
    synchronises per tensor."""
    out, dev = {}, None
    for k, v in tensors.items():
        if v.is_cuda:
            h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            h.copy_(v, non_blocking=True)
            out[k], dev = h, v.device
        else:
            out[k] = v
    if dev is not None:
        torch.cuda.current_stream(dev).synchronize()
    return out


class BasicRestorer(nn.Module):
    allowed_metrics = {"PSNR": psnr, "SSIM": ssim}

    def __init__(self, generator, pixel_loss, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.fp16_enabled = False
        self.generator = build_backbone(generator)
        self.init_weights(pretrained)
        self.pixel_loss = build_loss(pixel_loss)

    def init_weights(self, pretrained=None):
        self.generator.init_weights(pretrained)

    def forward(self, lq, gt=None, test_mode=False, **kwargs):
        if test_mode:
            return self.forward_test(lq, gt, **kwargs)
        return self.forward_train(lq, gt, **kwargs)

    def forward_train(self, lq, gt, **kwargs):
        """basic_restorer.py:84-99 (the implicit generators also need coord / cell)."""
        output = self.generator(lq, **kwargs)
        return dict(losses=dict(loss_pix=self.pixel_loss(output, gt)), num_samples=len(gt.data),
                    results=dict(lq=lq.cpu(), gt=gt.cpu(), output=output.detach().cpu()))

    @staticmethod
    def parse_losses(losses):
        """mmedit BaseModel.parse_losses: sum every entry whose key contains 'loss'; log_vars as Python floats
        (averaged over the process group when one is initialised)."""
        import torch.distributed as tdist
        log_vars = {}
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        for name in log_vars:
            v = log_vars[name].detach().clone()
            if tdist.is_available() and tdist.is_initialized():
                tdist.all_reduce(v.div_(tdist.get_world_size()))
            log_vars[name] = v.item()
        return loss, log_vars

    def evaluate(self, output, gt):
        """basic_restorer.py:101-124: metrics on uint8 images with crop_border / convert_to."""
        crop_border = self.test_cfg["crop_border"]
        convert_to = self.test_cfg.get("convert_to", None)
        if output.is_cuda and gt.is_cuda and output.shape[0] == 1:
            # same numbers without moving the frames to the host (metrics.py, "evaluated where the image is")
            fns = {"PSNR": psnr_device, "SSIM": ssim_device}
            return {m: fns[m](output, gt, crop_border, convert_to=convert_to) for m in self.test_cfg["metrics"]}
        output, gt = tensor2img(output), tensor2img(gt)
        return {m: self.allowed_metrics[m](output, gt, crop_border, convert_to=convert_to)
                for m in self.test_cfg["metrics"]}


class CiaoSR(BasicRestorer):
    def __init__(self, generator, pixel_loss, rgb_mean=(0.5, 0.5, 0.5), rgb_std=(0.5, 0.5, 0.5),
                 train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__(generator, pixel_loss, train_cfg=train_cfg, test_cfg=test_cfg,
                         pretrained=pretrained)
        rgb_mean, rgb_std = torch.FloatTensor(rgb_mean), torch.FloatTensor(rgb_std)
        self.lq_mean, self.lq_std = rgb_mean.view(1, -1, 1, 1), rgb_std.view(1, -1, 1, 1)
        self.gt_mean, self.gt_std = rgb_mean.view(1, 1, -1), rgb_std.view(1, 1, -1)

    def train_step(self, data_batch, optimizer):
        """ciaosr.py:60-109: normalise, generator(lq, coord, cell), pixel loss, one optimizer step."""
        coord, cell, lq, gt = data_batch["coord"], data_batch["cell"], data_batch["lq"], data_batch["gt"]
        self.lq_mean, self.lq_std = self.lq_mean.to(lq), self.lq_std.to(lq)
        self.gt_mean, self.gt_std = self.gt_mean.to(gt), self.gt_std.to(gt)
        lq = (lq - self.lq_mean) / self.lq_std
        gt = (gt - self.gt_mean) / self.gt_std
        pred = self.generator(lq, coord, cell)
        loss, log_vars = self.parse_losses(dict(loss_pix=self.pixel_loss(pred, gt)))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        log_vars.pop("loss")
        return dict(log_vars=log_vars, num_samples=len(gt.data),
                    results=dict(lq=lq.cpu(), gt=gt.cpu(), output=pred.detach().cpu()))

    def forward_test(self, lq, gt=None, coord=None, cell=None, meta=None, save_image=False,
                     save_path=None, iteration=None):
        """ciaosr.py:111-203."""
        self.lq_mean, self.lq_std = self.lq_mean.to(lq), self.lq_std.to(lq)
        lq = (lq - self.lq_mean) / self.lq_std
        self.gt_mean, self.gt_std = self.gt_mean.to(lq), self.gt_std.to(lq)
        model = self._test_generator()
        with torch.no_grad():
            if self.test_cfg is not None and self.test_cfg.get("tile", None):
                pred = self.clip_test(lq, model, denorm=True)
            else:
                if cdist.world()[1] > 1 and lq.shape[0] == 1 and (self.test_cfg or {}).get("shard_queries", True):
                    # one process per GPU: bands of the coordinate list per rank, one all-gather (dist.py)
                    pred = cdist.sharded_query_forward(model, lq, coord, cell, getattr(model, "eval_bsize", None))
                else:
                    pred = model(lq, coord, cell, test_mode=True)
                pred = pred * self.gt_std + self.gt_mean
                pred.clamp_(0, 1)
        ih, iw = lq.shape[-2:]
        if coord is not None:
            s = math.sqrt(coord.shape[1] / (ih * iw))
        else:
            s = self.test_cfg["scale"]
        shape = [lq.shape[0], round(ih * s), round(iw * s), 3]
        pred = pred.view(*shape).permute(0, 3, 1, 2).contiguous()
        if gt is not None:
            gt = gt.view(*shape).permute(0, 3, 1, 2).contiguous()
        if self.test_cfg is not None and self.test_cfg.get("metrics", None):
            assert gt is not None, "evaluation with metrics must have gt images."
            results = dict(eval_result=self.evaluate(pred, gt))
        else:
            results = dict(lq=lq, output=pred)
            if gt is not None:
                results["gt"] = gt
            results = _to_host(results)
        if save_image:
            import cv2
            key = "gt_path" if "gt_path" in meta[0] else "lq_path"
            folder_name = osp.splitext(osp.basename(meta[0][key]))[0]
            if isinstance(iteration, numbers.Number):
                save_path = osp.join(save_path, folder_name, f"{folder_name}-{iteration + 1:06d}.png")
            elif iteration is None:
                save_path = osp.join(save_path, f"{folder_name}.png")
            else:
                raise ValueError(f"iteration should be number or None, but got {type(iteration)}")
            cv2.imwrite(save_path, tensor2img(pred))
        return results

    def init_weights(self, pretrained=None, strict=True):
        self.generator.init_weights(pretrained, strict)

    def _test_generator(self):
        return self.generator

    @staticmethod
    def tile_origins(n, tile, overlap):
        """Tile start offsets along one axis (ciaosr.py:227-229)."""
        stride = tile - overlap
        return list(range(0, n - tile, stride)) + [n - tile]

    def clip_test(self, img_lq, model, denorm=False):
        """Overlap-average tiled inference (ciaosr.py:218-258) -> [B, Ho*Wo, 3].

        With torch.distributed initialised (one process per GPU) the tiles are dealt round-robin
        to the ranks, each rank runs the generator on its tiles, ONE all-gather collects the tile
        predictions, and every rank blends the full frame locally (ciaosr_b200/dist.py).
        """
        sf = self.test_cfg.get("scale", None)
        b, c, h, w = img_lq.size()
        tile = min(self.test_cfg.get("tile", None), h, w)
        overlap = self.test_cfg.get("tile_overlap", None)
        origins = [(y0, x0) for y0 in self.tile_origins(h, tile, overlap)
                   for x0 in self.tile_origins(w, tile, overlap)]
        ho, wo = h * sf, w * sf
        acc = torch.zeros(b, c, ho, wo, dtype=torch.float32, device=img_lq.device)
        cnt = torch.zeros_like(acc)
        th, tw = round(tile * sf), round(tile * sf)
        hr_coord = make_coord((th, tw)).unsqueeze(0).expand(b, -1, 2).to(img_lq).contiguous()
        cell = torch.ones_like(hr_coord)
        cell[:, :, 0] *= 2 / th
        cell[:, :, 1] *= 2 / tw

        # Tiles are independent generator calls (ciaosr.py:233-245) of one shape, so several of them ride in the
        # batch dimension of ONE call: per-tile kernels of a 128x128 tile are too small to fill 148 SMs (the
        # SwinIR trunk's Linear layers have 128 row tiles), and batch items are independent, so the result is
        # the same.  test_cfg['tile_batch'] overrides the group size.
        group = int(self.test_cfg.get("tile_batch", 0)) or max(1, min(8, round(73728 / (tile * tile))))

        def run_tiles(batch):
            """[(y0, x0), ...] -> list of [b, th*tw, 3] predictions, one generator call."""
            t = len(batch)
            patch = torch.cat([img_lq[..., y0:y0 + tile, x0:x0 + tile] for y0, x0 in batch], dim=0).contiguous()
            out = model(patch, hr_coord.repeat(t, 1, 1) if t > 1 else hr_coord,
                        cell.repeat(t, 1, 1) if t > 1 else cell, test_mode=True)
            return list(out.split(b, dim=0))

        def run_many(tiles):
            preds = []
            for i in range(0, len(tiles), group):
                preds += run_tiles(tiles[i:i + group])
            return preds

        if cdist.world()[1] > 1 and self.test_cfg.get("shard_tiles", True):
            preds = cdist.sharded_tile_predictions(origins, None, (b, th * tw, 3), img_lq, run_many=run_many)
        else:
            preds = run_many(origins)
        for (y0, x0), out in zip(origins, preds):
            native.tile_blend_accumulate(out, acc, cnt, y0 * sf, x0 * sf, th, tw)
        if denorm:
            return native.tile_blend_finish(acc, cnt, self.gt_mean.to(acc), self.gt_std.to(acc), True)
        return native.tile_blend_finish(acc, cnt)


class RealCiaoSR(CiaoSR):
    """Real-world variant (mmedited/models/restorers/real_ciaosr.py:28-373, configs/002_*.py): same
    inference path as ``CiaoSR`` but evaluated through the EMA copy of the generator
    (``generator_ema``, real_ciaosr.py:84-87, 262) -- its weights are what the released real-world
    checkpoints are tested with, so both copies appear in ``state_dict`` under the reference's names
    together with the ``step_counter`` buffer.  Its ``clip_test`` (real_ciaosr.py:336-373) is the same
    overlap-average blend as ``CiaoSR.clip_test`` for batch 1 (it builds batch-1 coordinates with
    ``.cuda()``); the batched form above covers it.

    The GAN-training pieces (discriminator, GAN / perceptual losses, the sharpened-GT switches, the
    degradation queue in ``train_step``) are outside the inference scope: their configs are accepted and
    kept as attributes, never built.
    """

    def __init__(self, generator, pixel_loss=None, perceptual_loss=None, discriminator=None, gan_loss=None,
                 rgb_mean=(0.5, 0.5, 0.5), rgb_std=(0.5, 0.5, 0.5), train_cfg=None, test_cfg=None,
                 pretrained=None, is_use_sharpened_gt_in_pixel=False, is_use_sharpened_gt_in_percep=False,
                 is_use_sharpened_gt_in_gan=False, is_use_ema=True):
        super().__init__(generator, pixel_loss if pixel_loss is not None else dict(type="L1Loss"),
                         rgb_mean=rgb_mean, rgb_std=rgb_std, train_cfg=train_cfg, test_cfg=test_cfg,
                         pretrained=pretrained)
        self.perceptual_loss_cfg, self.discriminator_cfg, self.gan_loss_cfg = perceptual_loss, discriminator, gan_loss
        self.is_use_sharpened_gt_in_pixel = is_use_sharpened_gt_in_pixel
        self.is_use_sharpened_gt_in_percep = is_use_sharpened_gt_in_percep
        self.is_use_sharpened_gt_in_gan = is_use_sharpened_gt_in_gan
        self.is_use_ema = is_use_ema
        self.generator_ema = copy.deepcopy(self.generator) if is_use_ema else None
        self.register_buffer("step_counter", torch.zeros(1))
        self.start_iter = train_cfg.get("start_iter", -1) if train_cfg is not None else -1

    def _test_generator(self):
        return self.generator_ema if self.is_use_ema else self.generator

    def train_step(self, data_batch, optimizer):
        raise NotImplementedError("GAN training of RealCiaoSR is out of scope for ciaosr_b200 (SURVEY.md 8f #4)")
