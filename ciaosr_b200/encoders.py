"""Encoders (host PyTorch, per BASELINE.json north_star: "host code stays
Python/PyTorch for the EDSR/RDN/SwinIR encoder").

mmedit 0.11's EDSR / RDN are not vendored by the reference; these expose the
attributes the generators hoist (ciaosr_net.py:314-318, 388-390) with the
upstream parameter names, so released checkpoints map onto
``generator.{sfe1,sfe2,rdbs.N.layers.M.conv,rdbs.N.lff,gff.*}`` /
``generator.{conv_first,body.N.conv1|conv2,conv_after_body}`` unchanged.
"""
import torch
import torch.nn as nn


class ResidualBlockNoBN(nn.Module):
    def __init__(self, mid_channels=64, res_scale=1.0):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return x + self.conv2(self.relu(self.conv1(x))) * self.res_scale


class EDSR(nn.Module):
    """Trunk of EDSR; the upsampler / conv_last / mean shift that upstream also owns
    are dropped by the generator (``del self.encoder`` at ciaosr_net.py:391), so they
    are not built here."""

    def __init__(self, in_channels=3, out_channels=3, mid_channels=64, num_blocks=16,
                 upscale_factor=4, res_scale=1, rgb_mean=(0.4488, 0.4371, 0.4040),
                 rgb_std=(1.0, 1.0, 1.0)):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mid_channels, self.num_blocks = mid_channels, num_blocks
        self.conv_first = nn.Conv2d(in_channels, mid_channels, 3, padding=1)
        self.body = nn.Sequential(*[ResidualBlockNoBN(mid_channels, res_scale)
                                    for _ in range(num_blocks)])
        self.conv_after_body = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1)


class DenseLayer(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=3 // 2)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return torch.cat([x, self.relu(self.conv(x))], 1)


class RDB(nn.Module):
    def __init__(self, in_channels, channel_growth, num_layers):
        super().__init__()
        self.layers = nn.Sequential(*[DenseLayer(in_channels + channel_growth * i, channel_growth)
                                      for i in range(num_layers)])
        self.lff = nn.Conv2d(in_channels + channel_growth * num_layers, channel_growth, kernel_size=1)

    def forward(self, x):
        return x + self.lff(self.layers(x))


class RDN(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=64, num_blocks=16, upscale_factor=4,
                 num_layers=8, channel_growth=64):
        super().__init__()
        self.mid_channels, self.channel_growth = mid_channels, channel_growth
        self.num_blocks, self.num_layers = num_blocks, num_layers
        self.sfe1 = nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=3 // 2)
        self.sfe2 = nn.Conv2d(mid_channels, mid_channels, kernel_size=3, padding=3 // 2)
        self.rdbs = nn.ModuleList([RDB(mid_channels, channel_growth, num_layers)])
        for _ in range(num_blocks - 1):
            self.rdbs.append(RDB(channel_growth, channel_growth, num_layers))
        self.gff = nn.Sequential(
            nn.Conv2d(channel_growth * num_blocks, mid_channels, kernel_size=1),
            nn.Conv2d(mid_channels, mid_channels, kernel_size=3, padding=3 // 2))
