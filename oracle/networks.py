"""Oracle: CPU PyTorch restatements of the two denoiser networks.

Test infrastructure.  Module trees reproduce the reference ``state_dict`` key
names so the shipped weight files load unchanged; forward passes are plain
fp32 ATen ops (autograd-capable, used by the fine-tune oracle).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def pixel_unshuffle_kair(x, r=2):
    """models/basicblock.py:104-126 — output channel = c*r*r + dy*r + dx."""
    n, c, H, W = x.shape
    v = x.contiguous().view(n, c, H // r, r, W // r, r)
    return v.permute(0, 1, 3, 5, 2, 4).contiguous().view(n, c * r * r, H // r, W // r)


class FFDNet(nn.Module):
    """models/network_ffdnet.py:27-69 (KAIR FFDNet: returns the DENOISED image).
    Keys: model.{0,2,...}.weight / .bias."""

    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=15, act_mode="R"):
        super().__init__()
        layers = [nn.Conv2d(in_nc * 4 + 1, nc, 3, 1, 1, bias=True), nn.ReLU(inplace=True)]
        for _ in range(nb - 2):
            layers += [nn.Conv2d(nc, nc, 3, 1, 1, bias=True), nn.ReLU(inplace=True)]
        layers += [nn.Conv2d(nc, out_nc * 4, 3, 1, 1, bias=True)]
        self.model = nn.Sequential(*layers)

    def forward(self, x, sigma):
        h, w = x.shape[-2:]
        pb, pr = (-h) % 2, (-w) % 2
        x = F.pad(x, (0, pr, 0, pb), mode="replicate")                # :56-59
        x = pixel_unshuffle_kair(x, 2)                                 # :61
        m = sigma.repeat(1, 1, x.shape[-2], x.shape[-1])               # :63
        x = torch.cat((x, m), 1)                                       # :64
        x = self.model(x)
        x = F.pixel_shuffle(x, 2)                                      # :66
        return x[..., :h, :w]


def _cv(ci, co, stride=1, groups=1):
    return nn.Conv2d(ci, co, 3, padding=1, stride=stride, groups=groups, bias=False)


class _CvBlock(nn.Module):      # packages/fastdvdnet/models.py:16-30
    def __init__(self, ci, co):
        super().__init__()
        self.convblock = nn.Sequential(_cv(ci, co), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                                       _cv(co, co), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.convblock(x)


class _InputCvBlock(nn.Module):  # models.py:32-48
    def __init__(self, nfr, co, ncolor=3):
        super().__init__()
        self.convblock = nn.Sequential(_cv(nfr * (ncolor + 1), nfr * 30, groups=nfr), nn.BatchNorm2d(nfr * 30),
                                       nn.ReLU(inplace=True),
                                       _cv(nfr * 30, co), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.convblock(x)


class _DownBlock(nn.Module):     # models.py:50-62
    def __init__(self, ci, co):
        super().__init__()
        self.convblock = nn.Sequential(_cv(ci, co, stride=2), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                                       _CvBlock(co, co))

    def forward(self, x):
        return self.convblock(x)


class _UpBlock(nn.Module):       # models.py:64-75
    def __init__(self, ci, co):
        super().__init__()
        self.convblock = nn.Sequential(_CvBlock(ci, ci), _cv(ci, co * 4), nn.PixelShuffle(2))

    def forward(self, x):
        return self.convblock(x)


class _OutputCvBlock(nn.Module):  # models.py:77-89
    def __init__(self, ci, co):
        super().__init__()
        self.convblock = nn.Sequential(_cv(ci, ci), nn.BatchNorm2d(ci), nn.ReLU(inplace=True), _cv(ci, co))

    def forward(self, x):
        return self.convblock(x)


class DenBlock(nn.Module):       # models.py:146-198
    def __init__(self, num_input_frames=3, ncolor=3):
        super().__init__()
        self.inc = _InputCvBlock(num_input_frames, 32, ncolor)
        self.downc0 = _DownBlock(32, 64)
        self.downc1 = _DownBlock(64, 128)
        self.upc2 = _UpBlock(128, 64)
        self.upc1 = _UpBlock(64, 32)
        self.outc = _OutputCvBlock(32, ncolor)

    def forward(self, in0, in1, in2, noise_map):
        x0 = self.inc(torch.cat((in0, noise_map, in1, noise_map, in2, noise_map), dim=1))
        x1 = self.downc0(x0)
        x2 = self.downc1(x1)
        x2 = self.upc2(x2)
        x1 = self.upc1(x1 + x2)
        x = self.outc(x0 + x1)
        return in1 - x


class FastDVDnet(nn.Module):     # models.py:200-253
    def __init__(self, num_input_frames=5, num_color_channels=3):
        super().__init__()
        self.num_input_frames = num_input_frames
        self.num_color_channels = num_color_channels
        self.temp1 = DenBlock(3, num_color_channels)
        self.temp2 = DenBlock(3, num_color_channels)

    def forward(self, x, noise_map):
        C = self.num_color_channels
        x0, x1, x2, x3, x4 = (x[:, m * C:m * C + C] for m in range(self.num_input_frames))
        x20 = self.temp1(x0, x1, x2, noise_map)
        x21 = self.temp1(x1, x2, x3, noise_map)
        x22 = self.temp1(x2, x3, x4, noise_map)
        return self.temp2(x20, x21, x22, noise_map)


class Wrapped(nn.Module):
    """Stand-in for the ``nn.DataParallel`` wrapper the FastDVDnet script uses
    (two_stage_ADMM_Online_FastDVD_Warm.py:240-241): exposes ``.module`` and
    prefixes state-dict keys with ``module.``; batch-1 forwards never split."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a):
        return self.module(*a)
