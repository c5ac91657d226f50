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


class FFDNetIPOL(nn.Module):
    """packages/ffdnet/models.py:70-110 + functions.py:16-53 (first layer) / :55-81 (last layer): the IPOL flavour used by
    the frame-wise gray adapter.  No conv bias, BatchNorm between the inner layers, returns the NOISE estimate.
    Keys: intermediate_dncnn.itermediate_dncnn.{i}.weight (sic) + BatchNorm entries."""

    def __init__(self, num_input_channels):
        super().__init__()
        gray = num_input_channels == 1
        nf, nl, cin, cout = (64, 15, 5, 4) if gray else (96, 12, 15, 12)            # models.py:76-88
        layers = [nn.Conv2d(cin, nf, 3, padding=1, bias=False), nn.ReLU(inplace=True)]
        for _ in range(nl - 2):
            layers += [nn.Conv2d(nf, nf, 3, padding=1, bias=False), nn.BatchNorm2d(nf), nn.ReLU(inplace=True)]
        layers.append(nn.Conv2d(nf, cout, 3, padding=1, bias=False))
        self.intermediate_dncnn = nn.Module()
        self.intermediate_dncnn.itermediate_dncnn = nn.Sequential(*layers)
        self.num_input_channels = num_input_channels

    def forward(self, x, noise_sigma):
        N, C, H, W = x.shape
        down = torch.zeros((N, 4 * C, H // 2, W // 2), dtype=x.dtype)
        for idx, (i, j) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):              # functions.py:36,48-50
            down[:, idx::4] = x[:, :, i::2, j::2]
        noise_map = noise_sigma.view(N, 1, 1, 1).repeat(1, C, H // 2, W // 2)         # :45
        h = self.intermediate_dncnn.itermediate_dncnn(torch.cat((noise_map, down), 1))  # :53: the noise planes come FIRST
        out = torch.zeros((N, C, H, W), dtype=x.dtype)
        for idx, (i, j) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):              # :76-79
            out[:, :, i::2, j::2] = h[:, idx::4]
        return out


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


# ---------------------------------------------------------------------------------------------------
# DDnet deep demosaicker (models/network_demosaicking.py).  No BatchNorm, no bias, base width 20.
# ---------------------------------------------------------------------------------------------------
def _seq_cv2(ci, co):            # network_demosaicking.py:34-46 (CvBlock): conv,ReLU,conv,ReLU
    return nn.Sequential(_cv(ci, co), nn.ReLU(inplace=True), _cv(co, co), nn.ReLU(inplace=True))


class _Blk(nn.Module):
    """Thin holder so that parameters appear under ``<name>.convblock.<i>`` as in the reference."""

    def __init__(self, seq):
        super().__init__()
        self.convblock = seq

    def forward(self, x):
        return self.convblock(x)


def _dd_input(nfr, per_frame, co):   # :48-81 InputCvBlock / InputCvBlock_2
    return _Blk(nn.Sequential(_cv(nfr * per_frame, nfr * 30, groups=nfr), nn.ReLU(inplace=True),
                              _cv(nfr * 30, co), nn.ReLU(inplace=True)))


def _dd_down(ci, co):                # :83-95
    return _Blk(nn.Sequential(_cv(ci, co, stride=2), nn.ReLU(inplace=True), _Blk(_seq_cv2(co, co))))


def _dd_up(ci, co):                  # :97-109
    return _Blk(nn.Sequential(_Blk(_seq_cv2(ci, ci)), _cv(ci, co * 4), nn.PixelShuffle(2)))


def _dd_out(ci, co):                 # :111-123
    return _Blk(nn.Sequential(_cv(ci, ci), nn.ReLU(inplace=True), _cv(ci, co)))


class DDDenBlock(nn.Module):
    """network_demosaicking.py:184-246 (``DenBlock``) and :310-375 (``DenBlock4ChBayer``, ``bayer4=True``).
    The noise-map input block ``inc`` exists (it is in the state_dict) but DDnet never uses it (:431-440)."""

    def __init__(self, ch_each_frame=3, bayer4=False):
        super().__init__()
        c0, c1, c2 = 20, 40, 80                                        # base_layer = 20 (:22)
        self.inc = _dd_input(3, 3 + 1, c0)
        self.inc_1 = _dd_input(3, ch_each_frame, c0)
        self.downc0 = _dd_down(c0, c1)
        self.downc1 = _dd_down(c1, c2)
        self.upc2 = _dd_up(c2, c1)
        self.upc1 = _dd_up(c1, c0)
        self.outc = _dd_out(c0, 4 if bayer4 else 3)
        self.bayer4 = bayer4
        if bayer4:
            self.upscale = nn.UpsamplingBilinear2d(scale_factor=2)     # align_corners=True
            self.fusion = _dd_out(4, 3)

    def forward(self, in0, in1, in2):
        x0 = self.inc_1(torch.cat((in0, in1, in2), dim=1))
        x1 = self.downc0(x0)
        x2 = self.downc1(x1)
        x2 = self.upc2(x2)
        x1 = self.upc1(x1 + x2)
        x = self.outc(x0 + x1)
        x = in1 + x                                                    # :242 (1-channel in1 broadcasts over 3)
        if self.bayer4:
            x = self.fusion(self.upscale(x))                           # :371-372
        return x


def _mosaic_to_4ch(m):
    """[N,H,W] -> [N,4,h,w], channel order (0,0),(0,1),(1,0),(1,1) (utils_image.py:145-151 via :420-424)."""
    return torch.stack((m[:, 0::2, 0::2], m[:, 0::2, 1::2], m[:, 1::2, 0::2], m[:, 1::2, 1::2]), dim=1)


class DDnet(nn.Module):
    """network_demosaicking.py:377-463.  Input [N, 5*3, H, W]: five sparse-RGB (one colour per site) frames."""

    def __init__(self, num_input_frames=5):
        super().__init__()
        self.num_input_frames = num_input_frames
        self.temp1 = DDDenBlock(ch_each_frame=1)
        self.temp2 = DDDenBlock(ch_each_frame=3)
        self.temp11 = DDDenBlock(ch_each_frame=4, bayer4=True)
        self.weight_tensor_in = nn.Parameter(torch.ones((9, 1, 1, 1, 1)))
        self.weight_tensor_in2 = nn.Parameter(torch.ones((9, 1, 4, 1, 1)))
        self.weight_tensor_out = nn.Parameter(torch.ones((2, 1, 3, 1, 1)))

    def forward(self, x, noise_map=None):
        fr = [x[:, 3 * m:3 * m + 3].sum(dim=1) for m in range(self.num_input_frames)]       # :411-416, [N,H,W]
        f4 = [_mosaic_to_4ch(f) for f in fr]
        f1 = [f.unsqueeze(1) for f in fr]
        a, a2, a3 = self.weight_tensor_in, self.weight_tensor_in2, self.weight_tensor_out
        y1 = [self.temp1(f1[j] * a[3 * j], f1[j + 1] * a[3 * j + 1], f1[j + 2] * a[3 * j + 2]) for j in range(3)]
        y2 = [self.temp11(f4[j] * a2[3 * j], f4[j + 1] * a2[3 * j + 1], f4[j + 2] * a2[3 * j + 2]) for j in range(3)]
        return a3[0] * self.temp2(*y1) + a3[1] * self.temp2(*y2)                              # :452-462
