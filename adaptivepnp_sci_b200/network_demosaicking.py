"""DDnet deep demosaicker as used by the reference: ``models/network_demosaicking.py:377-463``.

Parameter containers with the reference's ``state_dict`` key names (``temp1.inc_1.convblock.0.weight`` ...,
``weight_tensor_in`` ...), so ``model_zoo/ddnet1.pth`` (a ``{'state_dict': ...}`` checkpoint of the
``nn.DataParallel``-wrapped model, two_stage_ADMM_Online_FFD_Warm.py:229-232) loads unchanged through
``fastdvdnet_adapter.DataParallelLike``.  All convolutions are 3x3, pad 1, **no bias, no BatchNorm** (base width 20).
``forward`` keeps the reference convention (x [1,15,H,W]: five sparse-RGB frames) and runs on the native engine; the
hot path calls ``engine.DDnetEngine`` on the whole circular frame sequence.
"""
import torch
import torch.nn as nn

BASE = 20          # network_demosaicking.py:22


def _cv(ci, co, stride=1, groups=1):
    return nn.Conv2d(ci, co, kernel_size=3, padding=1, stride=stride, groups=groups, bias=False)


class _Holder(nn.Module):
    """Gives a layer list the ``<name>.convblock.<i>`` key prefix of the reference blocks."""

    def __init__(self, *mods):
        super().__init__()
        self.convblock = nn.Sequential(*mods)


def _two(ci, co):                   # CvBlock :34-46
    return _Holder(_cv(ci, co), nn.ReLU(inplace=True), _cv(co, co), nn.ReLU(inplace=True))


def _input(nfr, per_frame, co):     # InputCvBlock / InputCvBlock_2 :48-81
    return _Holder(_cv(nfr * per_frame, nfr * 30, groups=nfr), nn.ReLU(inplace=True), _cv(nfr * 30, co),
                   nn.ReLU(inplace=True))


def _down(ci, co):                  # DownBlock :83-95
    return _Holder(_cv(ci, co, stride=2), nn.ReLU(inplace=True), _two(co, co))


def _up(ci, co):                    # UpBlock :97-109
    return _Holder(_two(ci, ci), _cv(ci, co * 4), nn.PixelShuffle(2))


def _out(ci, co):                   # OutputCvBlock :111-123
    return _Holder(_cv(ci, ci), nn.ReLU(inplace=True), _cv(ci, co))


class DenBlock(nn.Module):
    """:184-246 (and, with ``bayer4=True``, ``DenBlock4ChBayer`` :310-375: 4-channel half-resolution frames, bilinear x2
    up-sampling and a 4->4->3 ``fusion`` block after the residual).  ``inc`` (the noise-map input block) is part of the
    state_dict but never executed by DDnet."""

    def __init__(self, num_input_frames=3, ch_each_frame=3, bayer4=False):
        super().__init__()
        c0, c1, c2 = BASE, 2 * BASE, 4 * BASE
        self.chs_lyr0, self.chs_lyr1, self.chs_lyr2 = c0, c1, c2
        self.ch_each_frame, self.bayer4 = ch_each_frame, bayer4
        self.inc = _input(num_input_frames, 3 + 1, c0)
        self.inc_1 = _input(num_input_frames, ch_each_frame, c0)
        self.downc0 = _down(c0, c1)
        self.downc1 = _down(c1, c2)
        self.upc2 = _up(c2, c1)
        self.upc1 = _up(c1, c0)
        self.outc = _out(c0, 4 if bayer4 else 3)
        if bayer4:
            self.upscale = nn.UpsamplingBilinear2d(scale_factor=2)
            self.fusion = _out(4, 3)

    def conv_specs(self):
        """The 16 convolutions of the U-shaped body in execution order: (conv, relu, stride, pixel_shuffle)."""
        def pair(h):
            s = h.convblock
            return [(s[0], True, 1, False), (s[2], True, 1, False)]
        i1, d0, d1, u2, u1, oc = (self.inc_1.convblock, self.downc0.convblock, self.downc1.convblock,
                                  self.upc2.convblock, self.upc1.convblock, self.outc.convblock)
        return ([(i1[0], True, 1, False), (i1[2], True, 1, False)]
                + [(d0[0], True, 2, False)] + pair(d0[2])
                + [(d1[0], True, 2, False)] + pair(d1[2])
                + pair(u2[0]) + [(u2[1], False, 1, True)]
                + pair(u1[0]) + [(u1[1], False, 1, True)]
                + [(oc[0], True, 1, False), (oc[2], False, 1, False)])

    def fusion_specs(self):
        f = self.fusion.convblock
        return [(f[0], True, 1, False), (f[2], False, 1, False)]


class DDnet(nn.Module):
    def __init__(self, num_input_frames=5):
        super().__init__()
        if num_input_frames != 5:
            raise NotImplementedError("DDnet is a 5-frame model (NUM_IN_FR_EXT = 5, DDnet_test.py:16)")
        self.num_input_frames = num_input_frames
        self.temp1 = DenBlock(3, ch_each_frame=1)
        self.temp2 = DenBlock(3, ch_each_frame=3)
        self.temp11 = DenBlock(3, ch_each_frame=4, bayer4=True)
        self.weight_tensor_in = nn.Parameter(torch.ones((9, 1, 1, 1, 1)))       # :398-400
        self.weight_tensor_in2 = nn.Parameter(torch.ones((9, 1, 4, 1, 1)))
        self.weight_tensor_out = nn.Parameter(torch.ones((2, 1, 3, 1, 1)))
        for m in self.modules():                                                 # :402-409
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, nonlinearity='relu')
        self._engine = None

    def engine(self):
        from .engine import DDnetEngine
        if self._engine is None:
            self._engine = DDnetEngine(self)
        return self._engine

    def forward(self, x, noise_map=None):
        """x [1,15,H,W] (5 sparse-RGB frames stacked frame-major) -> [1,3,H,W] (the centre frame, demosaicked)."""
        return self.engine().forward_window(x)
