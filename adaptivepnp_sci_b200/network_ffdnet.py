"""FFDNet (KAIR flavour) as used by the reference: ``models/network_ffdnet.py:27-69``.

The module is a PARAMETER CONTAINER with the reference's ``state_dict`` key names
(``model.{0,2,...}.weight/.bias``), so ``model_zoo/ffdnet_color.pth`` and
``ffdnet_gray.pth`` load unchanged.  ``forward(x, sigma)`` keeps the reference call
convention (``x`` [N,C,H,W], ``sigma`` [N,1,1,1], returns the DENOISED image) but runs
on the native sm_100a engine (``engine.FFDNetEngine``): pixel-unshuffle + sigma map ->
3x3 conv stack (implicit GEMM, bias+ReLU epilogues) -> pixel-shuffle.  No ATen conv is
ever called.
"""
import torch
import torch.nn as nn


class FFDNet(nn.Module):
    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=15, act_mode='R'):
        super().__init__()
        assert 'R' in act_mode or 'L' in act_mode, 'Examples of activation function: R, L, BR, BL, IR, IL'
        if act_mode != 'R':
            raise NotImplementedError("only act_mode='R' (conv+bias+ReLU) is on the hot path "
                                      "(two_stage_ADMM_Online_FFD_Warm.py:218-220)")
        self.in_nc, self.out_nc, self.nc, self.nb = in_nc, out_nc, nc, nb
        layers = [nn.Conv2d(in_nc * 4 + 1, nc, 3, 1, 1, bias=True), nn.ReLU(inplace=True)]
        for _ in range(nb - 2):
            layers += [nn.Conv2d(nc, nc, 3, 1, 1, bias=True), nn.ReLU(inplace=True)]
        layers += [nn.Conv2d(nc, out_nc * 4, 3, 1, 1, bias=True)]
        self.model = nn.Sequential(*layers)          # indices 0,2,4,... hold the convs, like B.sequential
        self._engine = None

    def conv_layers(self):
        return [m for m in self.model if isinstance(m, nn.Conv2d)]

    def engine(self):
        from .engine import FFDNetEngine
        if self._engine is None:
            self._engine = FFDNetEngine(self)
        return self._engine

    def forward(self, x, sigma):
        """x [N,in_nc,H,W] float32 CUDA, sigma [N,1,1,1] (one noise level per call on the hot path)."""
        return self.engine().forward_nchw(x, sigma)
