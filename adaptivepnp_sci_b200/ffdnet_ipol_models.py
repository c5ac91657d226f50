"""FFDNet, IPOL flavour, as used by the reference's frame-wise gray adapter: ``packages/ffdnet/models.py:17-110``
(model) and ``packages/ffdnet/functions.py:16-53,55-100`` (its first / last layer).

Parameter containers with the reference's ``state_dict`` key names (``intermediate_dncnn.itermediate_dncnn.{i}.weight``,
BatchNorm entries at the inner layers), so ``net_gray.pth`` / ``net_rgb.pth`` of the public FFDNet-pytorch code load
unchanged.  Differences from the KAIR flavour the solvers use (``network_ffdnet.py``): no conv bias, BatchNorm between
the inner layers, the noise map comes FIRST in the down-sampled stack (one copy per colour plane) and the network
returns the NOISE estimate.  ``forward`` runs on the native engine (``engine.FFDNetEngine``, inference only: the
reference never fine-tunes this model): the engine's input kernel builds [sub-images | one noise plane], so the first
convolution is handed to it with its input columns in that order (the C noise columns, which all see the same value,
added up) - a re-ordering of an exact sum.
"""
import torch
import torch.nn as nn


class IntermediateDnCNN(nn.Module):
    """models.py:27-68"""

    def __init__(self, input_features, middle_features, num_conv_layers):
        super().__init__()
        if input_features not in (5, 15):
            raise Exception('Invalid number of input features')
        self.input_features, self.middle_features, self.num_conv_layers = input_features, middle_features, num_conv_layers
        self.output_features = 4 if input_features == 5 else 12
        layers = [nn.Conv2d(input_features, middle_features, 3, padding=1, bias=False), nn.ReLU(inplace=True)]
        for _ in range(num_conv_layers - 2):
            layers += [nn.Conv2d(middle_features, middle_features, 3, padding=1, bias=False),
                       nn.BatchNorm2d(middle_features), nn.ReLU(inplace=True)]
        layers.append(nn.Conv2d(middle_features, self.output_features, 3, padding=1, bias=False))
        self.itermediate_dncnn = nn.Sequential(*layers)      # (sic) the reference's attribute name is part of the key names


class FFDNet(nn.Module):
    """models.py:70-110: gray = 15 layers x 64 features, colour = 12 layers x 96 features."""

    def __init__(self, num_input_channels):
        super().__init__()
        if num_input_channels not in (1, 3):
            raise Exception('Invalid number of input features')
        self.num_input_channels = num_input_channels
        gray = num_input_channels == 1
        self.num_feature_maps, self.num_conv_layers = (64, 15) if gray else (96, 12)
        self.downsampled_channels, self.output_features = (5, 4) if gray else (15, 12)
        self.intermediate_dncnn = IntermediateDnCNN(self.downsampled_channels, self.num_feature_maps, self.num_conv_layers)
        self.in_nc = self.out_nc = num_input_channels
        self._engine = None
        self._first = None           # (holder conv, version of the real first conv it was derived from)

    def _first_conv(self):
        """First convolution with its input columns as the engine's input kernel lays them out: [4C sub-images | sigma]."""
        real = self.intermediate_dncnn.itermediate_dncnn[0]
        C = self.num_input_channels
        key = (real.weight._version, real.weight.data_ptr())
        if self._first is None or self._first[0].weight.device != real.weight.device:
            holder = nn.Conv2d(4 * C + 1, self.num_feature_maps, 3, padding=1, bias=False).to(real.weight.device)
            holder.weight.requires_grad_(False)
            self._first = [holder, None]
        if self._first[1] != key:
            w = real.weight.data
            self._first[0].weight.data.copy_(torch.cat((w[:, C:], w[:, :C].sum(1, keepdim=True)), 1))
            self._first[1] = key
            if self._engine is not None:
                self._engine.dirty = True
        return self._first[0]

    def conv_layers(self):
        seq = list(self.intermediate_dncnn.itermediate_dncnn)
        out = []
        for i, m in enumerate(seq):
            if isinstance(m, nn.Conv2d):
                bn = seq[i + 1] if i + 1 < len(seq) and isinstance(seq[i + 1], nn.BatchNorm2d) else None
                out.append((self._first_conv() if i == 0 else m, bn))
        return out

    def engine(self):
        from .engine import FFDNetEngine
        if self.training:
            raise NotImplementedError("IPOL FFDNet on the native engine is inference only (BatchNorm uses its running "
                                      "statistics; the reference's adapter calls model.eval(), test_ffdnet_ipol.py:128)")
        self._first_conv()
        if self._engine is None:
            self._engine = FFDNetEngine(self)
        return self._engine

    def forward(self, x, noise_sigma):
        """x [N,C,H,W] (even H, W), noise_sigma: one level (models.py:100-110) -> predicted noise [N,C,H,W]."""
        if noise_sigma.numel() > 1 and not bool((noise_sigma == noise_sigma.flatten()[0]).all()):
            raise NotImplementedError("one noise level per call on the native path")
        return self.engine().forward(x.contiguous().float(), float(noise_sigma.flatten()[0]), train=False).clone()
