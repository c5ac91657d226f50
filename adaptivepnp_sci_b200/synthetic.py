"""Deterministic synthetic inputs used by bench.py, smoke() and the scripts.

No dataset ships with the reference (readme.md:22-23) and the FastDVDnet
weights are absent (.MISSING_LARGE_BLOBS), so throughput is measured on
synthetic mask-modulated cubes and a deterministic contractive FastDVDnet init.
``oracle/synthetic.py`` holds the checker's own copy; ``tests/test_synthetic.py``
asserts the two agree bit for bit.
"""
import numpy as np
import torch


def make_case(H, W, B, seed, bayer=True):
    """Smooth moving colour video + random binary mask + noiseless snapshot.

    Returns float32 ``meas[H,W]``, ``mask[H,W,B]``, ``orig[H,W,B]`` with the
    scripts' ``/255`` already applied (ADMM_TV_Warm_Start_save.py:118-121).
    Zero-sum mask pixels occur with P=2^-B and exercise ``Phi_sum==0 -> 1``
    (dvp_linear_inv_2_stage_ADMM_tensor_online.py:73).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    orig = np.empty((H, W, B), np.float32)
    for t in range(B):
        rgb = []
        for c in range(3):
            f = 0.5 + 0.35 * np.sin((xx + 3 * t + 5 * c) / 17.0) * np.cos((yy - 2 * t + 3 * c) / 23.0)
            f = f + 0.05 * rng.standard_normal((H, W))
            rgb.append(np.clip(f, 0.0, 1.0).astype(np.float32))
        if bayer:
            fr = np.empty((H, W), np.float32)
            fr[0::2, 0::2] = rgb[0][0::2, 0::2]
            fr[0::2, 1::2] = rgb[1][0::2, 1::2]
            fr[1::2, 0::2] = rgb[1][1::2, 0::2]
            fr[1::2, 1::2] = rgb[2][1::2, 1::2]
        else:
            fr = rgb[0]
        orig[:, :, t] = fr
    mask = (rng.random((H, W, B)) > 0.5).astype(np.float32)
    meas = (orig * mask).sum(2).astype(np.float32)
    return meas, mask, orig


def fastdvdnet_synthetic_state_dict(seed=4242, out_gain=0.005):
    """Contractive random init for FastDVDnet (the trained weights are absent,
    .MISSING_LARGE_BLOBS:3-7).  Kaiming-normal convs as in
    packages/fastdvdnet/models.py:216-223, default BatchNorm statistics, and the
    last conv of each DenBlock scaled by ``out_gain`` so that ``in1 - net(.)``
    (models.py:196) stays close to the identity.
    Key order/names are those of ``FastDVDnet().state_dict()``.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci_per_group, gain=1.0):
        std = float(np.sqrt(2.0 / (ci_per_group * 9)))
        sd[name + ".weight"] = torch.randn(co, ci_per_group, 3, 3, generator=g) * (std * gain)

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.05 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.02 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.02 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 1.0 + 0.1 * torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def cvblock(prefix, ci, co):
        conv(prefix + ".convblock.0", co, ci); bn(prefix + ".convblock.1", co)
        conv(prefix + ".convblock.3", co, co); bn(prefix + ".convblock.4", co)

    for blk in ("temp1", "temp2"):
        conv(blk + ".inc.convblock.0", 90, 4); bn(blk + ".inc.convblock.1", 90)
        conv(blk + ".inc.convblock.3", 32, 90); bn(blk + ".inc.convblock.4", 32)
        for name, ci, co in (("downc0", 32, 64), ("downc1", 64, 128)):
            conv(f"{blk}.{name}.convblock.0", co, ci); bn(f"{blk}.{name}.convblock.1", co)
            cvblock(f"{blk}.{name}.convblock.3", co, co)
        for name, ci, co in (("upc2", 128, 64), ("upc1", 64, 32)):
            cvblock(f"{blk}.{name}.convblock.0", ci, ci)
            conv(f"{blk}.{name}.convblock.1", co * 4, ci)
        conv(blk + ".outc.convblock.0", 32, 32); bn(blk + ".outc.convblock.1", 32)
        conv(blk + ".outc.convblock.3", 3, 32, gain=out_gain)
    return sd


def fastdvdnet_gray_synthetic_state_dict(seed=4242, out_gain=0.005):
    """Single-channel FastDVDnet (num_color_channels=1, packages/fastdvdnet/models.py:34-48,77-91): the colour synthetic
    state with the first layer's [r, sigma] columns of every frame group and the first output filter."""
    sd = fastdvdnet_synthetic_state_dict(seed, out_gain)
    for blk in ("temp1", "temp2"):
        sd[blk + ".inc.convblock.0.weight"] = sd[blk + ".inc.convblock.0.weight"][:, [0, 3]].clone()
        sd[blk + ".outc.convblock.3.weight"] = sd[blk + ".outc.convblock.3.weight"][0:1].clone()
    return sd


def ffdnet_ipol_synthetic_state_dict(num_input_channels=1, seed=99, out_gain=0.05):
    """Random state for the IPOL-flavour FFDNet (packages/ffdnet/models.py:70-110; net_gray.pth is not shipped): Kaiming
    convs, non-trivial BatchNorm running statistics, a small last layer (the network predicts the noise)."""
    gray = num_input_channels == 1
    nf, nl, cin, cout = (64, 15, 5, 4) if gray else (96, 12, 15, 12)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = "intermediate_dncnn.itermediate_dncnn."
    idx = 0
    for layer in range(nl):
        ci, co = (cin if layer == 0 else nf), (cout if layer == nl - 1 else nf)
        gain = out_gain if layer == nl - 1 else 1.0
        sd[p + "%d.weight" % idx] = torch.randn(co, ci, 3, 3, generator=g) * (float(np.sqrt(2.0 / (ci * 9))) * gain)
        idx += 1
        if 0 < layer < nl - 1:
            sd[p + "%d.weight" % idx] = 1.0 + 0.05 * torch.randn(co, generator=g)
            sd[p + "%d.bias" % idx] = 0.02 * torch.randn(co, generator=g)
            sd[p + "%d.running_mean" % idx] = 0.02 * torch.randn(co, generator=g)
            sd[p + "%d.running_var" % idx] = 1.0 + 0.1 * torch.rand(co, generator=g)
            sd[p + "%d.num_batches_tracked" % idx] = torch.tensor(0, dtype=torch.long)
            idx += 1
        if layer < nl - 1:
            idx += 1                       # ReLU
    return sd


def ddnet_synthetic_state_dict(seed=777, out_gain=0.002):
    """Random init for DDnet (the trained ``model_zoo/ddnet1.pth`` is absent, .MISSING_LARGE_BLOBS).
    Kaiming-normal convs (models/network_demosaicking.py:402-405), the last conv of every DenBlock and of the
    ``fusion`` block scaled by ``out_gain`` so the residual form ``in1 + net(.)`` (:242) stays bounded, the mixing
    scalars perturbed off 1 so that every one of them matters, and the output mix set to about (1, 0.1): the result
    is roughly the mosaic broadcast to three channels plus a small learned-looking correction (bounded ADMM loop).
    Key order/names are those of ``DDnet().state_dict()``."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci_per_group, gain=1.0):
        std = float(np.sqrt(2.0 / (ci_per_group * 9)))
        sd[name + ".weight"] = torch.randn(co, ci_per_group, 3, 3, generator=g) * (std * gain)

    def cv2(prefix, ci, co):
        conv(prefix + ".convblock.0", co, ci)
        conv(prefix + ".convblock.2", co, co)

    sd["weight_tensor_in"] = 1.0 + 0.1 * torch.randn(9, 1, 1, 1, 1, generator=g)
    sd["weight_tensor_in2"] = 1.0 + 0.1 * torch.randn(9, 1, 4, 1, 1, generator=g)
    sd["weight_tensor_in"][4] = 1.0         # gain of the residual path mosaic -> output: exactly 1 keeps the ADMM loop bounded
    sd["weight_tensor_out"] = (torch.tensor([1.0, 0.1]).view(2, 1, 1, 1, 1)
                               * (1.0 + torch.tensor([0.01, 0.05]).view(2, 1, 1, 1, 1) * torch.randn(2, 1, 3, 1, 1, generator=g)))
    for blk, per_frame, co_out in (("temp1", 1, 3), ("temp2", 3, 3), ("temp11", 4, 4)):
        conv(blk + ".inc.convblock.0", 90, 4)
        conv(blk + ".inc.convblock.2", 20, 90)
        conv(blk + ".inc_1.convblock.0", 90, per_frame)
        conv(blk + ".inc_1.convblock.2", 20, 90)
        for name, ci, co in (("downc0", 20, 40), ("downc1", 40, 80)):
            conv(f"{blk}.{name}.convblock.0", co, ci)
            cv2(f"{blk}.{name}.convblock.2", co, co)
        for name, ci, co in (("upc2", 80, 40), ("upc1", 40, 20)):
            cv2(f"{blk}.{name}.convblock.0", ci, ci)
            conv(f"{blk}.{name}.convblock.1", co * 4, ci)
        conv(blk + ".outc.convblock.0", 20, 20)
        conv(blk + ".outc.convblock.2", co_out, 20, gain=out_gain)
        if blk == "temp11":
            conv(blk + ".fusion.convblock.0", 4, 4)
            conv(blk + ".fusion.convblock.2", 3, 4, gain=0.1)
    return sd
