"""Full-size (BASELINE 512x512x8) checks of the tensor-core engines.  The CPU oracle is too slow at this size for every
test run, so the checker is the fp32 FFMA engine on the same device (``SCI_CONV_IMPL=ref``), which the small-size tests pin
to the oracle / the reference's golden vectors at <= 2.2e-6.  These runs cover what the 64x64 cases cannot: four 128-pixel
tiles per row, R-row super-tiles that do not divide evenly among 148 persistent CTAs, the 256-column split, the rs-stack
weight gradient accumulating over ~220 tiles per CTA.  Tolerances: north_star's 1e-3 max-abs for the TF32 denoisers;
gradients 2e-3 of their max (TF32 operands, fp32 accumulation)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fastdvd(impl):
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from oracle import synthetic
    os.environ["SCI_CONV_IMPL"] = impl
    m = DataParallelLike(FastDVDnet())
    m.load_state_dict({"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


def _inputs():
    from oracle import synthetic
    meas, mask, orig = synthetic.make_case(512, 512, 8, 3000, bayer=True)
    g = torch.Generator().manual_seed(5)
    rgb = torch.from_numpy(orig).permute(2, 0, 1).unsqueeze(1).repeat(1, 3, 1, 1) + 0.05 * torch.randn(8, 3, 512, 512, generator=g)
    phi = torch.from_numpy(mask).permute(2, 0, 1).contiguous()
    return rgb.contiguous().cuda(), phi.cuda(), torch.from_numpy(meas).cuda()


def test_fastdvdnet_forward_and_gradients_512(cuda, monkeypatch):
    from adaptivepnp_sci_b200.ffdnet_adapter import _tile_loss
    monkeypatch.setenv("SCI_CONV_IMPL", "tc")
    u, phi, y = _inputs()
    res = {}
    for impl in ("tc", "ref"):
        m = _fastdvd(impl)
        eng = m.module.engine()
        out = eng.forward(u, 12 / 255, train=False).clone()
        eng.prepare(training=True)
        o_tr = eng.forward(u, 12 / 255, train=True)
        loss = torch.zeros(1, dtype=torch.float64, device=u.device)
        dout = _tile_loss(eng, o_tr, phi, y, "dout", loss, None, True)
        eng.backward(dout)
        res[impl] = (out.cpu(), eng.bucket.grad.clone().cpu(), float(loss))
        del m, eng
        torch.cuda.empty_cache()
    (o_tc, g_tc, l_tc), (o_ref, g_ref, l_ref) = res["tc"], res["ref"]
    assert float((o_tc - o_ref).abs().max()) < 1e-3
    assert abs(l_tc - l_ref) < 1e-4 * abs(l_ref)
    gmax = float(g_ref.abs().max())
    assert gmax > 0 and float((g_tc - g_ref).abs().max()) < 2e-3 * gmax
    # gradient direction: cosine similarity of the whole flat bucket
    cos = float((g_tc.double() * g_ref.double()).sum() / (g_tc.double().norm() * g_ref.double().norm()))
    assert cos > 0.9999


def test_ddnet_forward_512(cuda, monkeypatch):
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.network_demosaicking import DDnet
    from oracle import synthetic
    _, _, orig = synthetic.make_case(512, 512, 8, 3000, bayer=True)
    mosaic = torch.from_numpy(orig).permute(2, 0, 1).contiguous().cuda()
    outs = {}
    for impl in ("tc", "ref"):
        monkeypatch.setenv("SCI_CONV_IMPL", impl)
        m = DataParallelLike(DDnet())
        m.load_state_dict({"module." + k: v for k, v in synthetic.ddnet_synthetic_state_dict().items()}, strict=True)
        m = m.eval().cuda()
        outs[impl] = m.module.engine().forward(mosaic).clone().cpu()
        del m
        torch.cuda.empty_cache()
    assert outs["tc"].shape == (8, 3, 512, 512)
    assert float((outs["tc"] - outs["ref"]).abs().max()) < 1e-3


@pytest.mark.parametrize("shape", [(5, 72, 104), (3, 36, 200)])
def test_fastdvdnet_ragged_shapes_tc_vs_fp32(cuda, monkeypatch, shape):
    """Shapes that do not fill the 128-pixel tiles / R-row super-tiles / 8x8 weight-gradient tiles evenly: the tensor-core
    engine against the fp32 FFMA engine (forward, measurement loss, all parameter gradients)."""
    from adaptivepnp_sci_b200.ffdnet_adapter import _tile_loss
    monkeypatch.setenv("SCI_CONV_IMPL", "tc")
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 100 + H)
    u = torch.rand(B, 3, H, W, generator=g).cuda()
    phi = (torch.rand(B, H, W, generator=g) > 0.5).float().cuda()
    y = (torch.rand(H, W, generator=g) * B / 2).cuda()
    res = {}
    for impl in ("tc", "ref"):
        m = _fastdvd(impl)
        eng = m.module.engine()
        out = eng.forward(u, 12 / 255, train=False).clone()
        o_tr = eng.forward(u, 12 / 255, train=True)
        loss = torch.zeros(1, dtype=torch.float64, device=u.device)
        dout = _tile_loss(eng, o_tr, phi, y, "dout", loss, None, True)
        eng.backward(dout)
        res[impl] = (out.cpu(), eng.bucket.grad.clone().cpu(), float(loss))
    (o_tc, g_tc, l_tc), (o_ref, g_ref, l_ref) = res["tc"], res["ref"]
    assert float((o_tc - o_ref).abs().max()) < 1e-3
    assert abs(l_tc - l_ref) < 1e-4 * abs(l_ref)
    gmax = float(g_ref.abs().max())
    assert gmax > 0 and float((g_tc - g_ref).abs().max()) < 3e-3 * gmax
    cos = float((g_tc.double() * g_ref.double()).sum() / (g_tc.double().norm() * g_ref.double().norm()))
    assert cos > 0.9999
