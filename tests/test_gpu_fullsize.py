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


# ---------------------------------------------------------------------------------------------------------------------
# Reference-pinned parity at the headline size: tests/golden/fullsize.npz holds what the REFERENCE ITSELF produced on the
# CPU for BASELINE configs[3] (FastDVDnet, the headline) and configs[2] (FFDNet-colour) at 512x512x8, stage 1 feeding
# stage 2 (tests/golden/make_golden_fullsize.py).  The CUDA path runs its OWN stage 1 and stage 2.  Tolerances are
# north_star's: 1e-5 relative for the TV stage, 1e-3 max-abs and 0.05 dB on the final reconstruction.
# ---------------------------------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.npz")


def _frame_sums(a, frame_axis):
    a = np.asarray(a, dtype=np.float64)
    ax = tuple(i for i in range(a.ndim) if i != frame_axis)
    return a.sum(axis=ax), (a * a).sum(axis=ax)


def _check_against_gold(d, prefix, x_bayer, x_rgb, tol):
    s, rs = int(d["shape"][4]), int(d["rgb_stride"])
    n_px = x_bayer.shape[0] * x_bayer.shape[1]
    assert np.max(np.abs(x_bayer[::s, ::s] - d[prefix + "_x_s"])) < tol
    sm, sq = _frame_sums(x_bayer, 2)
    # checksums over EVERY pixel: a per-frame mean within tol, an rms within tol
    assert np.max(np.abs(sm - d[prefix + "_x_sum"])) / n_px < tol
    assert np.max(np.abs(np.sqrt(sq / n_px) - np.sqrt(d[prefix + "_x_sq"] / n_px))) < tol
    if x_rgb is not None:
        assert np.max(np.abs(x_rgb[::rs, ::rs] - d[prefix + "_rgb_s"])) < tol
        sm, sq = _frame_sums(x_rgb, 3)
        assert np.max(np.abs(sm - d[prefix + "_rgb_sum"])) / (3 * n_px) < tol


@pytest.fixture(scope="module")
def warm512(cuda):
    """Stage 1 of the product at 512x512x8 (40 TV iterations) checked against the reference's warm start."""
    import io
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import admm_denoise_bayer_demosaic_pre
    from oracle import synthetic
    d = np.load(GOLD)
    meas, mask, orig = synthetic.make_case(512, 512, 8, 3000, bayer=True)
    r = admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None, X_orig=orig,
                                        show_iqa=True, logf=io.StringIO())
    return d, meas, mask, orig, r


def test_stage1_tv_512_vs_reference(warm512):
    d, meas, mask, orig, r = warm512
    _check_against_gold(d, "s1", r[0], None, 2e-5)                        # <= 1e-5 relative on values in [0, 1.x]
    assert np.max(np.abs(np.array(r[3]) - d["s1_psnr_all"])) < 1e-3       # dB, all 40 iterations
    assert np.max(np.abs(np.array(r[1]) - d["s1_psnr"])) < 1e-3 and np.max(np.abs(np.array(r[2]) - d["s1_ssim"])) < 1e-4


@pytest.mark.parametrize("engine_impl", ["tc", "ref"])
def test_config4_fastdvdnet_512_vs_reference(warm512, engine_impl, monkeypatch):
    """BASELINE configs[3], the headline workload, end to end against the reference's own output."""
    import io
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    d, meas, mask, orig, r1 = warm512
    monkeypatch.setenv("SCI_CONV_IMPL", engine_impl)
    m = _fastdvd(engine_impl)
    worker_init_fn(0)
    r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [21, 2], False, [12 / 255, 6 / 255],
                                   x0_bayer=np2tch_cuda(r1[0]), X_orig=orig, model_denoise=m, model_demosaic=None,
                                   show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=9,
                                   logf=io.StringIO(), update_=True, update_per_iter=2, update_times=-1)
    tol = {"ref": 2e-4, "tc": 1e-3}[engine_impl]
    _check_against_gold(d, "c4", r[1], r[0], tol)
    assert np.max(np.abs(np.array(r[4]) - d["c4_psnr_all"])) < 0.05       # dB, every one of the 23 iterations
    assert np.max(np.abs(np.array(r[2]) - d["c4_psnr"])) < 0.05 and np.max(np.abs(np.array(r[3]) - d["c4_ssim"])) < 1e-3
    del m
    torch.cuda.empty_cache()


@pytest.mark.parametrize("engine_impl", ["tc", "ref"])
def test_config3_ffdnet_512_vs_reference(warm512, engine_impl, monkeypatch):
    """BASELINE configs[2] (FFDNet-colour + Malvar, sigma 25/12/6, iterations 6/6/4) against the reference's output."""
    import io
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.network_ffdnet import FFDNet
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    d, meas, mask, orig, r1 = warm512
    monkeypatch.setenv("SCI_CONV_IMPL", engine_impl)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    m = FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode='R')
    m.load_state_dict(torch.load(os.path.join(root, "model_zoo", "ffdnet_color.pth")), strict=True)
    m = m.eval().cuda()
    worker_init_fn(0)
    r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [6, 6, 4], False, [25 / 255, 12 / 255, 6 / 255],
                                   x0_bayer=np2tch_cuda(r1[0]), X_orig=orig, model_denoise=m, model_demosaic=None,
                                   show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=6,
                                   logf=io.StringIO(), update_=True, update_per_iter=2)
    # MEASURED (tools/ffdnet_dev.py): fp32 engine vs the reference 1.2e-5; tensor-core engine 5.3e-4 with the fp16 value +
    # remainder inference chain (one pass deviates 1.4e-5 from the fp32 engine, tools/ffdnet_pass.py; the loop's figure is set by
    # the TF32 fine-tune steps), 2.6e-4 with the "3xTF32" chain (SCI_FFDNET_INF=tf32).  (It was 1.23e-3 - above
    # north_star's bound - while all products of the "3xTF32" scheme accumulated into ONE TMEM accumulator: the tensor core
    # truncates when it accumulates, and FFDNet returns the image itself, so the 16 ADMM iterations integrate that bias through
    # the dual variables.  The small products now have their own accumulator, sci_conv_desc.lo_channel0.)
    tol = {"ref": 2e-4, "tc": 1e-3}[engine_impl]
    _check_against_gold(d, "c3", r[1], r[0], tol)
    assert np.max(np.abs(np.array(r[4]) - d["c3_psnr_all"])) < 0.05
    assert np.max(np.abs(np.array(r[2]) - d["c3_psnr"])) < 0.05 and np.max(np.abs(np.array(r[3]) - d["c3_ssim"])) < 1e-3
    del m
    torch.cuda.empty_cache()
