"""GPU parity of the convolution engine (forward, dgrad, wgrad, packers, loss, Adam) against the oracle's
arithmetic (CPU PyTorch fp32) — run for both kernel families: the fp32 FFMA reference and the tcgen05 path.

Tolerances: fp32 reference kernels 1e-5 relative; tensor-core kernels use TF32 operands with fp32
accumulation (north_star: denoisers may run in bf16/tf32) -> 2e-3 relative per layer."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")

IMPLS = ["ref", "tc"]
TOL = {"ref": 1e-5, "tc": 2e-3}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _tc_available():
    from adaptivepnp_sci_b200._lib import lib
    return bool(lib.sci_conv_tc_available())


@pytest.fixture(params=IMPLS)
def impl(request, monkeypatch):
    if request.param == "tc" and not _tc_available():
        pytest.skip("tensor-core conv path not built yet")
    monkeypatch.setenv("SCI_CONV_IMPL", request.param)
    return request.param


CASES = [
    # Ci, Co, groups, stride, ps, relu, bn, bias, residual, H, W, N
    (13, 96, 1, 1, False, True, False, True, False, 20, 28, 2),      # FFDNet head
    (96, 96, 1, 1, False, True, False, True, False, 16, 24, 2),      # FFDNet body
    (96, 12, 1, 1, False, False, False, True, False, 16, 24, 1),     # FFDNet tail
    (12, 90, 3, 1, False, True, True, False, False, 16, 16, 2),      # FastDVDnet grouped input conv
    (32, 64, 1, 2, False, True, True, False, False, 24, 32, 2),      # DownBlock stride 2
    (64, 128, 1, 2, False, True, True, False, False, 16, 24, 2),     # DownBlock stride 2 (sub-pixel dgrad with 256 columns)
    (128, 256, 1, 1, True, False, False, False, True, 8, 12, 2),     # UpBlock conv + PixelShuffle + skip add
    (32, 3, 1, 1, False, False, False, False, False, 18, 20, 1),     # DenBlock output conv
]


@pytest.mark.parametrize("case", CASES)
def test_conv_layer_fwd_bwd(cuda, impl, case):
    from adaptivepnp_sci_b200 import engine
    from adaptivepnp_sci_b200._lib import call, ptr, stream
    Ci, Co, groups, stride, ps, relu, bn, bias, residual, H, W, N = case
    g = torch.Generator().manual_seed(Ci * 1000 + Co)
    conv = torch.nn.Conv2d(Ci, Co, 3, stride=stride, padding=1, groups=groups, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
        if bias:
            conv.bias.copy_(torch.randn(Co, generator=g) * 0.1)
    bnm = None
    if bn:
        bnm = torch.nn.BatchNorm2d(Co).eval()
        with torch.no_grad():
            bnm.weight.copy_(1 + 0.2 * torch.randn(Co, generator=g)); bnm.bias.copy_(0.1 * torch.randn(Co, generator=g))
            bnm.running_mean.copy_(0.1 * torch.randn(Co, generator=g)); bnm.running_var.copy_(1 + 0.3 * torch.rand(Co, generator=g))
    x = torch.randn(N, Ci, H, W, generator=g)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    # ---- oracle (CPU autograd)
    xr = x.clone().requires_grad_(True)
    z = conv(xr)
    if bn:
        z = bnm(z)
    if relu:
        z = F.relu(z)
    if ps:
        z = F.pixel_shuffle(z, 2)
    res = torch.randn(z.shape, generator=g) if residual else None
    yref = z + res if residual else z
    dy = torch.randn(yref.shape, generator=g)
    yref.backward(dy)
    # ---- device
    cconv = torch.nn.Conv2d(Ci, Co, 3, stride=stride, padding=1, groups=groups, bias=bias).cuda()
    cconv.load_state_dict(conv.state_dict())
    cbn = None
    if bn:
        cbn = torch.nn.BatchNorm2d(Co).cuda().eval()
        cbn.load_state_dict(bnm.state_dict())
    L = engine.ConvLayer(cconv, cbn, relu=relu, stride=stride, ps=ps)
    eng = engine._EngineBase(torch.nn.ModuleList([cconv] + ([cbn] if bn else [])), [L])
    eng.prepare(training=True)
    xd = torch.zeros(N, H, W, L.Ci_pad, device=cuda)
    xd[..., :Ci] = x.permute(0, 2, 3, 1).to(cuda)
    och = L.out_ch
    oh, ow = (2 * Ho, 2 * Wo) if ps else (Ho, Wo)
    y = torch.empty(N, oh, ow, och, device=cuda)
    resd = None
    if residual:
        resd = res.permute(0, 2, 3, 1).contiguous().to(cuda)
    eng.conv(L, xd, N, H, W, y, residual=resd, round_out=False)
    cout_valid = Co // 4 if ps else Co
    tol = TOL[impl]
    e_fwd = _rel(y[..., :cout_valid].permute(0, 3, 1, 2).cpu(), yref.detach())
    assert e_fwd < tol, "forward rel err %g" % e_fwd
    if not ps and L.Co_pad > Co:
        assert float(y[..., Co:].abs().max()) == 0.0                  # padded columns stay exactly zero
    # ---- backward: dz = act_bwd(dy), wgrad, dgrad
    dyd = torch.zeros(N, oh, ow, och, device=cuda)
    dyd[..., :cout_valid] = dy.permute(0, 2, 3, 1).to(cuda)
    if ps:
        dconv = torch.empty(N, Ho, Wo, L.Co_pad, device=cuda)
        call("sci_nhwc_pixel_unshuffle", ptr(dyd), ptr(dconv), N, Ho, Wo, och, stream())
        dyd, ystore = dconv, None
    else:
        # the ReLU mask is taken from the ORACLE's forward output: with TF32 operands a handful of near-zero
        # activations change sign, which would turn this per-layer check into a test of mask flips instead of
        # the dgrad / wgrad kernels (the end-to-end fine-tune tests cover the real mask)
        ystore = torch.zeros(N, oh, ow, och, device=cuda)
        ystore[..., :cout_valid] = yref.detach().permute(0, 2, 3, 1).to(cuda)
    eng.begin_backward()                               # zero the packed-gradient accumulators and the column sums
    dz = eng.act_bwd(L, dyd, ystore, N * Ho * Wo, L.Co_pad)
    eng.wgrad(L, xd, dz, N, H, W)
    dx = torch.empty(N, H, W, L.Ci_pad, device=cuda)
    if stride == 2:
        assert L.s2t                                   # stride-2 layers: sub-pixel data gradient at the low resolution
        eng.dgrad_s2(L, dz, N, Ho, Wo, dx)
        # the older formulation (zero-dilated dz convolved at full resolution) must agree with it
        wt = torch.empty(9 * L.Co_pad * L.Ci_pad, device=cuda)
        call("sci_conv_pack_weights", ptr(cconv.weight.data), ptr(wt), L.Co, L.Ci, 1, L.Co_pad, L.Ci_pad, 0, ptr(L.scale), 1,
             int(eng.tf32), 0, stream())
        dil = torch.empty(N, H, W, L.Co_pad, device=cuda)
        call("sci_nhwc_dilate2", ptr(dz), ptr(dil), N, Ho, Wo, L.Co_pad, stream())
        dx_old = torch.empty_like(dx)
        d = engine.ConvDesc(dil.data_ptr(), wt.data_ptr(), None, None, None, dx_old.data_ptr(), N, H, W, L.Co_pad, L.Ci_pad, 1, 0,
                            0, int(eng.tf32), 0, 0)
        call("sci_conv3x3_dgrad", ctypes.byref(d), eng.impl, stream())
        e_form = _rel(dx.cpu(), dx_old.cpu())
        assert e_form < {"ref": 1e-5, "tc": 1e-3}[impl], "sub-pixel vs dilated dgrad (outputs are TF32-rounded): %g" % e_form
    else:
        eng.dgrad(L, dz, N, Ho, Wo, dx)
    eng.param_grads(L)
    eng.finish_param_grads()                           # batched mode: unpack / BatchNorm / bias gradients of all layers in one launch
    btol = tol * 3
    errs = dict(dx=_rel(dx[..., :Ci].permute(0, 3, 1, 2).cpu(), xr.grad),
                dw=_rel(eng.bucket.grad_view(cconv.weight).cpu(), conv.weight.grad))
    if bias:
        errs["dbias"] = _rel(eng.bucket.grad_view(cconv.bias).cpu(), conv.bias.grad)
    if bn:
        errs["dgamma"] = _rel(eng.bucket.grad_view(cbn.weight).cpu(), bnm.weight.grad)
        errs["dbeta"] = _rel(eng.bucket.grad_view(cbn.bias).cpu(), bnm.bias.grad)
    assert all(v < btol for v in errs.values()), "backward rel errs %s (fwd %g)" % (errs, e_fwd)


def _ffdnet(cuda):
    from adaptivepnp_sci_b200.network_ffdnet import FFDNet
    m = FFDNet(3, 3, 96, 12, 'R')
    m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")), strict=True)
    return m.eval().cuda()


def _fastdvd(cuda):
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict
    m = DataParallelLike(FastDVDnet(num_input_frames=5))
    m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


def test_networks_vs_golden(cuda, impl):
    """Reference call conventions model(x, sigma) against outputs of the reference networks."""
    d = np.load(os.path.join(G, "networks.npz"))
    tol = {"ref": 2e-5, "tc": 3e-3}[impl]
    m = _ffdnet(cuda)
    y = m(torch.from_numpy(d["ffd_x"]).cuda(), torch.full((2, 1, 1, 1), 25 / 255).cuda())
    assert _rel(y.cpu(), d["ffd_y"]) < tol
    # odd size (23x37): replication pad to even, crop (models/network_ffdnet.py:56-59, :68)
    yo = m(torch.from_numpy(d["ffd_xo"]).cuda(), torch.full((1, 1, 1, 1), 25 / 255).cuda())
    assert tuple(yo.shape) == d["ffd_yo"].shape and _rel(yo.cpu(), d["ffd_yo"]) < tol
    f = _fastdvd(cuda)
    y5 = f(torch.from_numpy(d["fdvd_x"]).cuda(), torch.full((1, 1, 32, 48), 12 / 255).cuda())
    assert _rel(y5.cpu(), d["fdvd_y"]) < tol


def test_ffdnet_ragged_multi_tile(cuda, impl):
    """FFDNet inference where the half-resolution width spans several 128-pixel row tiles with a ragged last one (w/2 = 150,
    131) and odd sizes, colour (KAIR) and gray (KAIR + IPOL flavours), against the oracle networks on the CPU.  On the
    tensor-core path this is the fp16 value + remainder chain (sci_conv_desc.w_split with half_io)."""
    from adaptivepnp_sci_b200 import synthetic as syn
    from adaptivepnp_sci_b200.ffdnet_ipol_models import FFDNet as FFDNetIPOL
    from adaptivepnp_sci_b200.network_ffdnet import FFDNet
    from oracle import networks
    tol = {"ref": 2e-5, "tc": 2e-4}[impl]
    g = torch.Generator().manual_seed(21)
    sd = torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth"))
    om = networks.FFDNet(3, 3, 96, 12, 'R'); om.load_state_dict(sd, strict=True); om.eval()
    m = _ffdnet(cuda)
    for shape in ((2, 3, 36, 300), (1, 3, 19, 261)):
        x = torch.rand(*shape, generator=g)
        with torch.no_grad():
            want = om(x, torch.full((shape[0], 1, 1, 1), 12 / 255))
        got = m(x.cuda(), torch.full((shape[0], 1, 1, 1), 12 / 255).cuda())
        assert got.shape == want.shape and float((got.cpu() - want).abs().max()) < tol, shape
    # gray: KAIR flavour (15 layers x 64 features; random weights, the gray file is not shipped) and the IPOL flavour
    tg = torch.Generator().manual_seed(5)
    og = networks.FFDNet(1, 1, 64, 15, 'R')
    for p in og.parameters():
        p.data = torch.randn(p.shape, generator=tg) * (0.5 / max(1.0, float(np.sqrt(p[0].numel()))))
    mg = FFDNet(1, 1, 64, 15, 'R'); mg.load_state_dict(og.state_dict(), strict=True); mg = mg.eval().cuda()
    x = torch.rand(2, 1, 34, 262, generator=g)
    with torch.no_grad():
        want = og.eval()(x, torch.full((2, 1, 1, 1), 20 / 255))
    assert float((mg(x.cuda(), torch.full((2, 1, 1, 1), 20 / 255).cuda()).cpu() - want).abs().max()) < tol
    oi = networks.FFDNetIPOL(1); oi.load_state_dict(syn.ffdnet_ipol_synthetic_state_dict(1), strict=True); oi.eval()
    mi = FFDNetIPOL(1); mi.load_state_dict(syn.ffdnet_ipol_synthetic_state_dict(1), strict=True); mi = mi.eval().cuda()
    with torch.no_grad():
        want = oi(x, torch.FloatTensor([20 / 255, 20 / 255]))
    assert float((mi(x.cuda(), torch.full((2,), 20 / 255).cuda()).cpu() - want).abs().max()) < tol


def test_adapters_vs_golden(cuda, impl):
    """Both plug-in adapters, inference and online fine-tune, against the reference's outputs."""
    from adaptivepnp_sci_b200 import fastdvdnet_adapter as fa, ffdnet_adapter as ffa
    from oracle import sci_ops
    d = np.load(os.path.join(G, "adapters.npz"))
    tol = {"ref": 5e-5, "tc": 5e-3}[impl]
    x = torch.from_numpy(d["x"]).cuda()
    yall, Phiall, _, _ = (t.cuda() for t in sci_ops.bayer_split_init(torch.from_numpy(d["meas"]), torch.from_numpy(d["mask"]), None))
    lr = 2e-6
    # FFDNet
    m = _ffdnet(cuda)
    w0_before = m.state_dict()["model.0.weight"].clone()
    out = ffa.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, m, True, lr)
    assert out.shape == x.shape and _rel(out.cpu(), d["ffd_inf"]) < tol
    out, m2 = ffa.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, m, True, lr, True, 2)
    assert m2 is m and _rel(out.cpu(), d["ffd_upd"]) < tol
    losses = ffa.last_losses[0].cpu().numpy()
    assert np.allclose(losses, d["ffd_losses"], rtol=tol * 5)
    for key, gold in (("model.0.weight", d["ffd_w0_after"]), ("model.22.weight", d["ffd_w22_after"])):
        w = m.state_dict()[key].cpu().numpy()
        assert np.max(np.abs(w - gold)) <= 2 * 2 * lr * 1.01           # 2 Adam steps of at most ~lr each, either sign
        assert np.mean(np.abs(w - gold)) < 0.15 * lr
    assert float((m.state_dict()["model.0.weight"] - w0_before).abs().max()) > 0.5 * lr      # it did train
    # FastDVDnet
    f = _fastdvd(cuda)
    out = fa.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, f, True, lr)
    assert _rel(out.cpu(), d["fdvd_inf"]) < tol
    np.random.seed(42)
    out, f2 = fa.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, f, True, lr, True, 2)
    assert f2 is f and _rel(out.cpu(), d["fdvd_upd"]) < tol
    assert np.allclose(fa.last_losses[0].cpu().numpy(), d["fdvd_losses"], rtol=tol * 5)
    for key, gold in (("module.temp1.inc.convblock.0.weight", d["fdvd_w_first_after"]),
                      ("module.temp2.outc.convblock.1.weight", d["fdvd_bn_after"])):
        w = f.state_dict()[key].cpu().numpy()
        assert np.max(np.abs(w - gold)) <= 2 * 2 * lr * 1.01
        assert np.mean(np.abs(w - gold)) < 0.15 * lr


def test_adam_and_loss_kernels(cuda):
    from adaptivepnp_sci_b200._lib import call, ptr, stream
    g = torch.Generator().manual_seed(1)
    p = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) * s for s in (1.0, 1e-3, 1e-9)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=2e-6)
    pd, m, v = p.clone().cuda(), torch.zeros(1000, device=cuda), torch.zeros(1000, device=cuda)
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.clone()
        opt.step()
        gd = gr.cuda()
        call("sci_adam_step", ptr(pd), ptr(gd), ptr(m), ptr(v), 1000, 2e-6, 0.9, 0.999, 1e-8, step, stream())
        assert float((pd.cpu() - pr.detach()).abs().max()) <= 2.5e-7          # <= 2 ulp of a unit-scale fp32 weight
    # loss + gradient vs autograd
    from oracle import sci_ops
    B, H, W = 4, 12, 16
    xhat = torch.rand(B, 3, H, W, generator=g, requires_grad=True)
    phi = (torch.rand(B, H, W, generator=g) > 0.5).float()
    y = torch.rand(H, W, generator=g) * 2
    m4 = sci_ops.fourCh2OneCh(sci_ops.rgb_to_bayer4(xhat.permute(2, 3, 1, 0))).permute(2, 0, 1)
    loss = torch.nn.functional.mse_loss((m4 * phi).sum(0), y)
    loss.backward()
    dx = torch.empty(B, 3, H, W, device=cuda)
    lo = torch.zeros(1, dtype=torch.float64, device=cuda)
    xd, pd_, yd = xhat.detach().cuda(), phi.cuda(), y.cuda()       # keep the device tensors alive across the launch
    call("sci_meas_loss_fwd_bwd", ptr(xd), ptr(pd_), ptr(yd), ptr(dx), ptr(lo), H, W, B, 3, 0, stream())
    assert abs(float(lo) - float(loss.detach())) < 1e-5 * float(loss.detach())
    assert _rel(dx.cpu(), xhat.grad) < 1e-5
