"""GPU parity of the DDnet deep-demosaic path against golden vectors produced by the
reference (tests/golden/ddnet.npz; weights = oracle.synthetic.ddnet_synthetic_state_dict, the trained file is absent)."""
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")

from test_gpu_conv import _fastdvd, _ffdnet, impl  # noqa: E402,F401

TOL = {"ref": 2e-5, "tc": 1e-3}      # fp32 FFMA kernels / TF32 tensor-core kernels (north_star: 1e-3 max-abs)


def _ddnet(cuda):
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.network_demosaicking import DDnet
    from oracle import synthetic
    m = DataParallelLike(DDnet())
    m.load_state_dict({"module." + k: v for k, v in synthetic.ddnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


def _psnr(a, b):
    return 10 * np.log10(1.0 / np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))


def test_ddnet_state_dict_keys(cuda):
    from adaptivepnp_sci_b200.network_demosaicking import DDnet
    from oracle import networks
    a, b = DDnet().state_dict(), networks.DDnet().state_dict()
    assert list(a.keys()) == list(b.keys()) and all(a[k].shape == b[k].shape for k in a)


def test_ddnet_window_forward(cuda, impl):
    d = np.load(os.path.join(G, "ddnet.npz"))
    m = _ddnet(cuda)
    x = torch.from_numpy(d["net_x"]).cuda()
    for n in range(x.shape[0]):
        y = m(x[n:n + 1]).cpu().numpy()
        assert y.shape == (1, 3, 16, 24)
        assert np.max(np.abs(y - d["net_y"][n:n + 1])) < TOL[impl]


def test_ddnet_adapter(cuda, impl):
    from adaptivepnp_sci_b200.ddnet_adapter import test_ddnet as ddnet_plugin
    from adaptivepnp_sci_b200.utils_image import oneCh2ThreeCh
    d = np.load(os.path.join(G, "ddnet.npz"))
    v = oneCh2ThreeCh(torch.from_numpy(d["ad_mosaic"]).cuda())
    out = ddnet_plugin(v, None, None, _ddnet(cuda))
    assert out.shape == (32, 48, 3, 8)
    assert np.max(np.abs(out.cpu().numpy() - d["ad_inf"])) < TOL[impl]



def test_ddnet_self_supervised_update(cuda, impl):
    """args.dm_update (DDnet_test.py:231-276): two update steps (fresh Adam each), losses, updated weights incl. the mixing
    tensors, and the output after the update, against the reference's values."""
    from adaptivepnp_sci_b200 import ddnet_adapter
    from adaptivepnp_sci_b200.utils_image import oneCh2ThreeCh
    d = np.load(os.path.join(G, "ddnet.npz"))
    v = oneCh2ThreeCh(torch.from_numpy(d["ad_mosaic"]).cuda())
    lr = 1e-5

    class Args:
        dm_lr, dm_update_per_iter, dm_update = lr, 2, True
    m = _ddnet(cuda)
    before = {k: t.clone() for k, t in m.state_dict().items()}
    out, m2 = ddnet_adapter.test_ddnet(v, None, None, m, True, Args)
    assert m2 is m and out.shape == (32, 48, 3, 8)
    tol = {"ref": 5e-5, "tc": 1e-3}[impl]
    assert np.max(np.abs(out.cpu().numpy() - d["ad_upd"])) < tol
    losses = ddnet_adapter.last_losses[0].cpu().numpy()
    assert np.allclose(losses, d["ad_losses"], rtol={"ref": 1e-5, "tc": 2e-3}[impl])
    for key, gold in (("module.temp1.inc_1.convblock.0.weight", d["ad_w_first_after"]),
                      ("module.weight_tensor_in2", d["ad_mix_after"])):
        w = m.state_dict()[key].cpu().numpy()
        assert np.max(np.abs(w - gold)) <= 2 * 2 * lr * 1.01           # 2 Adam steps of at most ~lr each, either sign
        assert np.mean(np.abs(w - gold)) < 0.15 * lr
        assert float((m.state_dict()[key] - before[key]).abs().max()) > 0.5 * lr        # it did train
    # the never-executed noise-map input blocks keep their weights (torch skips parameters without gradient)
    assert torch.equal(m.state_dict()["module.temp1.inc.convblock.0.weight"], before["module.temp1.inc.convblock.0.weight"])


def test_stage2_with_deep_demosaic(cuda, impl):
    """The scripts' default path (deep_demosaicking=True): stage 2 with model_demosaic, both denoisers, online update."""
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    from oracle import synthetic
    d = np.load(os.path.join(G, "ddnet.npz"))
    warm = np.load(os.path.join(G, "loops.npz"))["s2_warm"]
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2,
              logf=io.StringIO())
    dm = _ddnet(cuda)
    r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                   x0_bayer=np2tch_cuda(warm), X_orig=orig, model_denoise=_ffdnet(cuda), model_demosaic=dm, **kw)
    tol = {"ref": 2e-4, "tc": 1e-3}[impl]
    assert r[6] is dm
    assert np.max(np.abs(r[0] - d["s2_rgb"])) < tol and np.max(np.abs(r[1] - d["s2_x"])) < tol
    assert abs(_psnr(r[1], orig) - _psnr(d["s2_x"], orig)) < 0.05
    assert np.max(np.abs(np.array(r[4]) - d["s2_psnr_all"])) < 0.05
    worker_init_fn(0)
    r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                   x0_bayer=np2tch_cuda(warm), X_orig=orig, model_denoise=_fastdvd(cuda), model_demosaic=dm,
                                   update_times=-1, **kw)
    assert np.max(np.abs(r[0] - d["s2f_rgb"])) < tol and np.max(np.abs(r[1] - d["s2f_x"])) < tol
    assert np.max(np.abs(np.array(r[4]) - d["s2f_psnr_all"])) < 0.05


def test_sequence_drivers_reflect_pad_to_multiple_of_4(cuda, impl):
    """Frames whose size is not a multiple of 4 are reflect-padded (right / bottom) before the network and cropped after it
    (DDnet_test.py:180-196, fastdvdnet.py:119-141): DDnet against the reference's output for a 30x46 input, FastDVDnet
    against the oracle's restatement of the sequence driver."""
    from adaptivepnp_sci_b200 import fastdvdnet_adapter
    from adaptivepnp_sci_b200.ddnet_adapter import test_ddnet as ddnet_plugin
    from adaptivepnp_sci_b200.utils_image import oneCh2ThreeCh
    from oracle import adapters, networks, synthetic
    d = np.load(os.path.join(G, "ddnet.npz"))
    v = oneCh2ThreeCh(torch.from_numpy(d["ad_mosaic"]).cuda())[:30, :46].contiguous()
    out = ddnet_plugin(v, None, None, _ddnet(cuda))
    assert out.shape == (30, 46, 3, 8)
    assert np.max(np.abs(out.cpu().numpy() - d["ad_inf_odd"])) < TOL[impl]
    g = torch.Generator().manual_seed(9)
    seq = torch.rand(6, 3, 30, 46, generator=g)
    om = networks.Wrapped(networks.FastDVDnet(num_input_frames=5))
    om.load_state_dict({"module." + k: t for k, t in synthetic.fastdvdnet_synthetic_state_dict().items()}, strict=True)
    with torch.no_grad():
        want = adapters.fastdvdnet_seqdenoise(seq, torch.tensor([12 / 255]), 5, om.eval())
    got = fastdvdnet_adapter.fastdvdnet_seqdenoise(seq.cuda(), torch.tensor([12 / 255]).cuda(), 5, _fastdvd(cuda))
    assert got.shape == (6, 3, 30, 46)
    assert float((got.cpu() - want).abs().max()) < TOL[impl] * 5


def test_fastdvdnet_framewise_adapter(cuda):
    """fastdvdnet_denoiser (test_fastdvdnet.py:149-235, colour, inference): numpy [H,W,F,3] in / out, equals the sequence driver."""
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike, fastdvdnet_denoiser, fastdvdnet_seqdenoise
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from oracle import adapters, networks, synthetic
    sd = {"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}
    m = DataParallelLike(FastDVDnet()); m.load_state_dict(sd, strict=True); m = m.eval().cuda()
    g = torch.Generator().manual_seed(3)
    v = torch.rand(30, 46, 6, 3, generator=g).numpy()                       # 30x46: exercises the reflect pad to x4
    out = fastdvdnet_denoiser(v, 12 / 255, m)
    assert out.shape == v.shape and out.dtype == np.float32
    mo = networks.Wrapped(networks.FastDVDnet()); mo.load_state_dict(sd, strict=True); mo.eval()
    with torch.no_grad():
        ref = adapters.fastdvdnet_seqdenoise(torch.from_numpy(v).permute(2, 3, 0, 1), torch.FloatTensor([12 / 255]), 5, mo)
    assert np.max(np.abs(out - ref.permute(2, 3, 0, 1).numpy())) < 1e-3
    with pytest.raises(NotImplementedError):
        fastdvdnet_denoiser(v, 12 / 255, m, updata_=True)


def test_gray_framewise_adapters(cuda, impl):
    """ffdnet_vdenoiser (test_ffdnet_ipol.py:103-181) with the IPOL-flavour gray FFDNet and
    fastdvdnet_denoiser(gray=True) (test_fastdvdnet.py:149-235) with the single-channel FastDVDnet, against outputs of the
    reference's own functions (tests/golden/gray_adapters.npz); odd sizes against the oracle (the reference fails there)."""
    from adaptivepnp_sci_b200 import synthetic as syn
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike, fastdvdnet_denoiser
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from adaptivepnp_sci_b200.ffdnet_adapter import ffdnet_vdenoiser
    from adaptivepnp_sci_b200.ffdnet_ipol_models import FFDNet as FFDNetIPOL
    from oracle import adapters, networks, synthetic
    d = np.load(os.path.join(G, "gray_adapters.npz"))
    m = FFDNetIPOL(1)
    m.load_state_dict(syn.ffdnet_ipol_synthetic_state_dict(1), strict=True)
    m = m.cuda()
    out = ffdnet_vdenoiser(d["v"], 20 / 255, model=m)
    assert out.shape == d["v"].shape and out.dtype == np.float64
    assert np.max(np.abs(out - d["ffd_v"])) < TOL[impl]
    v4 = d["v"].reshape(32, 48, 3, 2)                                  # [M,N,F1,F2] input: frames are the flattened tail (:132-133)
    assert np.array_equal(ffdnet_vdenoiser(v4, 20 / 255, model=m).reshape(32, 48, 6), out)
    m3 = FFDNetIPOL(3)                                                  # colour flavour of the class: three noise planes first
    m3.load_state_dict(syn.ffdnet_ipol_synthetic_state_dict(3), strict=True)
    m3 = m3.cuda().eval()
    noise = m3(torch.from_numpy(d["ffd_rgb_in"]).cuda(), torch.full((2,), 15 / 255).cuda())
    assert float(np.max(np.abs(noise.cpu().numpy() - d["ffd_rgb_noise"]))) < TOL[impl]
    with pytest.raises(Exception):
        ffdnet_vdenoiser(d["v"], 20 / 255, model=None)

    sd = {"module." + k: t for k, t in syn.fastdvdnet_gray_synthetic_state_dict().items()}
    g = DataParallelLike(FastDVDnet(num_input_frames=5, num_color_channels=1))
    g.load_state_dict(sd, strict=True)
    g = g.eval().cuda()
    assert list(g.state_dict().keys()) == list(sd.keys())             # the colour twin is not part of the state
    got = fastdvdnet_denoiser(d["vg"], 12 / 255, g, gray=True)
    assert got.shape == d["vg"].shape and got.dtype == np.float32
    assert np.max(np.abs(got - d["fdvd_gray"])) < TOL[impl] * 5
    og = networks.Wrapped(networks.FastDVDnet(num_input_frames=5, num_color_channels=1))
    og.load_state_dict({"module." + k: t for k, t in synthetic.fastdvdnet_gray_synthetic_state_dict().items()}, strict=True)
    vo = torch.rand(30, 46, 5, generator=torch.Generator().manual_seed(5)).numpy()
    assert np.max(np.abs(fastdvdnet_denoiser(vo, 12 / 255, g, gray=True) - adapters.fastdvdnet_denoiser(vo, 12 / 255, og, gray=True))) < TOL[impl] * 5
    # weights changed in place -> the twin follows
    with torch.no_grad():
        g.module.temp2.outc.convblock[3].weight.mul_(2.0)
        og.module.temp2.outc.convblock[3].weight.mul_(2.0)
    assert np.max(np.abs(fastdvdnet_denoiser(vo, 12 / 255, g, gray=True) - adapters.fastdvdnet_denoiser(vo, 12 / 255, og, gray=True))) < TOL[impl] * 5
