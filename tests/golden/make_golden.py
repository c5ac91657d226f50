"""Generate the committed golden vectors FROM THE REFERENCE ITSELF and pin the oracle.

Run in the build container only (needs the read-only tree at /root/reference):

    python tests/golden/make_golden.py

For every case it (1) runs the reference's own function (imported through
``oracle/ref_harness.py``: CPU, import shims, restated skimage), (2) runs the
restatement in ``oracle/`` on the same inputs and ASSERTS agreement, and
(3) stores inputs' seeds + the reference outputs under ``tests/golden/``.
The committed files are what ``tests/`` (CPU and GPU) check against; the
reference tree is never needed at test time.
"""
import copy
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import adapters, admm, demosaic, networks, ref_harness, sci_ops, synthetic  # noqa: E402

torch.set_num_threads(8)


def _eq(a, b, what, tol=0.0):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().numpy() if torch.is_tensor(b) else np.asarray(b)
    err = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
    print(f"  {what:48s} max|ref-oracle| = {err:.3e}")
    assert err <= tol, f"oracle restatement disagrees with the reference on {what}: {err}"


def colour_kats():
    """Known-answer vectors quoted in the reference's doctests
    (malvar2004.py:70-95, masks.py:42-63) — typed in from the docstrings, then
    checked against the oracle's numpy restatement."""
    kats = {
        "malvar_RGGB_cfa": [[0.30980393, 0.36078432, 0.30588236, 0.3764706],
                            [0.35686275, 0.39607844, 0.36078432, 0.40000001]],
        "malvar_RGGB_rgb": [[[0.30980393, 0.31666668, 0.32941177], [0.33039216, 0.36078432, 0.38112746],
                             [0.30588236, 0.32794118, 0.34877452], [0.36274511, 0.3764706, 0.38480393]],
                            [[0.34828432, 0.35686275, 0.36568628], [0.35318628, 0.38186275, 0.39607844],
                             [0.3379902, 0.36078432, 0.3754902], [0.37769609, 0.39558825, 0.40000001]]],
        "malvar_BGGR_cfa": [[0.3764706, 0.360784320, 0.40784314, 0.3764706],
                            [0.35686275, 0.30980393, 0.36078432, 0.29803923]],
        "malvar_BGGR_rgb": [[[0.35539217, 0.37058825, 0.3764706], [0.34264707, 0.36078432, 0.37450981],
                             [0.36568628, 0.39607844, 0.40784314], [0.36568629, 0.3764706, 0.3882353]],
                            [[0.34411765, 0.35686275, 0.36200981], [0.30980393, 0.32990197, 0.34975491],
                             [0.33039216, 0.36078432, 0.38063726], [0.29803923, 0.30441178, 0.31740197]]],
        "masks_RGGB_3x3": [[[1, 0, 1], [0, 0, 0], [1, 0, 1]],
                           [[0, 1, 0], [1, 0, 1], [0, 1, 0]],
                           [[0, 0, 0], [0, 1, 0], [0, 0, 0]]],
        "masks_BGGR_3x3": [[[0, 0, 0], [0, 1, 0], [0, 0, 0]],
                           [[0, 1, 0], [1, 0, 1], [0, 1, 0]],
                           [[1, 0, 1], [0, 0, 0], [1, 0, 1]]],
    }
    for pat in ("RGGB", "BGGR"):
        out = demosaic.malvar2004_numpy(np.array(kats[f"malvar_{pat}_cfa"]), pat)
        _eq(out, np.array(kats[f"malvar_{pat}_rgb"]), f"malvar numpy KAT {pat}", tol=5e-8)
        m = demosaic.masks_CFA_Bayer((3, 3), pat)
        _eq(np.stack(m).astype(int), np.array(kats[f"masks_{pat}_3x3"]), f"masks KAT {pat}")
    with open(os.path.join(HERE, "colour_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)


def operators(ns):
    print("operators / projections / remaps")
    meas, mask, orig = synthetic.make_case(32, 48, 8, 77, bayer=True)
    mask[0:2, 0:2, :] = 0            # force Phi_sum == 0 pixels in every Bayer phase
    y, Phi = torch.from_numpy(meas), torch.from_numpy(mask)
    g = torch.Generator().manual_seed(5)
    yall, Phiall, Psum, x0 = sci_ops.bayer_split_init(y, Phi, None)
    theta = torch.rand(x0.shape, generator=g)
    b = 0.1 * torch.randn(x0.shape, generator=g)
    out = {}
    # reference A_/At_ on strided Bayer views (utilspy.py:28-44)
    for ib in range(4):
        _eq(ns.utilspy.A_(theta[..., ib], Phiall[..., ib]), sci_ops.A_(theta[..., ib], Phiall[..., ib]), f"A_ ib={ib}")
        _eq(ns.utilspy.At_(yall[..., ib], Phiall[..., ib]), sci_ops.At_(yall[..., ib], Phiall[..., ib]), f"At_ ib={ib}")
    out["A"] = torch.stack([ns.utilspy.A_(theta[..., ib], Phiall[..., ib]) for ib in range(4)], -1).numpy()
    out["At"] = torch.stack([ns.utilspy.At_(yall[..., ib], Phiall[..., ib]) for ib in range(4)], -1).numpy()
    out["x0_At"] = x0.numpy()
    out["Phi_sum"] = Psum.numpy()
    # the reference has the projections inlined in the solvers: run one solver iteration with tv disabled?
    # -> not separable; projections are pinned through the full-loop goldens below and here by formula on
    #    the reference's own A_/At_ (dvp:389-391, :128-140).
    x1 = torch.empty_like(theta)
    x2 = torch.empty_like(theta)
    for ib in range(4):
        v = theta[..., ib] + b[..., ib]
        x1[..., ib] = v + 1 * ns.utilspy.At_((yall[..., ib] - ns.utilspy.A_(v, Phiall[..., ib])) / (Psum[..., ib] + 0.01), Phiall[..., ib])
        p = theta[..., ib] - (1 / 0.55) * b[..., ib]
        t = (yall[..., ib] - ns.utilspy.A_(p, Phiall[..., ib])) / (1 * 0.55 + Psum[..., ib])
        x2[..., ib] = p + Phiall[..., ib] * torch.repeat_interleave(t.unsqueeze(2), 8, dim=2)
    _eq(x1, sci_ops.project_stage1(theta, b, yall, Phiall, Psum, 1, 0.01), "project_stage1")
    _eq(x2, sci_ops.project_stage2(theta, b, yall, Phiall, Psum, 1, 0.55), "project_stage2")
    out["proj1"], out["proj2"] = x1.numpy(), x2.numpy()
    # remaps (utils/utils_image.py:130-171)
    _eq(ns.utils_image.fourCh2OneCh(theta), sci_ops.fourCh2OneCh(theta), "fourCh2OneCh 4d")
    _eq(ns.utils_image.fourCh2OneCh(yall), sci_ops.fourCh2OneCh(yall), "fourCh2OneCh 3d")
    one = ns.utils_image.fourCh2OneCh(theta)
    _eq(ns.utils_image.oneCh2FourCh(one), sci_ops.oneCh2FourCh(one), "oneCh2FourCh")
    _eq(ns.utils_image.oneCh2ThreeCh(one), sci_ops.oneCh2ThreeCh(one), "oneCh2ThreeCh")
    rgb = torch.rand(32, 48, 3, 8, generator=g)
    import packages.fastdvdnet.utils as fu
    _eq(fu.gen_bayer_img(rgb, 1), sci_ops.gen_bayer_img(rgb, 1), "gen_bayer_img 1ch")
    _eq(fu.gen_bayer_img(rgb, 4), sci_ops.gen_bayer_img(rgb, 4), "gen_bayer_img 4ch")
    _eq(fu.gen_bayer_img(rgb, 4), sci_ops.rgb_to_bayer4(rgb), "rgb_to_bayer4 == gen_bayer_img(.,4)")
    out["merge"] = one.numpy()
    out["rgb_to_bayer4"] = fu.gen_bayer_img(rgb, 4).numpy()
    rm = ns.utils_image.masks_CFA_Bayer_tensor((6, 8))
    om = sci_ops.masks_CFA_Bayer_tensor((6, 8))
    for a, c in zip(rm, om):
        _eq(a.int(), c.int(), "masks_CFA_Bayer_tensor")
    # tensor Malvar (malvar2004.py:169-246), odd-ish size to exercise borders
    cfa = torch.rand(20, 28, generator=g)
    R_m, G_m, B_m = ns.utils_image.masks_CFA_Bayer_tensor((20, 28))
    ref = ns.malvar.demosaicing_CFA_Bayer_Malvar2004_tensor(cfa, R_m, G_m, B_m)
    _eq(ref, demosaic.malvar2004_tensor(cfa, R_m, G_m, B_m), "malvar2004 tensor")
    out["malvar_cfa"], out["malvar_rgb"] = cfa.numpy(), ref.numpy()
    np.savez_compressed(os.path.join(HERE, "operators.npz"), theta=theta.numpy(), b=b.numpy(), rgb=rgb.numpy(),
                        meas=meas, mask=mask, **out)


def _ref_ffdnet(ns):
    m = ns.network_ffdnet.FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode='R')
    m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")), strict=True)
    return m.eval()


def _orc_ffdnet():
    m = networks.FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode='R')
    m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")), strict=True)
    return m.eval()


def _ref_fastdvd(ns):
    m = ns.fastdvd_models.FastDVDnet(num_input_frames=5)
    m = torch.nn.DataParallel(m)
    sd = {"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}
    m.load_state_dict(sd, strict=True)
    return m.eval()


def _orc_fastdvd():
    m = networks.Wrapped(networks.FastDVDnet(num_input_frames=5))
    sd = {"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}
    m.load_state_dict(sd, strict=True)
    return m.eval()


def _ddnet_sd():
    return {"module." + k: v for k, v in synthetic.ddnet_synthetic_state_dict().items()}


def _ref_ddnet(ns):
    m = torch.nn.DataParallel(ns.network_demosaicking.DDnet())
    m.load_state_dict(_ddnet_sd(), strict=True)
    return m.eval()


def _orc_ddnet():
    m = networks.Wrapped(networks.DDnet())
    m.load_state_dict(_ddnet_sd(), strict=True)
    return m.eval()


class _DmArgs:
    dm_lr, dm_update_per_iter, dm_update = 1e-5, 2, True


def ddnet_(ns):
    """DDnet deep demosaic: network forward, the plug-in (inference and self-supervised update), and a stage-2 loop
    with ``model_demosaic`` (the scripts' default ``deep_demosaicking=True`` path)."""
    print("DDnet")
    g = torch.Generator().manual_seed(31)
    out = {}
    x = torch.rand(2, 15, 16, 24, generator=g)
    with torch.no_grad():
        r = _ref_ddnet(ns)(x)
        _eq(r, _orc_ddnet().module(x), "DDnet forward")
    out.update(net_x=x.numpy(), net_y=r.numpy())
    H, W, B = 32, 48, 8
    meas, mask, orig = synthetic.make_case(H, W, B, 92, bayer=True)
    yall, Phiall, _, _ = sci_ops.bayer_split_init(torch.from_numpy(meas), torch.from_numpy(mask), None)
    mosaic = torch.from_numpy(orig) + 0.05 * torch.randn(H, W, B, generator=g)
    v = sci_ops.oneCh2ThreeCh(mosaic)
    r = ns.ddnet_adapter.test_ddnet(v, yall, Phiall, _ref_ddnet(ns))
    _eq(r, adapters.test_ddnet(v, yall, Phiall, _orc_ddnet()), "test_ddnet inference")
    out.update(ad_mosaic=mosaic.numpy(), ad_inf=r.numpy())
    v_odd = v[:30, :46]                                    # reflect-pad-to-4 path (DDnet_test.py:180-187)
    r = ns.ddnet_adapter.test_ddnet(v_odd, None, None, _ref_ddnet(ns))
    _eq(r, adapters.test_ddnet(v_odd, None, None, _orc_ddnet()), "test_ddnet inference (30x46: reflect pad)")
    out.update(ad_inf_odd=r.numpy())
    rm, om = _ref_ddnet(ns), _orc_ddnet()
    r, rm = ns.ddnet_adapter.test_ddnet(v, yall, Phiall, rm, True, _DmArgs)
    losses = []
    o, om = adapters.test_ddnet(v, yall, Phiall, om, True, _DmArgs, losses=losses)
    _eq(r, o, "test_ddnet update output")
    _eq(torch.cat([a.flatten() for a in rm.state_dict().values()]),
        torch.cat([c.flatten() for c in om.state_dict().values()]), "ddnet fine-tuned weights (all keys)")
    out.update(ad_upd=r.numpy(), ad_losses=np.array(losses),
               ad_w_first_after=rm.state_dict()["module.temp1.inc_1.convblock.0.weight"].numpy(),
               ad_mix_after=rm.state_dict()["module.weight_tensor_in2"].numpy())
    # stage 2 with the deep demosaicker, 64x64x8 (same case / warm start as loops())
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    warm = np.load(os.path.join(HERE, "loops.npz"))["s2_warm"]
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_ffdnet(ns),
                                          model_demosaic=_ref_ddnet(ns), logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_ffdnet(),
                                        model_demosaic=_orc_ddnet(), **kw)
    _eq(r[0], o[0], "stage2 ffdnet+ddnet xbgr3")
    _eq(r[1], o[1], "stage2 ffdnet+ddnet x_bayer")
    _eq(np.array(r[4]), np.array(o[4]), "stage2 ffdnet+ddnet psnr_all")
    print("   psnr_all:", np.round(np.array(r[4]), 2))
    out.update(s2_rgb=r[0], s2_x=r[1], s2_psnr_all=np.array(r[4]), s2_psnr=np.array(r[2]), s2_ssim=np.array(r[3]))
    kw = dict(kw, update_times=-1)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_fastdvd(ns),
                                          model_demosaic=_ref_ddnet(ns), logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_fastdvd(),
                                        model_demosaic=_orc_ddnet(), **kw)
    _eq(r[0], o[0], "stage2 fastdvd+ddnet xbgr3")
    _eq(r[1], o[1], "stage2 fastdvd+ddnet x_bayer")
    print("   psnr_all:", np.round(np.array(r[4]), 2))
    out.update(s2f_rgb=r[0], s2f_x=r[1], s2f_psnr_all=np.array(r[4]))
    np.savez_compressed(os.path.join(HERE, "ddnet.npz"), **out)


def closed_form(ns):
    """``close_form_demosaic=True`` branch of stage 2 (dvp:112-118, 175-182, 224-230): tau = 10, rho = 0.55 and the
    closed-form x_rgb update from k = 1 on (k = 0 demosaics with Malvar)."""
    print("closed-form demosaic branch")
    out = {}
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    warm = np.load(os.path.join(HERE, "loops.npz"))["s2_warm"]
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2,
              close_form_demosaic=True)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_ffdnet(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_ffdnet(), **kw)
    _eq(r[0], o[0], "closed-form ffdnet xbgr3")
    _eq(r[1], o[1], "closed-form ffdnet x_bayer")
    _eq(np.array(r[4]), np.array(o[4]), "closed-form ffdnet psnr_all")
    print("   psnr_all:", np.round(np.array(r[4]), 2))
    out.update(ffd_rgb=r[0], ffd_x=r[1], ffd_psnr_all=np.array(r[4]))
    kw = dict(kw, update_times=-1)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_fastdvd(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_fastdvd(), **kw)
    _eq(r[0], o[0], "closed-form fastdvd xbgr3")
    _eq(r[1], o[1], "closed-form fastdvd x_bayer")
    _eq(np.array(r[4]), np.array(o[4]), "closed-form fastdvd psnr_all")
    print("   psnr_all:", np.round(np.array(r[4]), 2))
    out.update(fdvd_rgb=r[0], fdvd_x=r[1], fdvd_psnr_all=np.array(r[4]))
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'tv', [6], False, [0], x0_bayer=torch.from_numpy(warm),
                                          X_orig=orig, show_iqa=True, logf=io.StringIO(), close_form_demosaic=True)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'tv', [6], False, [0], x0_bayer=torch.from_numpy(warm),
                                        X_orig=orig, close_form_demosaic=True)
    _eq(r[0], o[0], "closed-form flag with tv (rho = 0.55)")
    out.update(tv_x=r[0], tv_psnr_all=np.array(r[3]))
    np.savez_compressed(os.path.join(HERE, "closed_form.npz"), **out)


def nets(ns):
    print("networks")
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 24, 40, generator=g)
    s = torch.full((2, 1, 1, 1), 25 / 255)
    with torch.no_grad():
        r = _ref_ffdnet(ns)(x, s)
        _eq(r, _orc_ffdnet()(x, s), "FFDNet-color forward")
        xo = torch.rand(1, 3, 23, 37, generator=g)      # odd size: replication pad path (network_ffdnet.py:56-59)
        ro = _ref_ffdnet(ns)(xo, s[:1])
        _eq(ro, _orc_ffdnet()(xo, s[:1]), "FFDNet-color forward (odd size)")
        x5 = torch.rand(1, 15, 32, 48, generator=g)
        nm = torch.full((1, 1, 32, 48), 12 / 255)
        r5 = _ref_fastdvd(ns)(x5, nm)
        _eq(r5, _orc_fastdvd()(x5, nm), "FastDVDnet forward")
    np.savez_compressed(os.path.join(HERE, "networks.npz"), ffd_x=x.numpy(), ffd_y=r.numpy(), ffd_xo=xo.numpy(),
                        ffd_yo=ro.numpy(), fdvd_x=x5.numpy(), fdvd_y=r5.numpy())


def adapters_(ns):
    print("adapters (inference + online fine-tune)")
    g = torch.Generator().manual_seed(21)
    H, W, B = 32, 48, 8
    meas, mask, orig = synthetic.make_case(H, W, B, 91, bayer=True)
    yall, Phiall, _, _ = sci_ops.bayer_split_init(torch.from_numpy(meas), torch.from_numpy(mask), None)
    x = torch.rand(H, W, 3, B, generator=g)
    out = dict(x=x.numpy(), meas=meas, mask=mask)
    # FFDNet
    rm, om = _ref_ffdnet(ns), _orc_ffdnet()
    r = ns.ffdnet_adapter.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, rm, True, 2e-6)
    _eq(r, adapters.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, om, True, 2e-6), "ffdnet adapter inference")
    out["ffd_inf"] = r.numpy()
    r, rm = ns.ffdnet_adapter.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, rm, True, 2e-6, True, 2)
    losses = []
    o, om = adapters.ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, 25 / 255, om, True, 2e-6, True, 2, losses=losses)
    _eq(r, o, "ffdnet adapter fine-tune output")
    _eq(torch.cat([a.flatten() for a in rm.state_dict().values()]),
        torch.cat([c.flatten() for c in om.state_dict().values()]), "ffdnet fine-tuned weights (all keys)")
    out["ffd_upd"] = r.detach().numpy()
    out["ffd_losses"] = np.array(losses)
    out["ffd_w0_after"] = rm.state_dict()["model.0.weight"].numpy()
    out["ffd_w22_after"] = rm.state_dict()["model.22.weight"].numpy()
    # FastDVDnet
    rm, om = _ref_fastdvd(ns), _orc_fastdvd()
    r = ns.fastdvd_adapter.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, rm, True, 2e-6)
    _eq(r, adapters.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, om, True, 2e-6), "fastdvd adapter inference")
    out["fdvd_inf"] = r.numpy()
    ns.utilspy.worker_init_fn(0)
    r, rm = ns.fastdvd_adapter.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, rm, True, 2e-6, True, 2)
    ns.utilspy.worker_init_fn(0)
    losses = []
    o, om = adapters.fastdvdnet_denoiser_full_tensor_v2(x, 12 / 255, yall, Phiall, om, True, 2e-6, True, 2, losses=losses)
    _eq(r, o, "fastdvd adapter fine-tune output")
    _eq(torch.cat([a.flatten().float() for a in rm.state_dict().values()]),
        torch.cat([c.flatten().float() for c in om.state_dict().values()]), "fastdvd fine-tuned weights (all keys)")
    out["fdvd_upd"] = r.detach().numpy()
    out["fdvd_losses"] = np.array(losses)
    out["fdvd_w_first_after"] = rm.state_dict()["module.temp1.inc.convblock.0.weight"].numpy()
    out["fdvd_bn_after"] = rm.state_dict()["module.temp2.outc.convblock.1.weight"].numpy()
    np.savez_compressed(os.path.join(HERE, "adapters.npz"), **out)


def loops(ns):
    print("full ADMM loops")
    out = {}
    # stage 1: gray cube treated as 4 interleaved sub-problems (config 1, reduced size)
    meas, mask, orig = synthetic.make_case(64, 64, 8, 1001, bayer=False)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None,
                                               X_orig=orig, model=None, show_iqa=True, logf=io.StringIO())
    tr = {}
    o = admm.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None,
                                             X_orig=orig, trace=tr)
    _eq(r[0], o[0], "stage1 tv x_bayer")
    _eq(np.array(r[3]), np.array(o[3]), "stage1 tv psnr_all")
    _eq(np.array(r[1]), np.array(o[1]), "stage1 tv psnr_")
    _eq(np.array(r[2]), np.array(o[2]), "stage1 tv ssim_")
    stops = np.array(tr['tv_stops'])
    print("   tv early stops (i<5):", int((stops < 5).sum()), "of", stops.size)
    out.update(s1_x=r[0], s1_psnr_all=np.array(r[3]), s1_psnr=np.array(r[1]), s1_ssim=np.array(r[2]), s1_stops=stops)

    # stage 2, Bayer 64x64x8, warm start from a stage-1 run
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    ns.utilspy.worker_init_fn(0)
    warm = ns.dvp.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None,
                                                  X_orig=orig, model=None, show_iqa=False, logf=io.StringIO())[0]
    out.update(s2_warm=warm)
    # 'tv' branch of stage 2
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'tv', [6], False, [0], x0_bayer=torch.from_numpy(warm),
                                          X_orig=orig, show_iqa=True, logf=io.StringIO())
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'tv', [6], False, [0], x0_bayer=torch.from_numpy(warm),
                                        X_orig=orig)
    _eq(r[0], o[0], "stage2 tv x_bayer")
    _eq(np.array(r[3]), np.array(o[3]), "stage2 tv psnr_all")
    out.update(s2tv_x=r[0], s2tv_psnr_all=np.array(r[3]))
    # ffdnet_color with online update (two_stage_ADMM_Online_FFD_Warm.py:71-80 schedule, shortened)
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_ffdnet(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_ffdnet(), **kw)
    _eq(r[0], o[0], "stage2 ffdnet xbgr3")
    _eq(r[1], o[1], "stage2 ffdnet x_bayer")
    _eq(np.array(r[4]), np.array(o[4]), "stage2 ffdnet psnr_all")
    out.update(s2ffd_rgb=r[0], s2ffd_x=r[1], s2ffd_psnr_all=np.array(r[4]), s2ffd_psnr=np.array(r[2]),
               s2ffd_ssim=np.array(r[3]))
    # no-update variant (pure inference loop)
    kw0 = dict(kw, update_=False)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [3], False, [25 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_ffdnet(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw0)
    out.update(s2ffd0_rgb=r[0], s2ffd0_x=r[1], s2ffd0_psnr_all=np.array(r[4]))
    # fastdvd_color with online update (two_stage_ADMM_Online_FastDVD_Warm.py:72-79 schedule, shortened)
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2,
              update_times=-1)
    ns.utilspy.worker_init_fn(0)
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_ref_fastdvd(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw)
    ns.utilspy.worker_init_fn(0)
    o = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [5, 2], False, [12 / 255, 6 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=_orc_fastdvd(), **kw)
    _eq(r[0], o[0], "stage2 fastdvd xbgr3")
    _eq(r[1], o[1], "stage2 fastdvd x_bayer")
    _eq(np.array(r[4]), np.array(o[4]), "stage2 fastdvd psnr_all")
    out.update(s2fdvd_rgb=r[0], s2fdvd_x=r[1], s2fdvd_psnr_all=np.array(r[4]), s2fdvd_psnr=np.array(r[2]),
               s2fdvd_ssim=np.array(r[3]))
    np.savez_compressed(os.path.join(HERE, "loops.npz"), **out)


def stage1_deep(ns):
    """Deep branches of stage 1 (admm_denoise_bayer_demosaic_pre with 'ffdnet_color' / 'fastdvd_color', dvp:456-503, :552)."""
    print("stage-1 deep branches")
    out = {}
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    warm = np.load(os.path.join(HERE, "loops.npz"))["s2_warm"]
    for tag, den, ref_m, orc_m, sig, its in (("ffd", 'ffdnet_color', _ref_ffdnet, _orc_ffdnet, [25 / 255, 12 / 255], [3, 2]),
                                             ("fdvd", 'fastdvd_color', _ref_fastdvd, _orc_fastdvd, [12 / 255], [4])):
        for wtag, x0 in (("", torch.from_numpy(warm)), ("_cold", None)):
            r = ns.dvp.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, den, its, False, sig, x0_bayer=x0, X_orig=orig,
                                                       model=ref_m(ns), show_iqa=True, logf=io.StringIO())
            o = admm.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, den, its, False, sig, x0_bayer=x0, X_orig=orig,
                                                     model=orc_m(None) if False else (orc_m()), show_iqa=True)
            assert len(r) == 6
            _eq(r[0], o[0], "stage1 %s%s xbgr3" % (den, wtag))
            _eq(r[1], o[1], "stage1 %s%s x_bayer" % (den, wtag))
            _eq(np.array(r[4]), np.array(o[4]), "stage1 %s%s psnr_all" % (den, wtag))
            out.update({"%s%s_rgb" % (tag, wtag): r[0], "%s%s_x" % (tag, wtag): r[1], "%s%s_psnr_all" % (tag, wtag): np.array(r[4]),
                        "%s%s_psnr" % (tag, wtag): np.array(r[2])})
    np.savez_compressed(os.path.join(HERE, "stage1_deep.npz"), **out)


def gray_adapters(ns):
    """Frame-wise gray adapters: ffdnet_vdenoiser (test_ffdnet_ipol.py:103-181, IPOL-flavour gray FFDNet)
    and fastdvdnet_denoiser(gray=True) (test_fastdvdnet.py:149-235, single-channel FastDVDnet), inference."""
    print("gray frame-wise adapters")
    g = torch.Generator().manual_seed(77)
    v = torch.rand(32, 48, 6, generator=g).numpy().astype(np.float64)
    rm = ns.ffdnet_ipol_models.FFDNet(num_input_channels=1)
    rm.load_state_dict(synthetic.ffdnet_ipol_synthetic_state_dict(1), strict=True)
    om = networks.FFDNetIPOL(1)
    om.load_state_dict(synthetic.ffdnet_ipol_synthetic_state_dict(1), strict=True)
    r = ns.ffdnet_adapter.ffdnet_vdenoiser(v, 20 / 255, model=rm, useGPU=False)
    o = adapters.ffdnet_vdenoiser(v, 20 / 255, om)
    _eq(r, o, "ffdnet_vdenoiser")
    out = {"v": v, "ffd_v": r}
    rm3 = ns.ffdnet_ipol_models.FFDNet(num_input_channels=3).eval()     # colour flavour of the same class: network forward only
    rm3.load_state_dict(synthetic.ffdnet_ipol_synthetic_state_dict(3), strict=True)
    om3 = networks.FFDNetIPOL(3).eval()
    om3.load_state_dict(synthetic.ffdnet_ipol_synthetic_state_dict(3), strict=True)
    x3 = torch.rand(2, 3, 32, 48, generator=g)
    with torch.no_grad():
        r3 = rm3(x3, torch.FloatTensor([15 / 255, 15 / 255]))
        o3 = om3(x3, torch.FloatTensor([15 / 255, 15 / 255]))
    _eq(r3, o3, "IPOL FFDNet colour forward")
    out.update(ffd_rgb_in=x3.numpy(), ffd_rgb_noise=r3.numpy())
    sd = {"module." + k: t for k, t in synthetic.fastdvdnet_gray_synthetic_state_dict().items()}
    rg = torch.nn.DataParallel(ns.fastdvd_models.FastDVDnet(num_input_frames=5, num_color_channels=1))
    rg.load_state_dict(sd, strict=True)
    og = networks.Wrapped(networks.FastDVDnet(num_input_frames=5, num_color_channels=1))
    og.load_state_dict(sd, strict=True)
    # 32x48: the reference's sequence driver pads its noise map AGAIN for every frame (fastdvdnet.py:129), so it only runs
    # on sizes that are already multiples of 4 (the oracle and the CUDA path pad once; tests cover 30x46 against the oracle)
    vg = torch.rand(32, 48, 6, generator=g).numpy().astype(np.float32)
    r = ns.fastdvd_adapter.fastdvdnet_denoiser(vg, 12 / 255, model=rg.eval(), useGPU=False, gray=True)
    o = adapters.fastdvdnet_denoiser(vg, 12 / 255, og, gray=True)
    _eq(r, o, "fastdvdnet_denoiser gray")
    out.update(vg=vg, fdvd_gray=r)
    np.savez_compressed(os.path.join(HERE, "gray_adapters.npz"), **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "gray_adapters":
    gray_adapters(ref_harness.load())
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stage1_deep":
    stage1_deep(ref_harness.load())
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ddnet":
    ddnet_(ref_harness.load())        # regenerate only tests/golden/ddnet.npz (needs loops.npz)
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "closed_form":
    closed_form(ref_harness.load())
    sys.exit(0)

if __name__ == "__main__":
    if not ref_harness.available():
        sys.exit("reference tree not present: golden vectors can only be regenerated in the build container")
    ns = ref_harness.load()
    colour_kats()
    operators(ns)
    nets(ns)
    adapters_(ns)
    loops(ns)
    ddnet_(ns)
    closed_form(ns)
    stage1_deep(ns)
    gray_adapters(ns)
    print("golden vectors written to", HERE)
