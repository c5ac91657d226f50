"""Oracle: the two ADMM solvers (the hot loop), restated on CPU.

Test infrastructure.  Restates
  * ``admm_denoise_bayer_demosaic_pre``  dvp_linear_inv_2_stage_ADMM_tensor_online.py:326-552 (stage 1, 'tv')
  * ``twoStageAdmm_denoise_bayer``       dvp_linear_inv_2_stage_ADMM_tensor_online.py:40-324  (stage 2)
including the iteration-0 aliasing of ``xall``/``theta_all``:
both names are bound to ``x0all`` until ``torch.clip`` / the TV result rebinds
``theta_all``, so in stage 2 the k=0 assignment ``theta_all[...,c] = ...``
also overwrites ``xall`` and ``b`` receives ``theta_unclipped - theta_clipped``.
Only the branches reachable from the three scripts are restated
('tv', 'ffdnet_color', 'fastdvd_color'; Malvar demosaic or the DDnet deep demosaic ``model_demosaic``).
"""
import numpy as np
import torch

from . import BAYER
from .adapters import fastdvdnet_denoiser_full_tensor_v2, ffdnet_rgb_denoise_full_tensor, test_ddnet
from .demosaic import malvar2004_tensor
from .iqa import compare_psnr, compare_ssim
from .sci_ops import (bayer_merge, bayer_split_init, fourCh2ThreeCh, masks_CFA_Bayer_tensor, oneCh2ThreeCh,
                      project_stage1, project_stage2)
from .tv_chambolle import denoise_tv_chambolle


def _listify(sigma, iter_max):
    if not isinstance(sigma, list):
        sigma = [sigma]
    if not isinstance(iter_max, list):
        iter_max = [iter_max] * len(sigma)
    return sigma, iter_max


def _tv(cube4, nrow, ncol, nmask, stops=None):
    v = cube4.reshape([nrow // 2, ncol // 2, nmask * 4]).numpy()
    out, st = denoise_tv_chambolle(v, 0.1, n_iter_max=5, multichannel=True, return_stops=True)
    if stops is not None:
        stops.append(st)
    return torch.from_numpy(out).reshape([nrow // 2, ncol // 2, nmask, 4])


def _final_iqa(X_orig, x_bayer_np, nmask):
    psnr_, ssim_ = [], []
    if X_orig is not None:
        for t in range(nmask):
            psnr_.append(compare_psnr(X_orig[:, :, t], x_bayer_np[:, :, t], data_range=1.))
            ssim_.append(compare_ssim(X_orig[:, :, t], x_bayer_np[:, :, t], data_range=1.))
    return psnr_, ssim_


def admm_denoise_bayer_demosaic_pre(y_bayer, Phi_bayer, _lambda=1, gamma=0.01, denoiser='tv', iter_max=50,
                                    noise_estimate=True, sigma=None, x0_bayer=None, X_orig=None, model=None,
                                    show_iqa=True, trace=None, **_unused):
    """Stage 1.  'tv' (the warm start the scripts use) returns (x_bayer_np[H,W,B], psnr_, ssim_, psnr_all); the deep branches
    'ffdnet_color' / 'fastdvd_color' (:456-500, inference only: see the product's note on update_) return the 6-tuple of :552."""
    if denoiser != 'tv' and denoiser.lower() not in ('ffdnet_color', 'fastdvd_color'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if denoiser != 'tv':
        return _stage1_deep(y_bayer, Phi_bayer, _lambda, gamma, denoiser.lower(), iter_max, sigma, x0_bayer, X_orig, model,
                            show_iqa)
    y_bayer = torch.from_numpy(np.ascontiguousarray(y_bayer))
    Phi_bayer = torch.from_numpy(np.ascontiguousarray(Phi_bayer))
    sigma, iter_max = _listify(sigma, iter_max)
    nrow, ncol, nmask = Phi_bayer.shape
    yall, Phiall, Phi_sumall, x0all = bayer_split_init(y_bayer, Phi_bayer, x0_bayer)
    xall = x0all
    ball = torch.zeros_like(x0all)
    theta_all = x0all
    psnr_all = []
    k = 0
    for idx, nsig in enumerate(sigma):
        for it in range(iter_max[idx]):
            project_stage1(theta_all, ball, yall, Phiall, Phi_sumall, _lambda, gamma, out=xall)   # :389-391
            theta_all = _tv(xall - ball, nrow, ncol, nmask,
                            None if trace is None else trace.setdefault('tv_stops', []))          # :403-407
            theta_all = torch.clip(theta_all, 0, 1)                                               # :501
            ball = ball - (xall - theta_all)                                                      # :503
            if show_iqa and X_orig is not None:
                psnr_all.append(compare_psnr(X_orig, bayer_merge(xall).numpy(), data_range=1.))   # :507-512
            k += 1
    x_bayer_np = bayer_merge(xall).numpy()
    psnr_, ssim_ = _final_iqa(X_orig, x_bayer_np, nmask)
    if trace is not None:
        trace.update(theta=theta_all.clone(), b=ball.clone(), x=xall.clone())
    return x_bayer_np, psnr_, ssim_, psnr_all


def _stage1_deep(y_bayer, Phi_bayer, _lambda, gamma, denoiser, iter_max, sigma, x0_bayer, X_orig, model, show_iqa):
    """dvp...online.py:456-503 with update_ = False: single dual variable, PSNR of x, k = 0 aliasing of xall / theta_all."""
    y_bayer = torch.from_numpy(np.ascontiguousarray(y_bayer))
    Phi_bayer = torch.from_numpy(np.ascontiguousarray(Phi_bayer))
    sigma, iter_max = _listify(sigma, iter_max)
    nrow, ncol, nmask = Phi_bayer.shape
    yall, Phiall, Phi_sumall, x0all = bayer_split_init(y_bayer, Phi_bayer, x0_bayer)
    xall = x0all
    theta_all = x0all                                                # :375-377 (same tensor)
    ball = torch.zeros_like(x0all)
    R_m, G_m, B_m = masks_CFA_Bayer_tensor((nrow, ncol))
    psnr_all, xbgr3 = [], None
    for idx, nsig in enumerate(sigma):
        for it in range(iter_max[idx]):
            project_stage1(theta_all, ball, yall, Phiall, Phi_sumall, _lambda, gamma, out=xall)
            x_bayer = bayer_merge(xall - ball)
            x_rgb = torch.zeros([nrow, ncol, 3, nmask])
            for t in range(nmask):
                x_rgb[:, :, :, t] = malvar2004_tensor(x_bayer[:, :, t], R_m, G_m, B_m)
            if denoiser == 'ffdnet_color':
                xbgr3 = ffdnet_rgb_denoise_full_tensor(x_rgb, yall, Phiall, nsig, model, True, 1e-6)
            else:
                xbgr3 = fastdvdnet_denoiser_full_tensor_v2(x_rgb, nsig, yall, Phiall, model, True, 1e-6)
            theta_all[..., 0] = xbgr3[0::2, 0::2, 0, :]
            theta_all[..., 1] = xbgr3[0::2, 1::2, 1, :]
            theta_all[..., 2] = xbgr3[1::2, 0::2, 1, :]
            theta_all[..., 3] = xbgr3[1::2, 1::2, 2, :]
            theta_all = torch.clip(theta_all, 0, 1)
            ball = ball - (xall - theta_all)
            if show_iqa and X_orig is not None:
                psnr_all.append(compare_psnr(X_orig, bayer_merge(xall).numpy(), data_range=1.))
    x_bayer_np = bayer_merge(xall).numpy()
    psnr_, ssim_ = _final_iqa(X_orig, x_bayer_np, nmask)
    return xbgr3.detach().numpy(), x_bayer_np, psnr_, ssim_, psnr_all, model


def twoStageAdmm_denoise_bayer(y_bayer, Phi_bayer, _lambda=1, gamma=0.01, denoiser='tv', iter_max=50,
                               noise_estimate=True, sigma=None, x0_bayer=None, X_orig=None, model_denoise=None,
                               model_demosaic=None, show_iqa=True, demosaic_method='malvar2004', lr_=1e-6,
                               inital_iter=1, interval_iter=5, logf=None, useGPU=True, update_=False,
                               update_per_iter=1, close_form_demosaic=False, large=False, update_times=-1,
                               args=None, trace=None, grad_hook=None, rng=None):
    """Stage 2.  'tv' -> 4-tuple; deep denoisers -> (xbgr3_np[H,W,3,B], x_bayer_np, psnr_, ssim_,
    psnr_all, model_denoise, model_demosaic)."""
    y_bayer = torch.from_numpy(np.ascontiguousarray(y_bayer))
    Phi_bayer = torch.from_numpy(np.ascontiguousarray(Phi_bayer))
    sigma, iter_max = _listify(sigma, iter_max)
    nrow, ncol, nmask = Phi_bayer.shape
    yall, Phiall, Phi_sumall, x0all = bayer_split_init(y_bayer, Phi_bayer, x0_bayer)
    xall = x0all                                                     # :87
    ball = torch.zeros_like(x0all)
    theta_all = x0all                                                # :89 (same tensor)
    R_m, G_m, B_m = masks_CFA_Bayer_tensor((nrow, ncol))
    w = torch.zeros([nrow, ncol, 3, nmask])
    alpha = 0.01 if denoiser == 'tv' else 1                          # :101-104
    rou = 0.55 if denoiser == 'fastdvd_color' else 1                 # :106-109
    tau = 100
    if close_form_demosaic:                                          # :112-118
        tau = 10
        rou = 0.55
        bayer_mask = torch.stack([R_m, G_m, B_m], dim=2)             # gen_bayer_mask, utils_image.py:115-118 (bool [H,W,3])
        inv_3ch = torch.repeat_interleave((rou * bayer_mask + tau).unsqueeze(3), nmask, dim=3)
    psnr_all = []
    k = 0
    update_i = 0
    xbgr3 = None
    for idx, nsig in enumerate(sigma):
        for it in range(iter_max[idx]):
            project_stage2(theta_all, ball, yall, Phiall, Phi_sumall, alpha, rou, out=xall)      # :128-140
            if denoiser == 'tv':
                TV = True
                theta_all = _tv(xall + (1 / rou) * ball, nrow, ncol, nmask)                      # :153-160
            elif denoiser.lower() in ('ffdnet_color', 'fastdvd_color'):
                TV = False
                x_rgb = torch.zeros([nrow, ncol, 3, nmask])
                x_bayer = bayer_merge(xall + (1 / rou) * ball)                                   # :169-172
                if close_form_demosaic and k > 0:                                                # :175-182, :224-230
                    x_rgb = (rou * fourCh2ThreeCh(xall) + fourCh2ThreeCh(ball) + tau * xbgr3 + w) / inv_3ch
                    if denoiser.lower() == 'ffdnet_color':
                        x_rgb = x_rgb.clip(0, 1)                                                 # :182 (FFDNet branch only)
                elif model_demosaic is not None:                                                 # :192-194, :241-243
                    x_rgb = test_ddnet(oneCh2ThreeCh(x_bayer), yall, Phiall, model_demosaic)
                elif demosaic_method == 'malvar2004':                                            # :185-191 (App. D.2)
                    for t in range(nmask):
                        x_rgb[:, :, :, t] = malvar2004_tensor(x_bayer[:, :, t], R_m, G_m, B_m)
                x_rgb_w = x_rgb - (1 / tau) * w                                                  # :198
                do_update = update_ and k > inital_iter and k % interval_iter == 0
                losses = None if trace is None else trace.setdefault('losses', [])
                if denoiser.lower() == 'ffdnet_color':
                    if do_update:
                        xbgr3, model_denoise = ffdnet_rgb_denoise_full_tensor(
                            x_rgb_w, yall, Phiall, nsig, model_denoise, useGPU, lr_, update_, update_per_iter,
                            losses=losses)
                    else:
                        xbgr3 = ffdnet_rgb_denoise_full_tensor(x_rgb_w, yall, Phiall, nsig, model_denoise, useGPU, lr_)
                else:
                    if do_update and (update_i < update_times or update_times < 0):             # :247
                        xbgr3, model_denoise = fastdvdnet_denoiser_full_tensor_v2(
                            x_rgb_w, nsig, yall, Phiall, model_denoise, useGPU, lr_, update_, update_per_iter,
                            update_times=update_times, losses=losses, grad_hook=grad_hook, rng=rng)
                        update_i += 1
                    else:
                        xbgr3 = fastdvdnet_denoiser_full_tensor_v2(x_rgb_w, nsig, yall, Phiall, model_denoise,
                                                                   useGPU, lr_)
                theta_all[..., 0] = xbgr3[0::2, 0::2, 0, :]                                      # :206-209
                theta_all[..., 1] = xbgr3[0::2, 1::2, 1, :]
                theta_all[..., 2] = xbgr3[1::2, 0::2, 1, :]
                theta_all[..., 3] = xbgr3[1::2, 1::2, 2, :]
            else:
                raise ValueError('Unsupported denoiser {}!'.format(denoiser))
            theta_all = torch.clip(theta_all, 0, 1)                                              # :265
            ball = ball + (xall - theta_all)                                                     # :267
            if not TV:
                w = w + (x_rgb - xbgr3)                                                          # :271
            if show_iqa and X_orig is not None:
                psnr_all.append(compare_psnr(X_orig, bayer_merge(theta_all).numpy(), data_range=1.))  # :275-280
            if trace is not None and trace.get('per_iter') is not None:
                trace['per_iter'].append(dict(theta=theta_all.clone(), b=ball.clone(), x=xall.clone()))
            k += 1
    x_bayer_np = bayer_merge(theta_all).numpy()
    psnr_, ssim_ = _final_iqa(X_orig, x_bayer_np, nmask)
    if trace is not None:
        trace.update(theta=theta_all.clone(), b=ball.clone(), x=xall.clone(), w=w.clone())
    if denoiser == 'tv':
        return x_bayer_np, psnr_, ssim_, psnr_all
    return xbgr3.detach().numpy(), x_bayer_np, psnr_, ssim_, psnr_all, model_denoise, model_demosaic


def twoStageAdmm_denoise_gray(y, Phi, denoiser='ffdnet_gray', iter_max=50, sigma=None, x0=None, X_orig=None,
                              model_denoise=None, show_iqa=True, lr_=1e-6, inital_iter=1, interval_iter=5, update_=False,
                              update_per_iter=1, trace=None):
    """DERIVED oracle for BASELINE config 2 (two-stage ADMM + online FFDNet-gray on a grayscale cube).  The reference's
    ``twoStageAdmm_denoise_bayer`` has no gray branch (dvp...online.py:147,164,214,262); this is the same loop
    (:121-305) with the Bayer split / demosaic replaced by the identity, as specified in SURVEY §8(c):
        p = theta - b/rho ; x = p + Phi*((y - A p)/(alpha*rho + Phi_sum))      (alpha = 1, rho = 1, tau = 100)
        x_pre = x + b/rho ; xhat = FFDNet_gray(x_pre - w/tau) [online fine-tune] ; theta = clip(xhat)
        b += x - theta ; w += x_pre - xhat
    including the k = 0 aliasing of xall/theta_all.  Returns (xhat_np[H,W,B], theta_np[H,W,B], psnr_, ssim_, psnr_all, model)."""
    from .adapters import ffdnet_gray_denoise_full_tensor
    from .sci_ops import A_
    y = torch.from_numpy(np.ascontiguousarray(y))
    Phi = torch.from_numpy(np.ascontiguousarray(Phi))
    sigma, iter_max = _listify(sigma, iter_max)
    Phi_sum = torch.sum(Phi, dim=2)
    Phi_sum[Phi_sum == 0] = 1
    x0all = (y.unsqueeze(2) * Phi) if x0 is None else x0.clone()
    xall = x0all
    theta = x0all
    ball = torch.zeros_like(x0all)
    w = torch.zeros_like(x0all)
    alpha, rou, tau = 1, 1, 100
    psnr_all, k = [], 0
    xhat = None
    for idx, nsig in enumerate(sigma):
        for it in range(iter_max[idx]):
            p = theta - (1 / rou) * ball
            t = (y - A_(p, Phi)) / (alpha * rou + Phi_sum)
            xall[...] = p + Phi * t.unsqueeze(2)                       # in place: aliases theta at k = 0
            x_pre = xall + (1 / rou) * ball
            x_w = x_pre - (1 / tau) * w
            do_update = update_ and k > inital_iter and k % interval_iter == 0
            losses = None if trace is None else trace.setdefault('losses', [])
            if do_update:
                xhat, model_denoise = ffdnet_gray_denoise_full_tensor(x_w, y, Phi, nsig, model_denoise, lr_, True,
                                                                      update_per_iter, losses=losses)
            else:
                xhat = ffdnet_gray_denoise_full_tensor(x_w, y, Phi, nsig, model_denoise, lr_)
            theta[...] = xhat                                          # writes into xall too while they alias (k = 0)
            theta = torch.clip(theta, 0, 1)
            ball = ball + (xall - theta)
            w = w + (x_pre - xhat)
            if show_iqa and X_orig is not None:
                psnr_all.append(compare_psnr(X_orig, theta.numpy(), data_range=1.))
            k += 1
    theta_np = theta.numpy()
    psnr_, ssim_ = _final_iqa(X_orig, theta_np, Phi.shape[2])
    return xhat.detach().numpy(), theta_np, psnr_, ssim_, psnr_all, model_denoise
