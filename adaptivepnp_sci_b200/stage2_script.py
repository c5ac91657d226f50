"""Shared body of the two stage-2 entry points (two_stage_ADMM_Online_FFD_Warm.py / ..._FastDVD_Warm.py).

Keeps the reference scripts' flow: per-video hyper-parameter tables (two_stage_ADMM_Online_FFD_Warm.py:68-151,
two_stage_ADMM_Online_FastDVD_Warm.py:66-166; ``--deep-demosaicking`` selects the scripts' ``deep_demosaicking=True``
columns and demosaics with DDnet, the default is the Malvar columns),
warm start from results/savedmat/_Admm_tv_<name>8.mat, loop over measurement groups with ``reuse_model``
carry-over, log lines and the result .mat.  With torchrun the groups are sharded over ranks; ``--share-weights``
keeps one set of denoiser weights across ranks via the NCCL gradient all-reduce (BASELINE config 4)."""
import argparse
import os
import time
from statistics import mean

import numpy as np
import scipy.io as sio
import torch

from . import matio, parallel
from .dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer as reconstruct
from .utilspy import mkdir, worker_init_fn

# (sigma*255 list, iter_max list, lr, update_per_iter, interval_iter, update_times)
FFD_TABLE = {          # two_stage_ADMM_Online_FFD_Warm.py:65-151, deep_demosaicking=False values
    'Beauty_bayer': ([25, 12, 6], [15, 6, 4], 2e-6, 2, 15, -1), 'Bosphorus_bayer': ([50, 25, 12, 6], [8, 4, 4, 4], 2e-6, 2, 8, -1),
    'Jockey_bayer': ([25, 12, 6], [16, 8, 4], 2e-6, 2, 16, -1), 'Runner_bayer': ([50, 25, 12, 6], [8, 4, 4, 4], 2e-6, 2, 8, -1),
    'ShakeNDry_bayer': ([50, 25, 12, 6], [8, 4, 4, 4], 2e-6, 2, 10, -1), 'Traffic_bayer': ([50, 25], [16, 8], 2e-6, 2, 16, -1),
}
FASTDVD_TABLE = {      # two_stage_ADMM_Online_FastDVD_Warm.py:61-166, deep_demosaicking=False values
    'Beauty_bayer': ([8], [18], 2e-6, 2, 9, 1), 'Bosphorus_bayer': ([12, 6], [24, 12], 2e-7, 2, 12, -1),
    'Jockey_bayer': ([12], [24], 2e-7, 2, 12, -1), 'Runner_bayer': ([14], [24], 2e-7, 2, 12, -1),
    'ShakeNDry_bayer': ([10], [15], 2e-7, 1, 7, -1), 'Traffic_bayer': ([30], [22], 2e-7, 2, 11, -1),
}


# overrides when deep_demosaicking=True (same script lines): (sigma*255, iter_max, interval_iter)
FFD_DEEP = {
    'Beauty_bayer': ([25, 12, 6], [6, 6, 4], 6), 'Bosphorus_bayer': ([25, 12, 6], [4, 4, 2], 8),
    'Jockey_bayer': ([12, 6], [16, 8], 16), 'Runner_bayer': ([25, 12, 6], [8, 8, 4], 10),
    'ShakeNDry_bayer': ([25, 12, 6], [8, 8, 4], 10), 'Traffic_bayer': ([25, 12], [14, 7], 14),
}
FASTDVD_DEEP = {
    'Beauty_bayer': ([12, 6], [21, 2], 22), 'Bosphorus_bayer': ([8, 6], [24, 12], 25), 'Jockey_bayer': ([12, 6], [24, 6], 25),
    'Runner_bayer': ([12, 6], [40, 15], 41), 'ShakeNDry_bayer': ([12, 6], [14, 4], 15), 'Traffic_bayer': ([25, 12, 6], [36, 6, 2], 43),
}


def build_demosaicker():
    """two_stage_ADMM_Online_FFD_Warm.py:227-233: DataParallel(DDnet) + model_zoo/ddnet1.pth['state_dict']."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from .fastdvdnet_adapter import DataParallelLike
    from .network_demosaicking import DDnet
    from .synthetic import ddnet_synthetic_state_dict
    m = DataParallelLike(DDnet())
    path = os.path.join(root, 'model_zoo', 'ddnet1.pth')
    if os.path.exists(path):
        m.load_state_dict(torch.load(path)['state_dict'], strict=True)
    else:
        print('DDnet weights %s absent (as in the reference tree): using the synthetic init' % path)
        m.load_state_dict({'module.' + k: v for k, v in ddnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


def build_model(denoiser):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if denoiser == 'ffdnet_color':
        from .network_ffdnet import FFDNet
        m = FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode='R')                       # two_stage..FFD:218-225
        m.load_state_dict(torch.load(os.path.join(root, 'model_zoo', 'ffdnet_color.pth')), strict=True)
        return m.eval().cuda()
    from .fastdvdnet_adapter import DataParallelLike
    from .fastdvdnet_models import FastDVDnet
    from .synthetic import fastdvdnet_synthetic_state_dict
    m = DataParallelLike(FastDVDnet(num_input_frames=5))                                   # two_stage..FastDVD:234-241
    path = os.path.join(root, 'packages', 'fastdvdnet', 'model.pth')
    if os.path.exists(path):
        m.load_state_dict(torch.load(path), strict=True)
    else:
        print('FastDVDnet weights %s absent (as in the reference tree): using the synthetic init' % path)
        m.load_state_dict({'module.' + k: v for k, v in fastdvdnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


def main(denoiser):
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasetdir", default="./dataset/cacti/mid_scale")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--videos", type=int, default=6)
    ap.add_argument("--nmea", type=int, default=4)
    ap.add_argument("--no-update", action="store_true", help="plain PnP (update=False)")
    ap.add_argument("--deep-demosaicking", action="store_true", help="demosaic with DDnet (the reference scripts' default)")
    ap.add_argument("--share-weights", action="store_true", help="multi-GPU: one weight set, NCCL gradient all-reduce")
    args = ap.parse_args()
    ctx = parallel.init()
    worker_init_fn(0)
    update, reuse_model = not args.no_update, True
    table = FFD_TABLE if denoiser == 'ffdnet_color' else FASTDVD_TABLE
    resultsdir = "results/New1/" + str(int(time.time()))
    if ctx.rank == 0:
        mkdir(resultsdir + '/')
    f = open(resultsdir + '/log.txt', 'a') if ctx.rank == 0 else open(os.devnull, 'w')
    f.write('cacti midscale bayer: \n')
    average_psnr, average_ssim = [], []
    for datname in matio.VIDEOS[:args.videos]:
        sig255, iter_max, lr, update_per_iter, interval_iter, update_times = table[datname]
        if args.deep_demosaicking:
            sig255, iter_max, interval_iter = (FFD_DEEP if denoiser == 'ffdnet_color' else FASTDVD_DEEP)[datname]
        sigma = [s / 255 for s in sig255]
        f.write(datname + ':\n')
        meas_bayer, mask_bayer, orig_bayer = matio.load_video(args.datasetdir, datname, args.nmea,
                                                              force_synthetic=args.synthetic)
        recon_tv = matio.load_warm_start('./results/savedmat/', datname, mask_bayer.shape[2])
        nrows, ncols, nmea = meas_bayer.shape
        nmask = mask_bayer.shape[2]
        model_denoise = build_model(denoiser)
        model_demosaic = build_demosaicker() if args.deep_demosaicking else None
        MAXB = 255.
        results = {}
        for iframe in ctx.my_units(nmea):
            f.write('Measurement Frame {}.\n'.format(iframe))
            meas_t = meas_bayer[:, :, iframe] / MAXB
            orig_t = orig_bayer[:, :, iframe * nmask:(iframe + 1) * nmask] / MAXB
            v_tv = recon_tv[:, :, iframe * nmask:(iframe + 1) * nmask]
            begin = time.time()
            kw = dict(update_times=update_times) if denoiser == 'fastdvd_color' else {}
            out = reconstruct(meas_t, mask_bayer, 1, 0.01, denoiser, iter_max, False, sigma, x0_bayer=np2tch_cuda(v_tv),
                              X_orig=orig_t, model_denoise=model_denoise, model_demosaic=model_demosaic, show_iqa=True,
                              demosaic_method='malvar2004', lr_=lr, interval_iter=interval_iter, logf=f, update_=update,
                              update_per_iter=update_per_iter,
                              grad_sync=ctx.grad_sync if (args.share_weights and ctx.world > 1) else None, **kw)
            rgb, v, psnr_, ssim_, _, refined_model, _ = out
            if reuse_model and update:
                model_denoise = refined_model                                                # :270-275
            else:
                model_denoise = build_model(denoiser)
            msg = 'ADMM-{}--{}-{} PSNR {:2.2f} dB, SSIM {:.4f}, running time {:.1f} seconds.'.format(
                denoiser.upper(), datname, iframe, mean(psnr_), mean(ssim_), time.time() - begin)
            print(msg)
            f.write(msg + ' \n')
            results[iframe] = (v, np.asarray(psnr_, np.float32), np.asarray(ssim_, np.float32))
        results = ctx.gather_units(results)
        if ctx.rank == 0:
            v_all = np.concatenate([results[i][0] for i in range(nmea)], 2)
            psnr = np.concatenate([results[i][1] for i in range(nmea)]).reshape(-1, 1)
            ssim = np.concatenate([results[i][2] for i in range(nmea)]).reshape(-1, 1)
            print(round(float(psnr.mean()), 2), end=', ')
            print(round(float(ssim.mean()), 4))
            average_psnr.append(float(psnr.mean()))
            average_ssim.append(float(ssim.mean()))
            savedmatdir = resultsdir + '/savedmat/'
            os.makedirs(savedmatdir, exist_ok=True)
            tag = 'ffdnet' if denoiser == 'ffdnet_color' else 'fastdvd'
            sio.savemat('{}twoStageAdmm_{}_{}{:d}_sigma{:d}_all7_log.mat'.format(savedmatdir, denoiser.lower(), datname, nmask,
                                                                               int(sigma[-1] * MAXB)),
                        {'v_twoStageAdmm_%s_gray_bayer' % tag: v_all, 'psnr_%s_gray' % tag: psnr, 'ssim_%s_gray' % tag: ssim,
                         'meas_bayer': meas_bayer})
    if ctx.rank == 0 and average_psnr:
        print('all= ')
        print(round(mean(average_psnr), 2), end=', ')
        print(round(mean(average_ssim), 4))
    f.close()
    ctx.finalize()
