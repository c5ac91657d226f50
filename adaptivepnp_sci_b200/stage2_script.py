"""Shared body of the two stage-2 entry points (two_stage_ADMM_Online_FFD_Warm.py / ..._FastDVD_Warm.py).

Keeps the reference scripts' flow: per-video hyper-parameter tables (two_stage_ADMM_Online_FFD_Warm.py:68-151,
two_stage_ADMM_Online_FastDVD_Warm.py:66-166; ``--deep-demosaicking`` selects the scripts' ``deep_demosaicking=True``
columns and demosaics with DDnet, the default is the Malvar columns),
warm start from results/savedmat/_Admm_tv_<name>8.mat, loop over measurement groups with ``reuse_model``
carry-over, log lines and the result .mat (same keys as two_stage_ADMM_Online_FFD_Warm.py:320-330 /
two_stage_ADMM_Online_FastDVD_Warm.py:356-365).

Multi-GPU (torchrun), in order of fidelity to the reference:
* default: VIDEOS are dealt to the ranks.  The model is rebuilt per video (:218-241), so a video is the largest unit
  whose result does not depend on what ran before it - except for the fine-tune noise, which the FastDVDnet adapter
  draws from the global numpy stream: a rank therefore fast-forwards that stream over the videos it does not own
  (``skip_finetune_noise``: same draws, discarded), and every video's groups still run in order with the
  ``reuse_model`` carry-over (:270-275).  Results equal the one-GPU / reference run.
* ``--no-update``: nothing carries over between groups, so the (video, group) pairs are dealt to the ranks.
* ``--share-weights`` (BASELINE config 4; a SEMANTIC CHANGE): the groups of a video run in lock-step on
  the ranks, one set of denoiser weights kept identical by an NCCL mean all-reduce of the gradients.  Needs
  ``nmea % world == 0`` (a rank without a group would never join the all-reduce).
* ``--shard-groups``: the groups of a video are dealt to the ranks although the update carries weights over between
  them in the reference; every rank then starts from the pristine weights and from seed 42 - results DIFFER from
  the reference (a warning is printed)."""
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


def count_updates(iter_max, interval_iter, update_times, inital_iter=1):
    """Number of fine-tune calls one reconstruction makes (dvp...online.py:200,247)."""
    n, ui = 0, 0
    for k in range(int(sum(iter_max))):
        if k > inital_iter and k % interval_iter == 0 and (update_times < 0 or ui < update_times):
            n += 1
            ui += 1
    return n


def skip_finetune_noise(n_calls, shape):
    """Advance the global numpy RNG exactly as ``n_calls`` FastDVDnet fine-tune calls would (utils_image.py:186)."""
    from .fastdvdnet_adapter import _release_host_buffer, noise_stream
    for _ in range(n_calls):
        _release_host_buffer(noise_stream.get(shape))


def main(denoiser, argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasetdir", default="./dataset/cacti/mid_scale")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--synthetic-size", default="512x512x8", help="HxWxB of the synthetic videos (tests use a small one)")
    ap.add_argument("--videos", type=int, default=6)
    ap.add_argument("--nmea", type=int, default=4)
    ap.add_argument("--no-update", action="store_true", help="plain PnP (update=False)")
    ap.add_argument("--deep-demosaicking", action="store_true", help="demosaic with DDnet (the reference scripts' default)")
    ap.add_argument("--share-weights", action="store_true", help="multi-GPU: one weight set, NCCL gradient all-reduce")
    ap.add_argument("--shard-groups", action="store_true", help="multi-GPU: deal the groups of a video to the ranks even with "
                    "the online update (drops the reuse_model carry-over: results differ from the reference)")
    ap.add_argument("--resultsdir", default=None)
    args = ap.parse_args(argv)
    ctx = parallel.init()
    worker_init_fn(0)
    update, reuse_model = not args.no_update, True
    if args.share_weights and ctx.world > 1:
        ctx.check_even_split(args.nmea, "measurement groups")
    # what is dealt to the ranks (see the module docstring)
    by_group = ctx.world > 1 and (args.share_weights or args.shard_groups or not update)
    by_video = ctx.world > 1 and not by_group
    if args.shard_groups and update and not args.share_weights and ctx.world > 1 and ctx.rank == 0:
        print('WARNING: --shard-groups with the online update: every rank starts from the pristine weights and from seed 42; '
              'PSNR and the .mat output differ from a one-GPU / reference run (reuse_model carry-over dropped)')
    table = FFD_TABLE if denoiser == 'ffdnet_color' else FASTDVD_TABLE
    resultsdir = args.resultsdir or ("results/New1/" + str(int(time.time())))
    if ctx.rank == 0:
        mkdir(resultsdir + '/')
    ctx.barrier()
    f = open(resultsdir + '/log.txt', 'a') if (ctx.rank == 0 or by_video) else open(os.devnull, 'w')
    if ctx.rank == 0:
        f.write('cacti midscale bayer: \n')
    average_psnr, average_ssim = {}, {}
    shape = tuple(int(v) for v in args.synthetic_size.split('x'))
    for vi, datname in enumerate(matio.VIDEOS[:args.videos]):
        sig255, iter_max, lr, update_per_iter, interval_iter, update_times = table[datname]
        if args.deep_demosaicking:
            sig255, iter_max, interval_iter = (FFD_DEEP if denoiser == 'ffdnet_color' else FASTDVD_DEEP)[datname]
        sigma = [s / 255 for s in sig255]
        mine = (not by_video) or (vi % ctx.world == ctx.rank)
        if not mine:
            if denoiser == 'fastdvd_color' and update:
                # keep the global noise stream where the sequential run would have it after this video
                H_, W_, B_ = matio.video_shape(args.datasetdir, datname, shape, force_synthetic=args.synthetic)
                skip_finetune_noise(args.nmea * count_updates(iter_max, interval_iter, update_times), (B_, 3, H_, W_))
            continue
        f.write(datname + ':\n')
        meas_bayer, mask_bayer, orig_bayer, orig_real = matio.load_video(args.datasetdir, datname, args.nmea, synthetic_shape=shape,
                                                                         force_synthetic=args.synthetic, with_orig_real=True)
        recon_tv = matio.load_warm_start('./results/savedmat/', datname, mask_bayer.shape[2])
        nrows, ncols, nmea = meas_bayer.shape
        nmask = mask_bayer.shape[2]
        model_denoise = build_model(denoiser)
        model_demosaic = build_demosaicker() if args.deep_demosaicking else None
        MAXB = 255.
        results = {}
        for iframe in (ctx.my_units(nmea) if by_group else range(nmea)):
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
            rgb, v, psnr_, ssim_, psnr_all_t, refined_model, _ = out
            if reuse_model and update:
                model_denoise = refined_model                                                # :270-275
            else:
                model_denoise = build_model(denoiser)
            msg = 'ADMM-{}--{}-{} PSNR {:2.2f} dB, SSIM {:.4f}, running time {:.1f} seconds.'.format(
                denoiser.upper(), datname, iframe, mean(psnr_), mean(ssim_), time.time() - begin)
            print(msg)
            f.write(msg + ' \n')
            results[iframe] = (v, np.asarray(psnr_, np.float32), np.asarray(ssim_, np.float32), np.asarray(psnr_all_t, np.float64))
        if by_group:
            results = ctx.gather_units(results)
        if ctx.rank == 0 or by_video:
            v_all = np.concatenate([results[i][0] for i in range(nmea)], 2)
            psnr = np.concatenate([results[i][1] for i in range(nmea)]).reshape(-1, 1)
            ssim = np.concatenate([results[i][2] for i in range(nmea)]).reshape(-1, 1)
            print(round(float(psnr.mean()), 2), end=', ')
            print(round(float(ssim.mean()), 4))
            average_psnr[vi] = float(psnr.mean())
            average_ssim[vi] = float(ssim.mean())
            savedmatdir = resultsdir + '/savedmat/'
            os.makedirs(savedmatdir, exist_ok=True)
            # keys of the reference's result file (two_stage_ADMM_Online_FFD_Warm.py:320-330, ..._FastDVD_Warm.py:356-365)
            tag = 'ffd' if denoiser == 'ffdnet_color' else 'fastdvd'
            mat = {'v_twoStageAdmm_%s_gray_bayer' % tag: v_all, 'psnr_%s_gray' % tag: psnr, 'ssim_%s_gray' % tag: ssim,
                   'orig_real': orig_real, 'meas_bayer': meas_bayer}
            if denoiser == 'ffdnet_color':
                mat['psnr_all_iter'] = [results[i][3] for i in range(nmea)]                 # FFD script only (:327)
            sio.savemat('{}twoStageAdmm_{}_{}{:d}_sigma{:d}_all7_log.mat'.format(savedmatdir, denoiser.lower(), datname, nmask,
                                                                               int(sigma[-1] * MAXB)), mat)
    if by_video:
        average_psnr = ctx.gather_units(average_psnr)
        average_ssim = ctx.gather_units(average_ssim)
    if ctx.rank == 0 and average_psnr:
        print('all= ')
        print(round(mean(average_psnr.values()), 2), end=', ')
        print(round(mean(average_ssim.values()), 4))
    f.close()
    ctx.finalize()
    return resultsdir
