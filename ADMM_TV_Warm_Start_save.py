#!/usr/bin/env python
"""Stage 1 entry point (drop-in for the reference's ADMM_TV_Warm_Start_save.py): ADMM-TV reconstruction of
every measurement group of the six mid-scale Bayer videos, saved as the warm start of stage 2.

Same flow, hyper-parameters (sigma=[0], 40 iterations, lambda=1, gamma=0.01; :36-37,:130-135), log lines and
output file (results/savedmat/_Admm_tv_<name>8.mat, key v_Admm_tv_denoise) as the reference; the solver runs on
the B200-native kernels.  Measurement groups are independent: with torchrun they are sharded over the ranks
(one process per GPU) and gathered on rank 0.
"""
import argparse
import os
import time
from statistics import mean

import numpy as np
import torch

from adaptivepnp_sci_b200 import matio, parallel
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import admm_denoise_bayer_demosaic_pre as reconstruct
from adaptivepnp_sci_b200.utilspy import mkdir, worker_init_fn


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasetdir", default="./dataset/cacti/mid_scale")
    ap.add_argument("--synthetic", action="store_true", help="use the deterministic synthetic videos")
    ap.add_argument("--videos", type=int, default=6)
    ap.add_argument("--nmea", type=int, default=4)
    ap.add_argument("--synthetic-size", default="512x512x8", help="HxWxB of the synthetic videos (tests use a small one)")
    args = ap.parse_args(argv)
    shape = tuple(int(v) for v in args.synthetic_size.split('x'))
    ctx = parallel.init()
    worker_init_fn(0)
    resultsdir = "results/New1/" + str(int(time.time()))
    if ctx.rank == 0:
        mkdir(resultsdir + '/')
    f = open(resultsdir + '/log.txt', 'a') if ctx.rank == 0 else open(os.devnull, 'w')
    f.write('cacti midscale bayer: \n')
    sigma, iter_max = [0 / 255], [40]
    average_psnr, average_ssim = [], []
    for datname in matio.VIDEOS[:args.videos]:
        f.write(datname + ':\n')
        meas_bayer, mask_bayer, orig_bayer = matio.load_video(args.datasetdir, datname, args.nmea, synthetic_shape=shape,
                                                              force_synthetic=args.synthetic)
        nrows, ncols, nmea = meas_bayer.shape
        nmask = mask_bayer.shape[2]
        MAXB = 255.
        results = {}
        for iframe in ctx.my_units(nmea):
            f.write('Measurement Frame {}.\n'.format(iframe))
            meas_t = meas_bayer[:, :, iframe] / MAXB
            orig_t = orig_bayer[:, :, iframe * nmask:(iframe + 1) * nmask] / MAXB
            begin = time.time()
            v, psnr_, ssim_, _ = reconstruct(meas_t, mask_bayer, 1, 0.01, 'tv', iter_max, False, sigma, x0_bayer=None,
                                             X_orig=orig_t, model=None, show_iqa=True, logf=f)
            msg = 'ADMM-{} PSNR {:2.2f} dB, SSIM {:.4f}, running time {:.1f} seconds.'.format(
                'TV', mean(psnr_), mean(ssim_), time.time() - begin)
            print(msg)
            f.write(msg + ' \n')
            results[iframe] = (v, np.asarray(psnr_, np.float32), np.asarray(ssim_, np.float32))
        results = ctx.gather_units(results)
        if ctx.rank == 0:
            v_all = np.concatenate([results[i][0] for i in range(nmea)], 2)
            psnr = np.concatenate([results[i][1] for i in range(nmea)]).reshape(-1, 1)
            ssim = np.concatenate([results[i][2] for i in range(nmea)]).reshape(-1, 1)
            p = matio.save_warm_start('./results/savedmat/', datname, nmask, v_all, psnr, ssim)
            print(p + ' -- saved ')
            average_psnr.append(float(psnr.mean()))
            average_ssim.append(float(ssim.mean()))
    if ctx.rank == 0:
        print('all= ')
        print(mean(average_psnr))
        print(mean(average_ssim))
    f.close()
    ctx.finalize()


if __name__ == "__main__":
    main()
