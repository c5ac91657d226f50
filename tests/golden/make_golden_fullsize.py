"""Headline-size (512x512x8) golden vectors FROM THE REFERENCE ITSELF.

Run in the build container only (needs the read-only tree at /root/reference):

    python tests/golden/make_golden_fullsize.py [config4] [config3]

BASELINE.json configs[3] (two-stage ADMM + online FastDVDnet, the headline) and configs[2] (FFDNet-colour +
Malvar) are run END TO END with the reference's own functions on the CPU - stage 1
(``admm_denoise_bayer_demosaic_pre``, 40 TV iterations, dvp...online.py:326-552) feeding stage 2
(``twoStageAdmm_denoise_bayer``, dvp...online.py:40-324) with the scripts' schedules
(two_stage_ADMM_Online_FastDVD_Warm.py:68-83, two_stage_ADMM_Online_FFD_Warm.py:71-80) - on
``make_case(512, 512, 8, seed=3000, bayer=True)``.

The full outputs are 8 MB (Bayer) + 25 MB (RGB) per config, too much for the repository, so what is committed is
  * a strided sample of the warm start and of the final Bayer reconstruction (every 3rd row and column - an odd
    stride visits all four Bayer phases) and of the final RGB reconstruction (every 6th),
  * float64 per-frame sums and sums of squares of the FULL arrays (a checksum every pixel contributes to),
  * ``psnr_all`` per iteration, final per-frame PSNR / SSIM.
``tests/test_gpu_fullsize.py`` compares the CUDA path (its OWN stage 1 + stage 2) with these at the north_star
tolerances (1e-3 max-abs, 0.05 dB), and ``bench.py`` reports ``delta_psnr`` against them.
"""
import io
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_harness, synthetic  # noqa: E402
import make_golden as mg                   # noqa: E402  (model builders shared with the small goldens)

torch.set_num_threads(os.cpu_count() or 1)
H, W, B, SEED, STRIDE, RGB_STRIDE = 512, 512, 8, 3000, 3, 6


def _sums(a, frame_axis):
    a = np.asarray(a, dtype=np.float64)
    ax = tuple(i for i in range(a.ndim) if i != frame_axis)
    return a.sum(axis=ax), (a * a).sum(axis=ax)


def _pack(prefix, x_bayer, x_rgb, out):
    out[prefix + "_x_s"] = np.ascontiguousarray(x_bayer[::STRIDE, ::STRIDE])
    out[prefix + "_x_sum"], out[prefix + "_x_sq"] = _sums(x_bayer, 2)
    if x_rgb is not None:
        out[prefix + "_rgb_s"] = np.ascontiguousarray(x_rgb[::RGB_STRIDE, ::RGB_STRIDE])
        out[prefix + "_rgb_sum"], out[prefix + "_rgb_sq"] = _sums(x_rgb, 3)


def warm_start(ns, meas, mask, orig):
    ns.utilspy.worker_init_fn(0)
    t0 = time.time()
    r = ns.dvp.admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None,
                                               X_orig=orig, model=None, show_iqa=True, logf=io.StringIO())
    print("  reference stage 1 (40 TV iterations): %.1f s, PSNR %.3f dB" % (time.time() - t0, float(np.mean(r[1]))))
    return r


def config4(ns, meas, mask, orig, warm, out):
    """configs[3]: FastDVDnet, sigma [12,6]/255, iterations [21,2], fine-tune every 9th iteration, 2 Adam steps."""
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=9, update_=True, update_per_iter=2,
              update_times=-1)
    ns.utilspy.worker_init_fn(0)
    t0 = time.time()
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [21, 2], False, [12 / 255, 6 / 255],
                                          x0_bayer=torch.from_numpy(warm), X_orig=orig, model_denoise=mg._ref_fastdvd(ns),
                                          model_demosaic=None, logf=io.StringIO(), **kw)
    print("  reference stage 2 FastDVDnet: %.1f s; psnr_all %s" % (time.time() - t0, np.round(np.array(r[4]), 3)))
    _pack("c4", r[1], r[0], out)
    out.update(c4_psnr_all=np.array(r[4]), c4_psnr=np.array(r[2]), c4_ssim=np.array(r[3]))


def config3(ns, meas, mask, orig, warm, out):
    """configs[2]: FFDNet-colour, sigma [25,12,6]/255, iterations [6,6,4], fine-tune every 6th iteration, 2 Adam steps."""
    kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=6, update_=True, update_per_iter=2)
    ns.utilspy.worker_init_fn(0)
    t0 = time.time()
    r = ns.dvp.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [6, 6, 4], False,
                                          [25 / 255, 12 / 255, 6 / 255], x0_bayer=torch.from_numpy(warm), X_orig=orig,
                                          model_denoise=mg._ref_ffdnet(ns), model_demosaic=None, logf=io.StringIO(), **kw)
    print("  reference stage 2 FFDNet: %.1f s; psnr_all %s" % (time.time() - t0, np.round(np.array(r[4]), 3)))
    _pack("c3", r[1], r[0], out)
    out.update(c3_psnr_all=np.array(r[4]), c3_psnr=np.array(r[2]), c3_ssim=np.array(r[3]))


if __name__ == "__main__":
    if not ref_harness.available():
        sys.exit("reference tree not present: golden vectors can only be regenerated in the build container")
    which = sys.argv[1:] or ["config4", "config3"]
    ns = ref_harness.load()
    path = os.path.join(HERE, "fullsize.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    meas, mask, orig = synthetic.make_case(H, W, B, SEED, bayer=True)
    r1 = warm_start(ns, meas, mask, orig)
    warm = r1[0]
    _pack("s1", warm, None, out)
    out.update(s1_psnr_all=np.array(r1[3]), s1_psnr=np.array(r1[1]), s1_ssim=np.array(r1[2]),
               shape=np.array([H, W, B, SEED, STRIDE]), rgb_stride=np.array(RGB_STRIDE))
    if "config4" in which:
        config4(ns, meas, mask, orig, warm, out)
    if "config3" in which:
        config3(ns, meas, mask, orig, warm, out)
    np.savez_compressed(path, **out)
    print("written", path, "%.2f MB" % (os.path.getsize(path) / 1e6))
