"""Oracle: PSNR / SSIM as the reference obtains them from scikit-image.

Test infrastructure.  **PARITY UNPINNED** (scikit-image 0.18.1 is not in the
reference tree nor installed): restates ``peak_signal_noise_ratio(X, Y,
data_range=1.)`` and ``structural_similarity(X, Y, data_range=1.)`` with their
defaults, call sites dvp_linear_inv_2_stage_ADMM_tensor_online.py:279,320-321.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def compare_psnr(image_true, image_test, data_range=1.0):
    """fp32 difference and square, float64 mean (skimage.metrics.simple_metrics)."""
    a = np.asarray(image_true, dtype=np.float32)
    b = np.asarray(image_test, dtype=np.float32)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    return 10 * np.log10((data_range ** 2) / err)


def compare_ssim(im1, im2, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    """7x7 uniform window, sample covariance, float64, crop 3, mean."""
    X = np.asarray(im1, dtype=np.float64)
    Y = np.asarray(im2, dtype=np.float64)
    NP = win_size ** X.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(X, size=win_size)
    uy = uniform_filter(Y, size=win_size)
    uxx = uniform_filter(X * X, size=win_size)
    uyy = uniform_filter(Y * Y, size=win_size)
    uxy = uniform_filter(X * Y, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
