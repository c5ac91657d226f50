"""Image-quality reporting used by the solvers (PSNR from device-side SSE, SSIM on the device).

The reference obtains both from scikit-image on the host after a D2H copy in
every iteration (dvp_linear_inv_2_stage_ADMM_tensor_online.py:274-281,:318-321).
Here the squared-error sums are accumulated on the device (``sci_psnr_accum`` and
the fused kernels) and only scalars cross PCIe; SSIM (7x7 uniform window, sample
covariance, K1=.01, K2=.03, float64) is evaluated once per reconstruction by
``sci_ssim_accum`` on the device.  ``ssim`` below is the host
restatement kept for the tiled multi-GPU gather path and for cross-checks.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def psnr_from_sse(sse, n, data_range=1.0):
    return 10 * np.log10((data_range ** 2) / (np.asarray(sse, dtype=np.float64) / n))


def ssim(im1, im2, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    X = np.asarray(im1, dtype=np.float64)
    Y = np.asarray(im2, dtype=np.float64)
    npx = win_size ** X.ndim
    cov_norm = npx / (npx - 1)
    ux, uy = uniform_filter(X, size=win_size), uniform_filter(Y, size=win_size)
    vx = cov_norm * (uniform_filter(X * X, size=win_size) - ux * ux)
    vy = cov_norm * (uniform_filter(Y * Y, size=win_size) - uy * uy)
    vxy = cov_norm * (uniform_filter(X * Y, size=win_size) - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
