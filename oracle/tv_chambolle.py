"""Oracle: restatement of ``skimage.restoration.denoise_tv_chambolle``.

Test infrastructure.  **PARITY UNPINNED**: the arithmetic lives in
scikit-image (pinned 0.18.1 in the reference's readme.md:14; call sites
dvp_linear_inv_2_stage_ADMM_tensor_online.py:158-159 and :405-406 with
``weight=0.1, n_iter_max=5, multichannel=True``, default ``eps=2e-4``).
scikit-image is neither vendored under /root/reference nor installed here, so
this file restates the published algorithm (Chambolle 2004 projection as
implemented by ``_denoise_tv_chambolle_nd``) and is
itself the pin.

Energy accumulation: with the reference's pinned numpy 1.21 the energy scalar
``E`` is promoted to float64 by the ``weight * norm.sum()`` python-float
product, while the two array sums themselves are float32 pairwise sums.  The
restatement (and the CUDA kernel) computes the summands in float32 and
accumulates them in float64, which keeps the early-stop decision away from
summation-order noise.
"""
import numpy as np


def _tv_chambolle_2d(image, weight, eps, n_iter_max):
    """One 2-D channel.  Returns (out, n_updates) where n_updates is the index i
    at which the loop stopped (number of completed dual updates that shaped
    ``out`` is ``i``)."""
    f32 = np.float32
    h, w = image.shape
    p0 = np.zeros((h, w), f32)
    p1 = np.zeros((h, w), f32)
    g0 = np.zeros((h, w), f32)
    g1 = np.zeros((h, w), f32)
    d = np.zeros((h, w), f32)
    tau = f32(1.0 / (2.0 * 2))
    tw = f32(0.25 / float(weight))  # python-double tau/weight, cast on the fp32 multiply
    i = 0
    out = image
    E_init = E_prev = 0.0
    while i < n_iter_max:
        if i > 0:
            d = -(p0 + p1)
            d[1:, :] += p0[:-1, :]
            d[:, 1:] += p1[:, :-1]
            out = image + d
        else:
            out = image
        E = float(np.sum(d * d, dtype=np.float64))
        g0[:-1, :] = out[1:, :] - out[:-1, :]
        g1[:, :-1] = out[:, 1:] - out[:, :-1]
        norm = np.sqrt(g0 * g0 + g1 * g1)
        E += float(weight) * float(np.sum(norm, dtype=np.float64))
        norm = norm * tw + f32(1.0)
        p0 = (p0 - tau * g0) / norm
        p1 = (p1 - tau * g1) / norm
        E /= float(h * w)
        if i == 0:
            E_init = E
            E_prev = E
        else:
            if abs(E_prev - E) < eps * E_init:
                break
            E_prev = E
        i += 1
    return out, i


def denoise_tv_chambolle(image, weight=0.1, eps=2.0e-4, n_iter_max=200, multichannel=False,
                         return_stops=False):
    """float32 ``[h,w,C]`` (multichannel: each channel an independent 2-D image)
    or ``[h,w]``.  Float input is neither rescaled nor cast."""
    image = np.asarray(image)
    if image.dtype.kind != "f":
        image = image.astype(np.float64) / 255.0
    if not multichannel:
        out, n = _tv_chambolle_2d(image, weight, eps, n_iter_max)
        return (out, [n]) if return_stops else out
    out = np.zeros_like(image)
    stops = []
    for c in range(image.shape[-1]):
        o, n = _tv_chambolle_2d(np.ascontiguousarray(image[..., c]), weight, eps, n_iter_max)
        out[..., c] = o
        stops.append(n)
    return (out, stops) if return_stops else out
