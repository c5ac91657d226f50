"""Oracle: sensing operators, Bayer remaps and the two Euclidean projections.

Test infrastructure (see ``oracle/__init__.py``).  CPU PyTorch, fp32, the
reference's channels-last layouts: cubes ``[h,w,B]``, Bayer-split cubes
``[h,w,B,4]`` (Bayer phase fastest), RGB cubes ``[H,W,3,B]``.
"""
import torch

from . import BAYER


def A_(x, Phi):
    """utilspy.py:28-33 — y = sum_t x_t * Phi_t over dim 2."""
    return torch.sum(x * Phi, dim=2)


def At_(y, Phi):
    """utilspy.py:35-44 — x_t = y * Phi_t (the reference materialises the
    broadcast with repeat_interleave; the products are identical)."""
    return y.unsqueeze(2) * Phi


def bayer_split_init(y_bayer, Phi_bayer, x0_bayer=None):
    """dvp_linear_inv_2_stage_ADMM_tensor_online.py:59-82 / :347-370.

    Returns yall[h,w,4], Phiall[h,w,B,4], Phi_sumall[h,w,4] (zeros -> 1),
    x0all[h,w,B,4] (At(y,Phi) or the warm-start slices).
    """
    H, W, B = Phi_bayer.shape
    h, w = H // 2, W // 2
    yall = torch.zeros(h, w, 4)
    Phiall = torch.zeros(h, w, B, 4)
    Phi_sumall = torch.zeros(h, w, 4)
    x0all = torch.zeros(h, w, B, 4)
    for ib, (b0, b1) in enumerate(BAYER):
        yall[..., ib] = y_bayer[b0::2, b1::2]
        Phiall[..., ib] = Phi_bayer[b0::2, b1::2]
        s = torch.sum(Phiall[..., ib], dim=2)
        s[s == 0] = 1
        Phi_sumall[..., ib] = s
        if x0_bayer is None:
            x0all[..., ib] = At_(yall[..., ib], Phiall[..., ib])
        else:
            x0all[..., ib] = x0_bayer[b0::2, b1::2]
    return yall, Phiall, Phi_sumall, x0all


def project_stage1(theta_all, ball, yall, Phiall, Phi_sumall, _lambda, gamma, out=None):
    """dvp...online.py:389-391 (GAP projection of the TV warm start).
    ``out`` may alias ``theta_all`` exactly as ``xall`` does at k=0 (:375-377)."""
    xall = torch.empty_like(theta_all) if out is None else out
    for ib in range(4):
        v = theta_all[..., ib] + ball[..., ib]
        yb = A_(v, Phiall[..., ib])
        xall[..., ib] = v + _lambda * At_((yall[..., ib] - yb) / (Phi_sumall[..., ib] + gamma), Phiall[..., ib])
    return xall


def project_stage2(theta_all, ball, yall, Phiall, Phi_sumall, alpha, rou, out=None):
    """dvp...online.py:128-140 (ADMM projection, stage 2)."""
    xall = torch.empty_like(theta_all) if out is None else out
    B = Phiall.shape[2]
    for ib in range(4):
        p = theta_all[..., ib] - (1 / rou) * ball[..., ib]
        yb = A_(p, Phiall[..., ib])
        t = (yall[..., ib] - yb) / (alpha * rou + Phi_sumall[..., ib])
        t = Phiall[..., ib] * torch.repeat_interleave(t.unsqueeze(2), B, dim=2)
        xall[..., ib] = p + t
    return xall


def bayer_merge(cube4):
    """dvp...online.py:170-172 — [h,w,B,4] -> [H,W,B] (also utils_image.py:130-143)."""
    h, w, B, _ = cube4.shape
    out = torch.zeros(2 * h, 2 * w, B)
    for ib, (b0, b1) in enumerate(BAYER):
        out[b0::2, b1::2] = cube4[..., ib]
    return out


def fourCh2OneCh(RGGB):
    """utils/utils_image.py:130-143 (3-D [h,w,4] and 4-D [h,w,B,4] inputs)."""
    if RGGB.dim() == 3:
        h, w = RGGB.shape[:2]
        one = torch.zeros(2 * h, 2 * w)
    else:
        h, w, B = RGGB.shape[:3]
        one = torch.zeros(2 * h, 2 * w, B)
    for ib, (b0, b1) in enumerate(BAYER):
        one[b0::2, b1::2] = RGGB[..., ib]
    return one


def oneCh2FourCh(oneCh):
    """utils/utils_image.py:145-151 — [H,W,B] -> [h,w,B,4]."""
    H, W, B = oneCh.shape
    out = torch.zeros(H // 2, W // 2, B, 4)
    for ib, (b0, b1) in enumerate(BAYER):
        out[..., ib] = oneCh[b0::2, b1::2]
    return out


def oneCh2ThreeCh(oneCh):
    """utils/utils_image.py:153-161 — sparse 3-channel mosaic [H,W,3,B]."""
    H, W, B = oneCh.shape
    RGB = torch.zeros(H, W, 3, B)
    RGB[0::2, 0::2, 0, :] = oneCh[0::2, 0::2, :]
    RGB[0::2, 1::2, 1, :] = oneCh[0::2, 1::2, :]
    RGB[1::2, 0::2, 1, :] = oneCh[1::2, 0::2, :]
    RGB[1::2, 1::2, 2, :] = oneCh[1::2, 1::2, :]
    return RGB


def fourCh2ThreeCh(RGGB):
    """utils/utils_image.py:162-171 — [h,w,B,4] Bayer planes -> sparse 3-channel mosaic [H,W,3,B]."""
    h, w, B, _ = RGGB.shape
    RGB = torch.zeros(2 * h, 2 * w, 3, B)
    RGB[0::2, 0::2, 0, :] = RGGB[:, :, :, 0]
    RGB[0::2, 1::2, 1, :] = RGGB[:, :, :, 1]
    RGB[1::2, 0::2, 1, :] = RGGB[:, :, :, 2]
    RGB[1::2, 1::2, 2, :] = RGGB[:, :, :, 3]
    return RGB


def rgb_to_bayer4(xrgb):
    """dvp...online.py:206-209 / test_ffdnet_ipol.py:275-278 — RGGB samples of an
    RGB cube [H,W,3,B] -> [h,w,B,4].  Nothing is averaged."""
    H, W, _, B = xrgb.shape
    out = torch.zeros(H // 2, W // 2, B, 4, dtype=xrgb.dtype)
    out[..., 0] = xrgb[0::2, 0::2, 0, :]
    out[..., 1] = xrgb[0::2, 1::2, 1, :]
    out[..., 2] = xrgb[1::2, 0::2, 1, :]
    out[..., 3] = xrgb[1::2, 1::2, 2, :]
    return out


def masks_CFA_Bayer_tensor(shape):
    """utils/utils_image.py:106-112 — RGGB boolean masks."""
    R = torch.zeros(shape, dtype=torch.bool)
    G = torch.zeros(shape, dtype=torch.bool)
    Bm = torch.zeros(shape, dtype=torch.bool)
    R[0::2, 0::2] = True
    G[0::2, 1::2] = True
    G[1::2, 0::2] = True
    Bm[1::2, 1::2] = True
    return R, G, Bm


def gen_bayer_img(RGB, output_ch=1):
    """packages/fastdvdnet/utils.py:69-78 — differentiable RGB -> Bayer by
    mask-multiply and sum over the colour axis; [H,W,3,B] -> [H,W,B] or [h,w,B,4]."""
    R, G, Bm = masks_CFA_Bayer_tensor((RGB.shape[0], RGB.shape[1]))
    mask = torch.stack([R, G, Bm], dim=2).unsqueeze(3)
    img = torch.sum(RGB * mask, dim=2)
    if output_ch == 1:
        return img
    return oneCh2FourCh_autograd(img)


def oneCh2FourCh_autograd(oneCh):
    return torch.stack([oneCh[b0::2, b1::2] for (b0, b1) in BAYER], dim=3)
