"""Host-side mirror of the reference's ``utils/utils_image.py`` Bayer helpers.

Same names, argument meaning and layouts (cubes ``[H,W,B]``, Bayer stacks
``[h,w,B,4]``, RGB cubes ``[H,W,3,B]``); the index remaps run in the sm_100a
kernels behind the C ABI and are bit-exact.  The solvers do not call these per
iteration (their state is frame-planar on the device); they exist so that code
written against the reference's helpers keeps working.
"""
import torch

from . import ops
from ._lib import call, ptr, require_cuda_f32, stream

bayer = [[0, 0], [0, 1], [1, 0], [1, 1]]   # utils_image.py:91


def np2tch_cuda(a):
    """utils_image.py:93."""
    return torch.from_numpy(a).cuda()


def cuda2np(a):
    """utils_image.py:95.  Large results are read back through a pinned staging tensor (one DMA at PCIe speed instead of
    the staged pageable path); the returned numpy array owns that buffer."""
    a = a.detach()
    if a.is_cuda and a.numel() * a.element_size() >= (1 << 20):
        host = torch.empty(a.shape, dtype=a.dtype, pin_memory=True)
        host.copy_(a, non_blocking=False)
        return host.numpy()
    return a.cpu().numpy()


def masks_CFA_Bayer_tensor(shape, pattern='RGGB'):
    """utils_image.py:106-112 — boolean CUDA masks (R, G, B) of the CFA."""
    pattern = pattern.upper()
    ch = {c: torch.zeros(tuple(shape), dtype=torch.bool, device='cuda') for c in 'RGB'}
    for c, (y, x) in zip(pattern, [(0, 0), (0, 1), (1, 0), (1, 1)]):
        ch[c][y::2, x::2] = True
    return tuple(ch[c] for c in 'RGB')


def gen_bayer_mask(R_m, G_m, B_m):
    """utils_image.py:115-118."""
    return torch.cat([R_m.unsqueeze(2), G_m.unsqueeze(2), B_m.unsqueeze(2)], dim=2)


def fourCh2OneCh(RGGB):
    """utils_image.py:130-143 — [h,w,4] -> [H,W]  or  [h,w,B,4] -> [H,W,B]."""
    RGGB = RGGB.contiguous()
    require_cuda_f32(RGGB)
    if RGGB.dim() == 3:
        h, w, _ = RGGB.shape
        out = torch.empty((2 * h, 2 * w), dtype=torch.float32, device=RGGB.device)
        B = 1
    else:
        h, w, B, _ = RGGB.shape
        out = torch.empty((2 * h, 2 * w, B), dtype=torch.float32, device=RGGB.device)
    call("sci_bayer4_to_mosaic", ptr(RGGB), ptr(out), h, w, B, stream())
    return out


def oneCh2FourCh(oneCh):
    """utils_image.py:145-151 — [H,W,B] -> [h,w,B,4]."""
    oneCh = oneCh.contiguous()
    require_cuda_f32(oneCh)
    H, W, B = oneCh.shape
    out = torch.empty((H // 2, W // 2, B, 4), dtype=torch.float32, device=oneCh.device)
    call("sci_mosaic_to_bayer4", ptr(oneCh), ptr(out), H // 2, W // 2, B, stream())
    return out


def oneCh2ThreeCh(oneCh):
    """utils_image.py:153-161 — sparse 3-channel mosaic [H,W,3,B]."""
    oneCh = oneCh.contiguous()
    H, W, B = oneCh.shape
    planar = ops.pixlast_to_planar(oneCh, 1, B).view(B, H, W)
    rgb = ops.bayer_to_rgb_sparse(planar)
    return ops.planar_to_pixlast(rgb, 3, B).view(H, W, 3, B)


def fourCh2ThreeCh(RGGB):
    """utils_image.py:162-171."""
    return oneCh2ThreeCh(fourCh2OneCh(RGGB))


def gen_bayer_img(RGB, output_ch=1):
    """packages/fastdvdnet/utils.py:69-78 — RGGB samples of [H,W,3,B]."""
    RGB = RGB.contiguous()
    H, W, _, B = RGB.shape
    planar = ops.pixlast_to_planar(RGB, 3, B).view(B, 3, H, W)
    mosaic = ops.planar_to_pixlast(ops.rgb_to_bayer(planar), 1, B).view(H, W, B)
    return mosaic if output_ch == 1 else oneCh2FourCh(mosaic)
