"""DDnet deep-demosaic plug-in.

Mirror of packages/DDnet/DDnet_test.py:166-321 (``ddnet_seqdenoise``, ``test_ddnet``): same names, arguments, layouts
and return convention.  The solvers call it as ``test_ddnet(oneCh2ThreeCh(x_bayer), yall, Phiall, model_demosaic)``
(dvp_linear_inv_2_stage_ADMM_tensor_online.py:192-194, :241-243) — inference only, ``args`` is never passed — and the
native solver uses ``demosaic_planar`` on the frame-planar mosaic directly (the sum over the three sparse colour planes
that DDnet forms first, network_demosaicking.py:411-416, IS the mosaic).

The optional self-supervised update (``args.dm_update``, DDnet_test.py:231-276), reachable only by calling
``test_ddnet`` directly with an ``args`` object, is not built: it raises ``NotImplementedError``.
"""
import torch

from . import ops
from ._lib import SciError, call, ptr, stream
from .network_demosaicking import DDnet

NUM_IN_FR_EXT = 5          # DDnet_test.py:16


def _unwrap(model):
    m = model.module if hasattr(model, "module") else model
    if not isinstance(m, DDnet):
        raise SciError("ddnet adapter expects adaptivepnp_sci_b200.network_demosaicking.DDnet, got %s" % type(m).__name__)
    return m


def demosaic_planar(mosaic, model):
    """mosaic [B,H,W] planar device tensor -> demosaicked [B,3,H,W] (engine-owned buffer, valid until the next call)."""
    return _unwrap(model).engine().forward(mosaic)


def rgb_sum(rgb):
    """[B,3,H,W] -> [B,H,W]: the channel sum DDnet applies to each input frame (network_demosaicking.py:411-416)."""
    B, _, H, W = rgb.shape
    out = torch.empty((B, H, W), dtype=torch.float32, device=rgb.device)
    call("sci_rgb_sum", ptr(rgb), ptr(out), H, W, B, stream())
    return out


def ddnet_seqdenoise(seq, windsize, model):
    """seq [N,3,H,W] -> [N,3,H,W]; circular ``windsize``-frame window around every frame (DDnet_test.py:166-204)."""
    if windsize != NUM_IN_FR_EXT:
        raise NotImplementedError("DDnet is a 5-frame model")
    seq = seq.contiguous().float()
    return demosaic_planar(rgb_sum(seq), model).clone()


def test_ddnet(vnoisy, yall, Phiall, model=None, useGPU=True, args=None, gray=False):
    """vnoisy [H,W,3,B] (sparse RGB mosaic, pixel-last as in the reference) -> demosaicked [H,W,3,B]."""
    if not useGPU:
        raise SciError("the B200 path has no CPU mode")
    if gray:
        raise NotImplementedError("DDnet demosaics Bayer mosaics; gray=True is not a path of the reference solvers")
    if args is not None and getattr(args, "dm_update", False):
        raise NotImplementedError("online update of the demosaicker (args.dm_update) is not built; the solvers never "
                                  "enable it (dvp_linear_inv_2_stage_ADMM_tensor_online.py:193,243 pass no args)")
    H, W, C, B = vnoisy.shape
    planar = ops.pixlast_to_planar(vnoisy.contiguous().float(), C, B).view(B, C, H, W)
    out = demosaic_planar(rgb_sum(planar), model)
    return ops.planar_to_pixlast(out, 3, B).view(H, W, 3, B)
