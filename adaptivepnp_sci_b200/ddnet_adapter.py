"""DDnet deep-demosaic plug-in.

Mirror of packages/DDnet/DDnet_test.py:166-321 (``ddnet_seqdenoise``, ``test_ddnet``): same names, arguments, layouts
and return convention.  The solvers call it as ``test_ddnet(oneCh2ThreeCh(x_bayer), yall, Phiall, model_demosaic)``
(dvp_linear_inv_2_stage_ADMM_tensor_online.py:192-194, :241-243) — inference only, ``args`` is never passed — and the
native solver uses ``demosaic_planar`` on the frame-planar mosaic directly (the sum over the three sparse colour planes
that DDnet forms first, network_demosaicking.py:411-416, IS the mosaic).

The optional self-supervised update (``args.dm_update``, DDnet_test.py:231-276; reachable only by calling ``test_ddnet``
directly with an ``args`` object) runs on the same engine: per step one training forward, the re-mosaicking loss
``MSE(vnoisy, sites(out))`` (:208-216, :268), the full backward (all convolutions and the three mixing tensors), and one
Adam step with a FRESH optimizer state (the reference constructs ``torch.optim.Adam`` inside the loop, :270).
"""
import torch

from . import ops
from ._lib import SciError, call, ptr, stream
from .network_demosaicking import DDnet

NUM_IN_FR_EXT = 5          # DDnet_test.py:16
last_losses = []


def _unwrap(model):
    m = model.module if hasattr(model, "module") else model
    if not isinstance(m, DDnet):
        raise SciError("ddnet adapter expects adaptivepnp_sci_b200.network_demosaicking.DDnet, got %s" % type(m).__name__)
    return m


def demosaic_planar(mosaic, model):
    """mosaic [B,H,W] planar device tensor -> demosaicked [B,3,H,W] (engine-owned buffer, valid until the next call).
    Frames whose size is not a multiple of 4 are reflect-padded and the result cropped, as ddnet_seqdenoise does
    (DDnet_test.py:180-196)."""
    padded, H, W = ops.pad_to_multiple(mosaic, 4)
    out = _unwrap(model).engine().forward(padded)
    return ops.crop_to(out, H, W)


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


def update_and_demosaic(planar, model, lr, update_per_iter):
    """planar [B,3,H,W] sparse-RGB input.  ``update_per_iter`` self-supervised steps, then the demosaicked sequence."""
    eng = _unwrap(model).engine()
    B, _, H, W = planar.shape
    mosaic = rgb_sum(planar)
    loss = torch.zeros(update_per_iter, dtype=torch.float64, device=planar.device)
    dout = eng.ws.get("dd_dout", (B, 3, H, W), planar.device)
    for it in range(update_per_iter):
        out = eng.forward(mosaic, train=True)                                         # :259-261
        call("sci_ddnet_loss_fwd_bwd", ptr(planar), ptr(out), ptr(dout), ptr(loss[it:it + 1]), B, H, W, stream())   # :266-268
        eng.backward(dout)                                                            # :272
        eng.bucket.new_optimizer()                                                    # :270 a fresh Adam every step
        eng.bucket.adam_step(lr)                                                      # :273
        eng.after_step()
    last_losses[:] = [loss]
    return eng.forward(mosaic, train=False)                                           # :279-282


def test_ddnet(vnoisy, yall, Phiall, model=None, useGPU=True, args=None, gray=False):
    """vnoisy [H,W,3,B] (sparse RGB mosaic, pixel-last as in the reference) -> demosaicked [H,W,3,B]
    (``(out, model)`` when ``args.dm_update``)."""
    if not useGPU:
        raise SciError("the B200 path has no CPU mode")
    if gray:
        raise NotImplementedError("DDnet demosaics Bayer mosaics; gray=True is not a path of the reference solvers")
    updata_ = bool(args is not None and args.dm_update)
    H, W, C, B = vnoisy.shape
    planar = ops.pixlast_to_planar(vnoisy.contiguous().float(), C, B).view(B, C, H, W)
    if updata_:
        out = update_and_demosaic(planar, model, args.dm_lr, args.dm_update_per_iter)
        for val in last_losses[0].cpu().numpy():
            print('ddn loss:', end=' ')
            print('tensor(%.4e)' % val)
    else:
        out = demosaic_planar(rgb_sum(planar), model)
    outv = ops.planar_to_pixlast(out, 3, B).view(H, W, 3, B)
    return (outv, model) if updata_ else outv
