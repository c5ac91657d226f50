"""Oracle: the two plug-in denoiser adapters, incl. the online fine-tune.

Test infrastructure.  CPU restatements of
  * ``ffdnet_rgb_denoise_full_tensor``      packages/ffdnet/test_ffdnet_ipol.py:240-359
  * ``fastdvdnet_seqdenoise``               packages/fastdvdnet/fastdvdnet.py:82-146
  * ``fastdvdnet_denoiser_full_tensor_v2``  packages/fastdvdnet/test_fastdvdnet.py:325-500
  * ``ddnet_seqdenoise`` / ``test_ddnet``   packages/DDnet/DDnet_test.py:166-321
  * ``ffdnet_vdenoiser``                    packages/ffdnet/test_ffdnet_ipol.py:103-181   (frame-wise, gray)
  * ``fastdvdnet_denoiser``                 packages/fastdvdnet/test_fastdvdnet.py:149-235 (frame-wise, inference)
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .sci_ops import fourCh2OneCh, gen_bayer_img, rgb_to_bayer4

NUM_IN_FR_EXT = 5   # test_fastdvdnet.py:23


def _ffdnet_frames(x, sigma, model):
    """8 x model(img[1,3,H,W], sigma[1,1,1,1]) -> [H,W,3,B] (:266-273 / :344-354)."""
    outs = []
    for t in range(x.shape[3]):
        img = x[:, :, :, t].permute(2, 0, 1).float().unsqueeze(0)
        s = torch.full((1, 1, 1, 1), sigma).type_as(img)
        outs.append(model(img, s)[0].permute(1, 2, 0))
    return torch.stack(outs, dim=3)


def ffdnet_rgb_denoise_full_tensor(x, yall, Phiall, sigma, model, useGPU=True, lr_=1e-6,
                                   updata_=False, update_per_iter=4, losses=None):
    """x[H,W,3,B], yall[h,w,4], Phiall[h,w,B,4].  With ``updata_`` runs
    ``update_per_iter`` Adam steps on the measurement loss first (fresh optimizer
    per call, :251) and returns ``(outv, model)``."""
    if updata_:
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr_)
        mse = nn.MSELoss()
        for _ in range(update_per_iter):
            xb = _ffdnet_frames(x, sigma, model)
            xall = rgb_to_bayer4(xb)                                  # :275-278
            up_meas = torch.sum(xall * Phiall, dim=2)                 # :289
            loss = mse(up_meas, yall)                                 # :291
            opt.zero_grad()
            loss.backward()
            opt.step()
            if losses is not None:
                losses.append(float(loss.detach()))
        model.eval()
        with torch.no_grad():                                          # reference keeps autograd on (:303-315); values equal
            outv = _ffdnet_frames(x, sigma, model)
            if losses is not None:
                losses.append(float(mse(torch.sum(rgb_to_bayer4(outv) * Phiall, dim=2), yall)))
        return outv, model
    with torch.no_grad():
        return _ffdnet_frames(x, sigma, model)


def fastdvdnet_seqdenoise(seq, noise_std, windsize, model):
    """fastdvdnet.py:82-146 — circular temporal window, reflect pad to x4.  (The reference re-pads its noise map inside the
    frame loop, :129, and therefore fails from the second frame on when H or W is not a multiple of 4; padded once here.)"""
    N, C, H, W = seq.shape
    hw = (windsize - 1) // 2
    out = torch.empty((N, C, H, W))
    noise_map = noise_std.expand((1, 1, H, W))
    wpad, hpad = (-W) % 4, (-H) % 4
    if wpad or hpad:
        noise_map = F.pad(noise_map, (0, wpad, 0, hpad), mode="reflect")
    for f in range(N):
        idx = (torch.arange(f, f + windsize) - hw) % N                 # :115
        ns = seq[idx].reshape((1, -1, H, W))
        if wpad or hpad:
            ns = F.pad(ns, (0, wpad, 0, hpad), mode="reflect")
        den = model(ns, noise_map)
        out[f] = den[0, :, :H, :W]
    return out


def fastdvdnet_denoiser_full_tensor_v2(vnoisy, sigma, y_bayer=None, Phi=None, model=None, useGPU=True,
                                       lr_=1e-6, updata_=False, update_per_iter=1, gray=False,
                                       update_times=-1, losses=None, grad_hook=None, rng=None):
    """vnoisy[H,W,3,B]; y_bayer[h,w,4]; Phi[h,w,B,4]; ``model`` exposes ``.module``.

    Fine-tune quirk reproduced verbatim (test_fastdvdnet.py:359 with
    utils/utils_image.py:183-192): the helper returns ``meas + noise``, so the
    training input is ``vnoisy + float32(float64(vnoisy) + N(0,(5/255)^2))``
    = 2*vnoisy + noise, the noise drawn from the GLOBAL numpy RNG."""
    noisestd = torch.FloatTensor([sigma])
    if updata_:
        n_update_iter, lr_all = ([update_per_iter], [lr_]) if isinstance(update_per_iter, int) else (update_per_iter, lr_)
        mse = nn.MSELoss()
        v = vnoisy.permute(3, 2, 0, 1)                                   # [B,3,H,W]
        # ``rng`` / ``grad_hook`` exist for the MODIFIED oracle of the shared-weight configuration only (oracle/shared.py)
        noise = (rng or np.random).normal(0, 5 / 255, tuple(v.shape))    # utils_image.py:186
        v_plus = v + torch.from_numpy(v.detach().numpy() + noise).float()  # :359
        Phi_bayer = fourCh2OneCh(Phi)                                     # :362
        y_one = fourCh2OneCh(y_bayer)                                     # :363
        model.train()
        for m in model.module.modules():                                  # :376-379 BN frozen
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        for lr_i, nit in zip(lr_all, n_update_iter):
            opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=lr_i)
            for _ in range(nit):
                N, C, H, W = v.shape
                noise_map = noisestd.expand((1, 1, H, W))
                frames = []
                for f in range(N):
                    idx = (torch.arange(f, f + NUM_IN_FR_EXT) - 2) % N
                    frames.append(model(v_plus[idx].reshape((1, -1, H, W)), noise_map)[0])
                outv = torch.stack(frames, 0).permute(2, 3, 1, 0)          # [H,W,3,B]
                x_bayer = gen_bayer_img(outv, 1)                           # :428
                up_meas = torch.sum(x_bayer * Phi_bayer, dim=2)            # :430
                loss = mse(up_meas, y_one)                                 # :431
                opt.zero_grad()
                loss.backward()
                if grad_hook is not None:
                    grad_hook([p for p in model.parameters() if p.requires_grad])
                opt.step()
                if losses is not None:
                    losses.append(float(loss.detach()))
        with torch.no_grad():                                              # :453-458, on the CLEAN input
            outv = fastdvdnet_seqdenoise(v, noisestd, NUM_IN_FR_EXT, model).permute(2, 3, 1, 0)
            if losses is not None:
                losses.append(float(mse(torch.sum(rgb_to_bayer4(outv) * Phi, dim=2), y_bayer)))
        return outv, model
    model.eval()
    with torch.no_grad():
        v = vnoisy.permute(3, 2, 0, 1)
        return fastdvdnet_seqdenoise(v, noisestd, NUM_IN_FR_EXT, model).permute(2, 3, 1, 0)


def ffdnet_gray_denoise_full_tensor(x, y, Phi, sigma, model, lr_=1e-6, updata_=False, update_per_iter=4, losses=None):
    """DERIVED restatement (no reference function exists, SURVEY §8(c) "config without a reference function"):
    ``ffdnet_rgb_denoise_full_tensor`` (test_ffdnet_ipol.py:240-359) with the colour/Bayer handling removed —
    x [H,W,B] gray cube, frames denoised one by one with FFDNet-gray (two_stage_ADMM_Online_FFD_Warm.py:37-40 builds
    FFDNet(1,1,64,15,'R')), loss MSE(sum_t xhat_t * Phi_t, y) over H*W, fresh Adam per call."""
    def frames(m):
        outs = []
        for t in range(x.shape[2]):
            img = x[:, :, t].float()[None, None]
            s = torch.full((1, 1, 1, 1), sigma).type_as(img)
            outs.append(m(img, s)[0, 0])
        return torch.stack(outs, dim=2)
    if updata_:
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr_)
        mse = nn.MSELoss()
        for _ in range(update_per_iter):
            loss = mse(torch.sum(frames(model) * Phi, dim=2), y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if losses is not None:
                losses.append(float(loss.detach()))
        model.eval()
        with torch.no_grad():
            return frames(model), model
    with torch.no_grad():
        return frames(model)


# ---------------------------------------------------------------------------------------------------
# DDnet deep demosaic plug-in (packages/DDnet/DDnet_test.py)
# ---------------------------------------------------------------------------------------------------
def ddnet_seqdenoise(seq, windsize, model):
    """DDnet_test.py:166-204.  seq [N,C,H,W]: circular window of ``windsize`` frames around every frame, reflect
    pad (right/bottom) to multiples of 4, ``model(window[1, windsize*C, H, W])``, un-pad."""
    N, C, H, W = seq.shape
    hw = (windsize - 1) // 2
    wpad, hpad = (-W) % 4, (-H) % 4
    out = torch.empty((N, C, H, W))
    for f in range(N):
        idx = (torch.arange(f, f + windsize) - hw) % N
        win = F.pad(seq[idx].reshape(1, -1, H, W), (0, wpad, 0, hpad), mode='reflect')
        out[f] = model(win)[:, :, :H, :W]
    return out


def sparse_rgb_sites(rgb):
    """DDnet_test.py:208-216 ``gen_bayer_img``: [H,W,3,B] -> the RGGB sites only, as [B,3,H,W]."""
    s = torch.zeros_like(rgb)
    s[0::2, 0::2, 0, :] = rgb[0::2, 0::2, 0, :]
    s[0::2, 1::2, 1, :] = rgb[0::2, 1::2, 1, :]
    s[1::2, 0::2, 1, :] = rgb[1::2, 0::2, 1, :]
    s[1::2, 1::2, 2, :] = rgb[1::2, 1::2, 2, :]
    return s.permute(3, 2, 0, 1)


def test_ddnet(vnoisy, yall, Phiall, model=None, useGPU=True, args=None, gray=False, losses=None):
    """DDnet_test.py:218-321.  vnoisy [H,W,3,B] sparse RGB mosaic -> demosaicked [H,W,3,B].
    ``args.dm_update`` switches on the self-supervised fine-tune (loss = MSE between the input mosaic and the
    re-mosaicked output, a fresh Adam per step, :268-276); the solvers never pass ``args`` (dvp:193,243)."""
    updata_ = False
    if args is not None:
        lr_, update_per_iter, updata_ = args.dm_lr, args.dm_update_per_iter, args.dm_update
    if gray:
        vnoisy = vnoisy.unsqueeze(3)
    seq = vnoisy.permute(3, 2, 0, 1)
    if updata_:
        model.train()
        for _ in range(update_per_iter):
            outv = ddnet_seqdenoise(seq, NUM_IN_FR_EXT, model).permute(2, 3, 1, 0)
            if gray:
                outv = outv.squeeze(3)
            loss = nn.MSELoss()(seq, sparse_rgb_sites(outv))
            opt = torch.optim.Adam(model.parameters(), lr=lr_)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if losses is not None:
                losses.append(float(loss.detach()))
    else:
        model.eval()
    with torch.no_grad():
        outv = ddnet_seqdenoise(seq, NUM_IN_FR_EXT, model)
    outv = outv.permute(2, 3, 1, 0)
    if gray:
        outv = outv.squeeze(3)
    return (outv, model) if updata_ else outv


def ffdnet_vdenoiser(vnoisy, sigma, model):
    """test_ffdnet_ipol.py:103-181: numpy [M,N,F...] -> float64 numpy, frame by frame ``frame - model(frame, sigma)``
    (IPOL-flavour model = noise estimate), no clipping (:177)."""
    model.eval()
    vshape = vnoisy.shape
    v = vnoisy.reshape(*vshape[0:2], -1)
    outv = np.zeros(v.shape)
    with torch.no_grad():
        for k in range(v.shape[-1]):
            im = torch.Tensor(v[:, :, k][None, None])
            outv[:, :, k] = (im - model(im, torch.FloatTensor([sigma])))[0, 0].numpy()
    return outv.reshape(vshape)


def fastdvdnet_denoiser(vnoisy, sigma, model, gray=False):
    """test_fastdvdnet.py:149-235, inference branch (:207-231): numpy [H,W,F,3] (gray: [H,W,F]) -> same shape."""
    model.eval()
    v = torch.from_numpy(vnoisy).float()
    if gray:
        v = v.unsqueeze(3)
    v = v.permute(2, 3, 0, 1)
    with torch.no_grad():
        out = fastdvdnet_seqdenoise(v, torch.FloatTensor([sigma]), NUM_IN_FR_EXT, model)
    out = out.permute(2, 3, 0, 1)
    if gray:
        out = out.squeeze(3)
    return out.numpy()
