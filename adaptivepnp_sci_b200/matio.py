"""Minimal I/O for the scripts: the stage-1 -> stage-2 hand-off file and the dataset loader.

The reference reads MATLAB v7.3 datasets with h5py (ADMM_TV_Warm_Start_save.py:69-74) and hands the warm start
to stage 2 through ``results/savedmat/_Admm_tv_<name>8.mat`` (key ``v_Admm_tv_denoise``, [H,W,B*nmea] float32;
:174-178 <-> two_stage_ADMM_Online_FFD_Warm.py:171-176).  The hand-off file is written/read with scipy.io exactly
as in the reference.  Dataset files (MATLAB -v7.3 = HDF5) are read with h5py when it is installed and with the package's
own minimal HDF5 reader (``h5lite.py``) otherwise; no dataset ships with the reference (readme.md:22-23), so the scripts
fall back to the deterministic synthetic videos of ``synthetic.py`` when ``dataset/cacti/mid_scale/<name>.mat`` is absent
(``--synthetic`` forces it)."""
import os

import numpy as np
import scipy.io as sio

from .synthetic import make_case

VIDEOS = ['Beauty_bayer', 'Bosphorus_bayer', 'Jockey_bayer', 'Runner_bayer', 'ShakeNDry_bayer', 'Traffic_bayer']


def video_shape(datasetdir, datname, synthetic_shape=(512, 512, 8), force_synthetic=False):
    """(H, W, B) of a video without keeping its data (used to fast-forward the fine-tune noise stream over videos another
    rank owns)."""
    path = os.path.join(datasetdir, datname + '.mat')
    if not force_synthetic and os.path.exists(path):
        mask = load_video(datasetdir, datname)[1]
        return mask.shape
    return tuple(synthetic_shape)


def load_video(datasetdir, datname, nmea=4, synthetic_shape=(512, 512, 8), force_synthetic=False, with_orig_real=False):
    """Returns meas_bayer [H,W,nmea] (0..255 scale), mask_bayer [H,W,B], orig_bayer [H,W,B*nmea] (0..255 scale)
    [, orig_real: the file's 'orig' array as stored (ADMM_TV_Warm_Start_save.py:74), which the scripts copy into their
    result files; the synthetic videos have no RGB ground truth on file, their Bayer ground truth stands in]."""
    path = os.path.join(datasetdir, datname + '.mat')
    if not force_synthetic and os.path.exists(path):
        # MATLAB -v7.3 = HDF5 (ADMM_TV_Warm_Start_save.py:69-74 reads it with h5py).  h5py when it is installed, otherwise
        # the package's own reader of the HDF5 subset such files use (h5lite.py): same arrays, same (C-order) shapes
        try:
            import h5py
            f = h5py.File(path, 'r')
        except ImportError:
            from . import h5lite
            f = h5lite.File(path)
        meas = np.float32(np.array(f['meas_bayer']))
        mask = np.float32(np.array(f['mask_bayer'])).transpose((2, 1, 0))
        orig = np.float32(np.array(f['orig_bayer'])).transpose((2, 1, 0))
        orig_real = np.array(f['orig']) if 'orig' in f else orig
        if hasattr(f, "close"):
            f.close()
        meas = meas.transpose((1, 0))[:, :, None] if meas.ndim < 3 else meas.transpose((2, 1, 0))
        return (meas, mask, orig, orig_real) if with_orig_real else (meas, mask, orig)
    H, W, B = synthetic_shape
    vid = VIDEOS.index(datname) if datname in VIDEOS else 0
    meas_l, orig_l, mask = [], [], None
    for g in range(nmea):
        m, msk, o = make_case(H, W, B, 3000 + 10 * vid + g, bayer=True)
        mask = msk if mask is None else mask
        meas_l.append((o * mask).sum(2) * 255.0)
        orig_l.append(o * 255.0)
    meas, orig = np.stack(meas_l, 2).astype(np.float32), np.concatenate(orig_l, 2).astype(np.float32)
    return (meas, mask, orig, orig) if with_orig_real else (meas, mask, orig)


def warm_start_path(savedmatdir, datname, nmask):
    return '{}_Admm_{}_{}{:d}.mat'.format(savedmatdir, 'tv', datname, nmask)


def save_warm_start(savedmatdir, datname, nmask, v, psnr, ssim):
    os.makedirs(savedmatdir, exist_ok=True)
    p = warm_start_path(savedmatdir, datname, nmask)
    sio.savemat(p, {'v_Admm_tv_denoise': v, 'psnr_Admm_tv_denoise': psnr, 'ssim_Admm_tv_denoise': ssim})
    return p


def load_warm_start(savedmatdir, datname, nmask):
    return np.array(sio.loadmat(warm_start_path(savedmatdir, datname, nmask))['v_Admm_tv_denoise'], dtype=np.float32)
