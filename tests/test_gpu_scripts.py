"""The three script entry points end to end on the GPU: ADMM_TV_Warm_Start_save -> hand-off .mat ->
two_stage_ADMM_Online_FFD_Warm / ..._FastDVD_Warm, on one small synthetic video with one measurement group; the stage-2
results are checked against the oracle's restatement of the reference loop on the same inputs, and the result .mat files
against the reference's key names (two_stage_ADMM_Online_FFD_Warm.py:320-330, ..._FastDVD_Warm.py:356-365)."""
import glob
import io
import os
import sys

import numpy as np
import pytest
import scipy.io as sio
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("scripts")
    old = os.getcwd()
    os.chdir(d)
    yield d
    os.chdir(old)


ARGS = ["--synthetic", "--synthetic-size", "64x64x8", "--videos", "1", "--nmea", "1"]


def test_scripts_end_to_end(cuda, workdir, monkeypatch):
    monkeypatch.setenv("SCI_CONV_IMPL", "ref")            # fp32 engine: the comparison below is with the fp32 oracle
    sys.path.insert(0, ROOT)
    import ADMM_TV_Warm_Start_save as s1
    from adaptivepnp_sci_b200 import matio, stage2_script
    from oracle import admm, networks, synthetic
    s1.main(ARGS)
    warm_file = matio.warm_start_path('./results/savedmat/', 'Beauty_bayer', 8)
    assert os.path.exists(warm_file)
    warm = matio.load_warm_start('./results/savedmat/', 'Beauty_bayer', 8)
    meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
    # the script scales by 255 and back (ADMM_TV_Warm_Start_save.py:118-121): compare with the oracle on the same path
    meas255 = (orig * mask).sum(2) * 255.0
    ref1 = admm.admm_denoise_bayer_demosaic_pre(np.float32(meas255) / 255., mask, 1, 0.01, 'tv', [40], False, [0],
                                                X_orig=np.float32(orig * 255.0) / 255.)
    assert warm.shape == (64, 64, 8) and np.max(np.abs(warm - ref1[0])) < 2e-5

    # ---- stage 2, FFDNet (Beauty_bayer row of the table: sigma 25/12/6, iterations 15/6/4, interval 15)
    out = stage2_script.main('ffdnet_color', ARGS + ["--resultsdir", "results/ffd"])
    files = glob.glob(os.path.join(out, "savedmat", "twoStageAdmm_ffdnet_color_Beauty_bayer8_sigma6_all7_log.mat"))
    assert len(files) == 1
    m = sio.loadmat(files[0])
    for key in ("v_twoStageAdmm_ffd_gray_bayer", "psnr_ffd_gray", "ssim_ffd_gray", "psnr_all_iter", "orig_real", "meas_bayer"):
        assert key in m, key
    assert m["v_twoStageAdmm_ffd_gray_bayer"].shape == (64, 64, 8) and m["psnr_ffd_gray"].shape == (8, 1)
    assert np.asarray(m["psnr_all_iter"]).shape[-1] == 25
    mo = networks.FFDNet(3, 3, 96, 12, 'R')
    mo.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")), strict=True)
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    worker_init_fn(0)
    r = admm.twoStageAdmm_denoise_bayer(np.float32(meas255) / 255., mask, 1, 0.01, 'ffdnet_color', [15, 6, 4], False,
                                        [25 / 255, 12 / 255, 6 / 255], x0_bayer=torch.from_numpy(warm), X_orig=np.float32(orig * 255.0) / 255.,
                                        model_denoise=mo.eval(), show_iqa=True, lr_=2e-6, interval_iter=15, update_=True, update_per_iter=2)
    assert np.max(np.abs(m["v_twoStageAdmm_ffd_gray_bayer"] - r[1])) < 2e-4
    assert np.max(np.abs(m["psnr_ffd_gray"].ravel() - np.asarray(r[2]))) < 1e-2
    log = open(os.path.join(out, "log.txt")).read()
    assert "Measurement Frame 0." in log and "ADMM-FFDNET_COLOR--Beauty_bayer-0 PSNR" in log

    # ---- stage 2, FastDVDnet (Beauty_bayer: sigma 8, 18 iterations, interval 9, update_times 1)
    out = stage2_script.main('fastdvd_color', ARGS + ["--resultsdir", "results/fdvd"])
    files = glob.glob(os.path.join(out, "savedmat", "twoStageAdmm_fastdvd_color_Beauty_bayer8_sigma8_all7_log.mat"))
    assert len(files) == 1
    m = sio.loadmat(files[0])
    for key in ("v_twoStageAdmm_fastdvd_gray_bayer", "psnr_fastdvd_gray", "ssim_fastdvd_gray", "orig_real", "meas_bayer"):
        assert key in m, key
    assert "psnr_all_iter" not in m                       # the FastDVD script does not store it (:356-365)
    mo = networks.Wrapped(networks.FastDVDnet())
    mo.load_state_dict({"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}, strict=True)
    worker_init_fn(0)
    r = admm.twoStageAdmm_denoise_bayer(np.float32(meas255) / 255., mask, 1, 0.01, 'fastdvd_color', [18], False, [8 / 255],
                                        x0_bayer=torch.from_numpy(warm), X_orig=np.float32(orig * 255.0) / 255., model_denoise=mo.eval(),
                                        show_iqa=True, lr_=2e-6, interval_iter=9, update_=True, update_per_iter=2, update_times=1)
    assert np.max(np.abs(m["v_twoStageAdmm_fastdvd_gray_bayer"] - r[1])) < 2e-4
