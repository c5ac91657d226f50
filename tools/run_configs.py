"""Per-config timing of BASELINE.json configs 1-3 on ONE GPU (configs 4 and 5 are bench.py and tools/run_config5.py).

Device time with CUDA events around the public solver call with device-resident outputs where the API allows it
(``return_device=True``), otherwise end to end (numpy in / numpy out).  (CPU timings of the reference's port come from
``bench.py --impl reference`` / the ``cpu_baseline`` leg — nothing outside tests/, smoke() and bench.py touches oracle/.)"""
import argparse, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import (admm_denoise_bayer_demosaic_pre,
                                                                            twoStageAdmm_denoise_bayer, twoStageAdmm_denoise_gray)
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
from adaptivepnp_sci_b200.synthetic import make_case
from adaptivepnp_sci_b200.utilspy import worker_init_fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    argparse.ArgumentParser().parse_args()
    out = {}
    worker_init_fn(0)
    # config 1: TV warm start, 256x256x8 gray cube, 40 iterations (ADMM_TV_Warm_Start_save.py:36-37,132-135)
    meas, mask, orig = make_case(256, 256, 8, 1001, bayer=False)
    run1 = lambda: admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None, X_orig=orig,
                                                   show_iqa=True, logf=io.StringIO())
    ms = timed(run1)
    r = run1()
    out["config1_tv_256x256x8"] = {"ms_per_recon_e2e": ms, "iters": 40, "iters_per_sec": 40e3 / ms, "psnr_db": float(np.mean(r[1]))}
    warm1 = r[0]
    # config 2: two-stage + online FFDNet-gray on the same cube (derived loop); sigma 25/12/6, iters 6/6/4
    sd = torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_gray.pth"))

    def gray_model():
        m = FFDNet(1, 1, 64, 15, 'R'); m.load_state_dict(sd, strict=True)
        return m.eval().cuda()
    kw = dict(iter_max=[6, 6, 4], sigma=[25 / 255, 12 / 255, 6 / 255], X_orig=orig, lr_=2e-6, interval_iter=6, update_=True,
              update_per_iter=2)
    run2 = lambda: twoStageAdmm_denoise_gray(meas, mask, 'ffdnet_gray', x0=torch.from_numpy(warm1).cuda(), model_denoise=gray_model(), **kw)
    ms = timed(run2)
    r = run2()
    out["config2_ffdnet_gray_256x256x8"] = {"ms_per_recon_e2e": ms, "iters": 16, "iters_per_sec": 16e3 / ms,
                                            "psnr_db": float(np.mean(r[2]))}
    # config 3: mid-scale Bayer 512x512x8, FFDNet-colour + Malvar, one measurement group
    meas, mask, orig = make_case(512, 512, 8, 3000, bayer=True)
    w3 = admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], x0_bayer=None, X_orig=orig, show_iqa=False,
                                         logf=io.StringIO())[0]
    sdc = torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth"))

    def color_model():
        m = FFDNet(3, 3, 96, 12, 'R'); m.load_state_dict(sdc, strict=True)
        return m.eval().cuda()
    run3 = lambda: twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [6, 6, 4], False, [25 / 255, 12 / 255, 6 / 255],
                                              x0_bayer=torch.from_numpy(w3).cuda(), X_orig=orig, model_denoise=color_model(),
                                              show_iqa=True, lr_=2e-6, interval_iter=6, logf=io.StringIO(), update_=True,
                                              update_per_iter=2)
    ms = timed(run3)
    r = run3()
    out["config3_ffdnet_color_512x512x8"] = {"ms_per_recon_e2e": ms, "iters": 16, "iters_per_sec": 16e3 / ms,
                                             "psnr_db": float(np.mean(r[2])), "psnr_warm_start_db": None}
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs_1_3.json", "w"), indent=1)


main()
