"""Is the headline step host-bound?  Host enqueue time (perf_counter, no sync) vs GPU time (events) of one reconstruction."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import admm_denoise_bayer_demosaic_pre, twoStageAdmm_denoise_bayer
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict, make_case
from adaptivepnp_sci_b200.utilspy import worker_init_fn
meas, mask, orig = make_case(512, 512, 8, 3000, bayer=True)
warm = admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [40], False, [0], X_orig=None, show_iqa=False)[0]
m = DataParallelLike(FastDVDnet()); m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()}); m = m.eval().cuda()
d = [torch.from_numpy(a).cuda() for a in (meas, mask, warm)]
worker_init_fn(0)
for upd in (True, False):
    def step():
        return twoStageAdmm_denoise_bayer(d[0], d[1], 1, 0.01, 'fastdvd_color', [21, 2], False, [12 / 255, 6 / 255], x0_bayer=d[2], X_orig=None,
                                          show_iqa=False, model_denoise=m, lr_=2e-6, interval_iter=9, update_=upd, update_per_iter=2,
                                          update_times=-1, return_device=True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    hs, gs = [], []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); a.record(); step(); b.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        hs.append((t1 - t0) * 1e3); gs.append(a.elapsed_time(b))
    print("update=%s: host enqueue %.1f ms, GPU %.1f ms per reconstruction" % (upd, np.median(hs), np.median(gs)))
