"""Driver used under ncu: one shortened stage-2 FastDVDnet reconstruction (10+1 iterations, one fine-tune at k=9)."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict, make_case
from adaptivepnp_sci_b200.utilspy import worker_init_fn
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [10, 1]
meas, mask, orig = make_case(H, H, 8, 3000, bayer=True)
warm = np.clip(meas[:, :, None] * mask / np.maximum(mask.sum(2, keepdims=True), 1), 0, 1).astype(np.float32)
m = DataParallelLike(FastDVDnet())
m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()})
m = m.eval().cuda()
worker_init_fn(0)
r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', iters, False, [12 / 255, 6 / 255][:len(iters)],
                               x0_bayer=torch.from_numpy(warm).cuda(), X_orig=orig, model_denoise=m, show_iqa=True,
                               lr_=2e-6, interval_iter=9, logf=io.StringIO(), update_=True, update_per_iter=2)
torch.cuda.synchronize()
print("psnr", float(np.mean(r[2])))
