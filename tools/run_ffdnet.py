"""BASELINE config 3 on one GPU: two-stage ADMM + online FFDNet-colour, 512x512x8 Bayer, schedule of
two_stage_ADMM_Online_FFD_Warm.py (sigma 25/12/6, iters 6/6/4, update every 6th iteration, 2 Adam steps, lr 2e-6).
Prints seconds per reconstruction (CUDA events, device-resident inputs) and the FFDNet pass / step times."""
import io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
from adaptivepnp_sci_b200.synthetic import make_case
from adaptivepnp_sci_b200.utilspy import worker_init_fn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
meas, mask, orig = make_case(H, H, 8, 3000, bayer=True)
warm = np.clip(meas[:, :, None] * mask / np.maximum(mask.sum(2, keepdims=True), 1), 0, 1).astype(np.float32)
sd = torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth"))


def model():
    m = FFDNet(3, 3, 96, 12, 'R'); m.load_state_dict(sd, strict=True)
    return m.eval().cuda()


def recon(m):
    return twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [6, 6, 4], False, [25 / 255, 12 / 255, 6 / 255],
                                      x0_bayer=torch.from_numpy(warm).cuda(), X_orig=None, model_denoise=m, show_iqa=False,
                                      lr_=2e-6, interval_iter=6, logf=io.StringIO(), update_=True, update_per_iter=2,
                                      return_device=True)


worker_init_fn(0)
for _ in range(2):
    recon(model())
ts = []
for _ in range(3):
    m = model()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); recon(m); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
eng = m.engine() if hasattr(m, "engine") else None
u = torch.rand(8, 3, H, H, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.forward(u, 25 / 255)
a.record()
for _ in range(10):
    eng.forward(u, 25 / 255)
b.record(); torch.cuda.synchronize()
out = {"size": [H, H, 8], "ms_per_recon": sorted(ts)[1], "iters": 16, "iters_per_sec": 16e3 / sorted(ts)[1],
       "ffdnet_inference_pass_ms": a.elapsed_time(b) / 10, "alg_gflop_per_pass": 111.55 * 8 * (H / 512) ** 2}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ffdnet_config3_%d.json" % H, "w"), indent=1)
