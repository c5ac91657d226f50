import io, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
from adaptivepnp_sci_b200 import synthetic
d = np.load("tests/golden/loops.npz")
meas, mask, orig = synthetic.make_case(64, 64, 8, 3000, bayer=True)
def mk():
    m = FFDNet(3, 3, 96, 12, 'R'); m.load_state_dict(torch.load("model_zoo/ffdnet_color.pth"), strict=True); return m.eval().cuda()
kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, logf=io.StringIO())
r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255],
                               x0_bayer=np2tch_cuda(d["s2_warm"]), X_orig=orig, model_denoise=mk(), model_demosaic=None, update_=True, update_per_iter=2, **kw)
print("wsplit env", os.environ.get("SCI_FFDNET_TRAIN_WSPLIT"), "max rgb", np.max(np.abs(r[0] - d["s2ffd_rgb"])), "max x", np.max(np.abs(r[1] - d["s2ffd_x"])))
