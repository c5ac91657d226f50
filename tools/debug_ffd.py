import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import np2tch_cuda, twoStageAdmm_denoise_bayer
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
from adaptivepnp_sci_b200.synthetic import make_case
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = np.load(os.path.join(ROOT, "tests/golden/loops.npz"))
def net():
    m = FFDNet(3, 3, 96, 12, 'R'); m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo/ffdnet_color.pth"))); return m.eval().cuda()
meas, mask, orig = make_case(64, 64, 8, 3000, bayer=True)
kw = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, logf=io.StringIO())
def stats(a, b, name):
    e = np.abs(a - b).ravel()
    print("%-28s max %.2e  p99.9 %.2e  mean %.2e  argmax %s" % (name, e.max(), np.percentile(e, 99.9), e.mean(), np.unravel_index(np.argmax(np.abs(a-b)), a.shape)))
r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [3], False, [25 / 255], x0_bayer=np2tch_cuda(d["s2_warm"]), X_orig=orig, model_denoise=net(), model_demosaic=None, update_=False, **kw)
stats(r[0], d["s2ffd0_rgb"], "inference-only rgb"); stats(r[1], d["s2ffd0_x"], "inference-only bayer")
r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4, 3], False, [25 / 255, 12 / 255], x0_bayer=np2tch_cuda(d["s2_warm"]), X_orig=orig, model_denoise=net(), model_demosaic=None, update_=True, update_per_iter=2, **kw)
stats(r[0], d["s2ffd_rgb"], "online rgb"); stats(r[1], d["s2ffd_x"], "online bayer")
print("psnr_all diff", np.abs(np.array(r[4]) - d["s2ffd_psnr_all"]).max())
# single forward pass check vs golden networks.npz
g = np.load(os.path.join(ROOT, "tests/golden/networks.npz"))
y = net()(torch.from_numpy(g["ffd_x"]).cuda(), torch.full((2, 1, 1, 1), 25 / 255).cuda()).cpu().numpy()
stats(y, g["ffd_y"], "single forward")
