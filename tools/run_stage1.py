"""Small driver used under ncu: one stage-1 reconstruction (config 1 size by default)."""
import io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import admm_denoise_bayer_demosaic_pre
from adaptivepnp_sci_b200.synthetic import make_case
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
its = int(sys.argv[2]) if len(sys.argv) > 2 else 10
meas, mask, orig = make_case(H, H, 8, 1001, bayer=True)
r = admm_denoise_bayer_demosaic_pre(meas, mask, 1, 0.01, 'tv', [its], False, [0], X_orig=orig, logf=io.StringIO())
torch.cuda.synchronize()
print("psnr", float(sum(r[1]) / len(r[1])))
