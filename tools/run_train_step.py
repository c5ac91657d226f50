"""Driver used under ncu: one FastDVDnet fine-tune step (training forward, loss, backward, Adam) at 8x512x512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike, finetune_and_denoise
from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict, make_case
m = DataParallelLike(FastDVDnet())
m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()})
m = m.eval().cuda()
meas, mask, orig = make_case(512, 512, 8, 3000, bayer=True)
v = torch.from_numpy(orig).permute(2, 0, 1).unsqueeze(1).repeat(1, 3, 1, 1).contiguous().cuda()
phi = torch.from_numpy(mask).permute(2, 0, 1).contiguous().cuda()
y = torch.from_numpy(meas).cuda()
np.random.seed(0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    finetune_and_denoise(v, phi, y, 12 / 255, m, 2e-6, 1)
torch.cuda.synchronize()
print("ok")
