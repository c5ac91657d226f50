"""Driver used under ncu: FastDVDnet inference passes only (8 frames 512x512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict
m = DataParallelLike(FastDVDnet())
m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()})
m = m.eval().cuda()
eng = m.module.engine()
u = torch.rand(8, 3, 512, 512, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng.forward(u, 12 / 255)
torch.cuda.synchronize()
print("ok")
