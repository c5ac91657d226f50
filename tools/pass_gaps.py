"""How much of a FastDVDnet inference pass is NOT inside conv kernels (launch gaps, prologue/tail bubbles, packers)?"""
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
for _ in range(5):
    eng.forward(u, 12 / 255)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
a.record()
for _ in range(n):
    eng.forward(u, 12 / 255)
b.record()
torch.cuda.synchronize()
wall = a.elapsed_time(b) / n
eng.profile = []
eng.forward(u, 12 / 255)
torch.cuda.synchronize()
conv = sum(x.elapsed_time(y) for x, y, _, _ in eng.profile)
print("pass wall %.3f ms, sum of conv launches (event-bracketed) %.3f ms, %d conv launches" % (wall, conv, len(eng.profile)))
