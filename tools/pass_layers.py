"""Per-layer CUDA-event timings of one FastDVDnet inference pass (8x512x512 by default) + the wall time of a pass.

    python tools/pass_layers.py [B H W]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 512, 512)
m = DataParallelLike(FastDVDnet())
m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()})
m = m.eval().cuda()
eng = m.module.engine()
u = torch.rand(B, 3, H, W, device="cuda")
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
acc = {}
reps = 5
for _ in range(reps):
    eng.profile = []
    eng.forward(u, 12 / 255)
    torch.cuda.synchronize()
    for i, (x, y, fl, tag) in enumerate(eng.profile):
        k = (i % (len(eng.profile) // 2), tag)
        t, f = acc.get(k, (0.0, fl))
        acc[k] = (t + x.elapsed_time(y), fl)
eng.profile = None
tot = 0.0
for (i, tag), (t, fl) in sorted(acc.items()):
    ms = t / (2 * reps)
    tot += ms
    print("  %-28s %8.4f ms %8.1f TFLOP/s" % (tag, ms, fl / ms / 1e9))
print("DenBlock sum %.3f ms; pass wall %.3f ms (%.1f TFLOP/s algorithmic over %d frames)" %
      (tot, wall, 2 * sum(f for _, f in acc.values()) / wall / 1e9, B))
