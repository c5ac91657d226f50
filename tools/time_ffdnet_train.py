"""FFDNet-colour engine at 8x512x512: time of the training forward, the backward and the inference pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = FFDNet(3, 3, 96, 12, 'R'); m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth"))); m = m.eval().cuda()
eng = m.engine()
u = torch.rand(8, 3, 512, 512, device="cuda")
d = torch.rand(8, 3, 512, 512, device="cuda") * 1e-6


def t(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("wsplit(train layers):", eng.layers[1].wsplit)
print("inference pass   %.3f ms" % t(lambda: eng.forward(u, 25 / 255, train=False)))
print("training forward %.3f ms" % t(lambda: eng.forward(u, 25 / 255, train=True)))
eng.forward(u, 25 / 255, train=True)
print("backward         %.3f ms" % t(lambda: eng.backward(d)))
