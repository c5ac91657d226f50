"""FFDNet-colour inference pass at 8x512x512: time + deviation from the fp32 FFMA engine, per inference form.

    python tools/ffdnet_pass.py            (SCI_FFDNET_INF=half | tf32)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200.network_ffdnet import FFDNet
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model(impl):
    os.environ["SCI_CONV_IMPL"] = impl
    m = FFDNet(3, 3, 96, 12, 'R')
    m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")))
    return m.eval().cuda()


u = torch.rand(8, 3, 512, 512, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
ref = model("ref").engine().forward(u, 12 / 255).clone()
eng = model("tc").engine()
out = eng.forward(u, 12 / 255).clone()
print("form %s: max-abs deviation from the fp32 engine %.3e (mean %.3e), output range [%.3f, %.3f]" %
      (os.environ.get("SCI_FFDNET_INF", "half"), float((out - ref).abs().max()), float((out - ref).abs().mean()), float(ref.min()), float(ref.max())))
for _ in range(3):
    eng.forward(u, 12 / 255)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    eng.forward(u, 12 / 255)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
flops = 2.0 * 8 * 256 * 256 * 9 * (15 * 96 + 10 * 96 * 96 + 96 * 12)
print("FFDNet-colour inference pass 8x512x512: %.3f ms = %.1f TFLOP/s algorithmic" % (ms, flops / ms / 1e9))
