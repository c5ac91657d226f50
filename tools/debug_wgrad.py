import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200 import engine
from adaptivepnp_sci_b200._lib import call, stream
dev = torch.device('cuda')
def run(x, dz, impl, N, H, W, Ci, Co, stride=1):
    dw = torch.zeros(9, Co, Ci, device=dev)
    d = engine.WgradDesc(x.data_ptr(), dz.data_ptr(), None, dw.data_ptr(), N, H, W, Ci, Co, stride)
    call("sci_conv3x3_wgrad", ctypes.byref(d), impl, stream())
    torch.cuda.synchronize()
    return dw
for (N, H, W, Ci, Co, st) in [(1, 16, 16, 32, 32, 1), (2, 20, 28, 96, 96, 1), (2, 24, 32, 32, 64, 2), (1, 8, 12, 128, 256, 1)]:
    x = torch.randn(N, H, W, Ci, device=dev); dz = torch.randn(N, (H - 1) // st + 1, (W - 1) // st + 1, Co, device=dev)
    r, t = run(x, dz, 1, N, H, W, Ci, Co, st), run(x, dz, 0, N, H, W, Ci, Co, st)
    print((N, H, W, Ci, Co, st), "max|ref|", float(r.abs().max()), "max|tc|", float(t.abs().max()), "rel err", float((r - t).abs().max() / r.abs().max()))
