"""Driver used under ncu: each HBM-bound kernel of the ADMM loop twice at 512x512x8 and at 2048x2048x24."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptivepnp_sci_b200 import ops
dev = torch.device("cuda:0")
for H, W, B in ((512, 512, 8), (2048, 2048, 24)):
    g = torch.Generator(device=dev).manual_seed(0)
    theta = torch.rand(B, H, W, device=dev, generator=g)
    b = 0.1 * torch.randn(B, H, W, device=dev, generator=g)
    phi = (torch.rand(B, H, W, device=dev, generator=g) > 0.5).float()
    y = (theta * phi).sum(0)
    phisum = phi.sum(0).clamp_(min=1)
    x, b2 = torch.empty_like(theta), torch.empty_like(b)
    w = torch.randn(B, 3, H, W, device=dev, generator=g)
    x_rgb, u = torch.empty_like(w), torch.empty_like(w)
    xhat = torch.rand(B, 3, H, W, device=dev, generator=g)
    ws = ops.TvWorkspace(H, W, B, dev)
    torch.cuda.synchronize()
    for _ in range(2):
        ops.project_stage1(theta, b, phi, y, phisum, x, 1.0, 0.01)
        ops.project_stage2(theta, b, phi, y, phisum, x, 1.0, 0.55)
        ops.tv_chambolle(x, b, -1.0, theta, b2, -1.0, True, ws)
        ops.malvar2004(x, b, 1.0, w, 0.01, x_rgb, u)
        ops.dual_update_rgb(xhat, x_rgb, w, x, b, theta, False)
    torch.cuda.synchronize()
print("ok")
