"""Host-side mirror of the reference's ``utilspy.py`` (sensing operators, seeding).

``A_`` / ``At_`` keep the reference signatures (utilspy.py:28-44) and accept the
same arbitrary-stride CUDA views (e.g. ``Phiall[..., ib]``); the arithmetic runs
in the sm_100a kernels ``sci_A`` / ``sci_At``.  Inside the solvers these
operators are not called one by one: they are fused into ``sci_project_stage1/2``.
"""
import math
import os

import numpy as np
import torch

from ._lib import SciError, call, ptr, stream


def mkdir(path):
    """utilspy.py:7-20."""
    path = path.strip().rstrip("\\")
    if not os.path.exists(path):
        os.makedirs(path)
        print(path + ' Successfully')
        return True
    print(path + ' Item existed')
    return False


def worker_init_fn(pid):
    """utilspy.py:22-25 — numpy / torch / CUDA seeds = 42 + pid."""
    np.random.seed(42 + pid)
    torch.manual_seed(42 + pid)
    torch.cuda.manual_seed(42 + pid)


def _check(*ts):
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32):
            raise SciError("A_/At_ expect float32 CUDA tensors")


def A_(x, Phi):
    """y[h,w] = sum_t x[h,w,t] * Phi[h,w,t]   (utilspy.py:28-33)."""
    _check(x, Phi)
    h, w, B = Phi.shape
    y = torch.empty((h, w), dtype=torch.float32, device=x.device)
    call("sci_A", ptr(x), *x.stride(), ptr(Phi), *Phi.stride(), ptr(y), *y.stride(), h, w, B, stream())
    return y


def At_(y, Phi):
    """x[h,w,t] = y[h,w] * Phi[h,w,t]   (utilspy.py:35-44)."""
    _check(y, Phi)
    h, w, B = Phi.shape
    x = torch.empty((h, w, B), dtype=torch.float32, device=y.device)
    call("sci_At", ptr(y), *y.stride(), ptr(Phi), *Phi.stride(), ptr(x), *x.stride(), h, w, B, stream())
    return x


def psnr(ref, img):
    """utilspy.py:46-54 (reporting helper; host scalar)."""
    mse = float(torch.mean((ref - img) ** 2))
    if mse == 0:
        return 100
    return 20 * math.log10(1.0 / math.sqrt(mse))
