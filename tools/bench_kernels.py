"""Per-kernel timing of the HBM-bound kernels (CUDA events, L2 flushed between launches).

    python tools/bench_kernels.py [--sizes 512x512x8,2048x2048x24] [--json gpurun_out/kernels.json]

Reports achieved algorithmic GB/s (SURVEY.md §8(d) byte counts) against MEASURED_PEAKS.json.
Not the headline benchmark (that is bench.py); this is the per-kernel roofline evidence.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adaptivepnp_sci_b200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    times = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)          # > L2 write: evicts the working set
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e-3)
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="256x256x8,512x512x8,2048x2048x24")
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    pk, kind = peak_gbs()
    flush = None if a.no_flush else torch.zeros(64 * 1024 * 1024, device=dev)   # 256 MB > 126 MB L2
    rows = []
    for s in a.sizes.split(","):
        H, W, B = (int(v) for v in s.split("x"))
        cube, plane = H * W * B * 4, H * W * 4
        g = torch.Generator(device=dev).manual_seed(0)
        theta = torch.rand(B, H, W, device=dev, generator=g)
        b = 0.1 * torch.randn(B, H, W, device=dev, generator=g)
        phi = (torch.rand(B, H, W, device=dev, generator=g) > 0.5).float()
        y = (theta * phi).sum(0)
        phisum = phi.sum(0).clamp_(min=1)
        x = torch.empty_like(theta)
        b2 = torch.empty_like(b)
        w = torch.randn(B, 3, H, W, device=dev, generator=g)
        x_rgb = torch.empty_like(w)
        u = torch.empty_like(w)
        xhat = torch.rand(B, 3, H, W, device=dev, generator=g)
        ws = ops.TvWorkspace(H, W, B, dev)
        cases = [
            ("project_stage1", lambda: ops.project_stage1(theta, b, phi, y, phisum, x, 1.0, 0.01), 4 * cube + 2 * plane),
            ("project_stage2", lambda: ops.project_stage2(theta, b, phi, y, phisum, x, 1.0, 0.55), 4 * cube + 2 * plane),
            ("tv_chambolle(+clip+dual)", lambda: ops.tv_chambolle(x, b, -1.0, theta, b2, -1.0, True, ws), 4 * cube),
            ("malvar2004(+w/tau)", lambda: ops.malvar2004(x, b, 1.0, w, 0.01, x_rgb, u), 11 * cube),
            ("dual_update_rgb", lambda: ops.dual_update_rgb(xhat, x_rgb, w, x, b, theta, False), 13 * cube),
        ]
        for name, fn, nbytes in cases:
            med, best = timeit(fn, flush)
            rows.append(dict(kernel=name, size=s, bytes=nbytes, ms_median=med * 1e3, ms_best=best * 1e3,
                             gbs_median=nbytes / med / 1e9, frac_of_peak=nbytes / med / 1e9 / pk, peak_kind=kind))
            print("%-28s %-14s %9.1f MB  median %8.3f ms  best %8.3f ms  %8.1f GB/s  %5.1f%% of %s peak"
                  % (name, s, nbytes / 1e6, med * 1e3, best * 1e3, nbytes / med / 1e9, 100 * nbytes / med / 1e9 / pk, kind))
    if a.json:
        os.makedirs(os.path.dirname(a.json), exist_ok=True)
        json.dump(dict(peak_gbs=pk, peak_kind=kind, rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
