"""GPU: spatial row-strip tiling with halo exchange (BASELINE config 5 mechanism) is exact.

Two processes share the one test GPU (gloo transport through host memory; on a multi-GPU box the same code uses
NCCL point-to-point over NVLink), each owns half of the rows of a 128x64x4 Bayer frame; the tiled result must equal
the un-tiled reconstruction, including an online fine-tune step (loss normalised by the whole frame, gradients summed)."""
import io
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W, B = 128, 64, 4
KW = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=2, update_=True, update_per_iter=1)


def _model():
    from adaptivepnp_sci_b200.network_ffdnet import FFDNet
    m = FFDNet(3, 3, 96, 12, 'R')
    m.load_state_dict(torch.load(os.path.join(ROOT, "model_zoo", "ffdnet_color.pth")), strict=True)
    return m.eval().cuda()


def _case():
    from adaptivepnp_sci_b200.synthetic import make_case
    meas, mask, orig = make_case(H, W, B, 515, bayer=True)
    warm = np.clip(meas[:, :, None] * mask / np.maximum(mask.sum(2, keepdims=True), 1), 0, 1).astype(np.float32)
    return meas, mask, orig, warm


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      SCI_CONV_IMPL="ref")
    from adaptivepnp_sci_b200 import parallel
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
    ctx = parallel.init(backend="gloo")
    tile = parallel.TileContext(ctx, H, W)
    meas, mask, orig, warm = _case()
    m = _model()
    r = twoStageAdmm_denoise_bayer(tile.slice_rows(meas), tile.slice_rows(mask), 1, 0.01, 'ffdnet_color', [4], False,
                                   [25 / 255], x0_bayer=torch.from_numpy(tile.slice_rows(warm)).cuda(),
                                   X_orig=tile.slice_rows(orig), model_denoise=m, logf=io.StringIO(), tile=tile, **KW)
    q.put((rank, r[0], r[1], np.array(r[2]), np.array(r[4]), m.state_dict()["model.10.weight"].cpu().numpy()))
    ctx.finalize()


def _fdvd():
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict
    m = DataParallelLike(FastDVDnet())
    m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()}, strict=True)
    return m.eval().cuda()


FH, FW, FB = 352, 32, 3          # 2 strips of 176 rows: 80-row halo < strip, FastDVDnet sees only part of the neighbour


def _fcase():
    from adaptivepnp_sci_b200.synthetic import make_case
    meas, mask, orig = make_case(FH, FW, FB, 77, bayer=True)
    warm = np.clip(meas[:, :, None] * mask / np.maximum(mask.sum(2, keepdims=True), 1), 0, 1).astype(np.float32)
    return meas, mask, orig, warm


def _fworker(rank, world, port, q, impl="ref"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      SCI_CONV_IMPL=impl)
    from adaptivepnp_sci_b200 import parallel
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    ctx = parallel.init(backend="gloo")
    tile = parallel.TileContext(ctx, FH, FW)
    meas, mask, orig, warm = _fcase()
    m = _fdvd()
    worker_init_fn(0)
    r = twoStageAdmm_denoise_bayer(tile.slice_rows(meas), tile.slice_rows(mask), 1, 0.01, 'fastdvd_color', [4], False,
                                   [12 / 255], x0_bayer=torch.from_numpy(tile.slice_rows(warm)).cuda(),
                                   X_orig=tile.slice_rows(orig), model_denoise=m, logf=io.StringIO(), tile=tile,
                                   update_times=-1, **KW)
    q.put((rank, r[0], r[1], np.array(r[4]), m.state_dict()["module.temp2.inc.convblock.3.weight"].cpu().numpy(),
           tile.p2p is not None))
    ctx.finalize()


@pytest.mark.parametrize("impl", ["ref", "tc"])
def test_tiled_fastdvdnet_equals_untiled(cuda, monkeypatch, impl):
    """Strips + one 40-row P2P halo exchange per DenBlock (inference iterations) / 80-row overlap (the online update) equal
    the un-tiled run, on the fp32 engine and on the tensor-core (fp16 inference / TF32 training) engine."""
    monkeypatch.setenv("SCI_CONV_IMPL", impl)
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    meas, mask, orig, warm = _fcase()
    m = _fdvd()
    worker_init_fn(0)
    ref = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', [4], False, [12 / 255],
                                     x0_bayer=torch.from_numpy(warm).cuda(), X_orig=orig, model_denoise=m,
                                     logf=io.StringIO(), update_times=-1, **KW)
    w_ref = m.state_dict()["module.temp2.inc.convblock.3.weight"].cpu().numpy()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_fworker, args=(r, 2, port, q, impl)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the tensor-core engine computes every output pixel with the same operands in the same order whether tiled or not,
    # except for the fine-tune gradients (summed over strips in a different order)
    tol = 2e-5 if impl == "ref" else 2e-4
    for rank, rgb, xb, psnr_all, w, used_p2p in outs:
        assert used_p2p                                   # the boundary rows went through sci_halo_send / sci_halo_assemble
        assert np.max(np.abs(rgb - ref[0])) < tol and np.max(np.abs(xb - ref[1])) < tol
        assert np.max(np.abs(psnr_all - np.array(ref[4]))) < 1e-3
        assert np.max(np.abs(w - w_ref)) <= 2 * 2e-6 * 1.01 and np.mean(np.abs(w - w_ref)) < (0.1 if impl == "ref" else 0.5) * 2e-6


def test_tiled_equals_untiled(cuda, monkeypatch):
    monkeypatch.setenv("SCI_CONV_IMPL", "ref")          # fp32 engine: the comparison isolates the tiling logic
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
    meas, mask, orig, warm = _case()
    m = _model()
    ref = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'ffdnet_color', [4], False, [25 / 255],
                                     x0_bayer=torch.from_numpy(warm).cuda(), X_orig=orig, model_denoise=m,
                                     logf=io.StringIO(), **KW)
    w_ref = m.state_dict()["model.10.weight"].cpu().numpy()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rgb, xb, psnr_, psnr_all, w in outs:
        assert rgb.shape == (H, W, 3, B) and xb.shape == (H, W, B)         # full frame on every rank
        # strips of 64 rows + 28-row halo (FFDNet receptive field -24..+25): not the whole image, still exact
        assert np.max(np.abs(rgb - ref[0])) < 2e-5 and np.max(np.abs(xb - ref[1])) < 2e-5
        assert np.max(np.abs(psnr_all - np.array(ref[4]))) < 1e-3 and np.max(np.abs(psnr_ - np.array(ref[2]))) < 1e-3
        assert np.max(np.abs(w - w_ref)) <= 2 * 2e-6 * 1.01 and np.mean(np.abs(w - w_ref)) < 0.1 * 2e-6
