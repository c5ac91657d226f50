"""BASELINE config 4 semantics (what SCALE measures): two ranks, one measurement group each, ONE shared set of FastDVDnet
weights kept identical by a mean all-reduce of the gradient bucket - checked against the MODIFIED oracle of SURVEY 8(e)
(oracle/shared.py: the restated reference loop per group, gradients averaged over groups before every Adam step).

Two processes share the one test GPU (gloo carries the all-reduce; on a multi-GPU box the same ``grad_sync`` is NCCL)."""
import io
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
H, W, B = 64, 64, 8
ITERS, SIGMA = [5, 2], [12 / 255, 6 / 255]
KW = dict(show_iqa=True, demosaic_method='malvar2004', lr_=2e-6, interval_iter=3, update_=True, update_per_iter=2, update_times=-1)
KEYS = ["module.temp1.inc.convblock.0.weight", "module.temp2.outc.convblock.1.weight", "module.temp2.upc1.convblock.1.weight"]


def _case(g):
    from oracle import synthetic
    meas, mask, orig = synthetic.make_case(H, W, B, 3000 + g, bayer=True)
    warm = np.clip(meas[:, :, None] * mask / np.maximum(mask.sum(2, keepdims=True), 1), 0, 1).astype(np.float32)
    return meas, mask, warm, orig


def _worker(rank, world, port, q, impl):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      SCI_CONV_IMPL=impl)
    from adaptivepnp_sci_b200 import parallel
    from adaptivepnp_sci_b200.dvp_linear_inv_2_stage_ADMM_tensor_online import twoStageAdmm_denoise_bayer
    from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from adaptivepnp_sci_b200.synthetic import fastdvdnet_synthetic_state_dict
    from adaptivepnp_sci_b200.utilspy import worker_init_fn
    ctx = parallel.init(backend="gloo")
    ctx.check_even_split(world, "measurement groups")
    meas, mask, warm, orig = _case(rank)
    m = DataParallelLike(FastDVDnet())
    m.load_state_dict({"module." + k: v for k, v in fastdvdnet_synthetic_state_dict().items()}, strict=True)
    m = m.eval().cuda()
    worker_init_fn(0)
    r = twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, 'fastdvd_color', ITERS, False, SIGMA, x0_bayer=torch.from_numpy(warm).cuda(),
                                   X_orig=orig, model_denoise=m, logf=io.StringIO(), grad_sync=ctx.grad_sync, **KW)
    sd = m.state_dict()
    q.put((rank, r[0], r[1], np.array(r[4]), {k: sd[k].cpu().numpy() for k in KEYS}))
    ctx.finalize()


@pytest.mark.parametrize("impl", ["ref", "tc"])
def test_shared_weight_finetune_vs_modified_oracle(cuda, impl):
    from oracle import networks, shared, synthetic
    mo = networks.Wrapped(networks.FastDVDnet())
    sd0 = {"module." + k: v for k, v in synthetic.fastdvdnet_synthetic_state_dict().items()}
    mo.load_state_dict(sd0, strict=True)
    ref = shared.shared_weight_runs([_case(0), _case(1)], mo.eval(), 'fastdvd_color', ITERS, SIGMA, **KW)
    w_ref = [{k: r[5].state_dict()[k].numpy() for k in KEYS} for r in ref]
    for k in KEYS:                                                # the oracle's two copies took identical steps
        assert np.array_equal(w_ref[0][k], w_ref[1][k])
        assert np.max(np.abs(w_ref[0][k] - sd0[k].numpy())) > 1e-7        # ... and they did move
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q, impl)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tol = {"ref": 2e-4, "tc": 1e-3}[impl]
    for k in KEYS:                                                # identical weights on both ranks, bit for bit
        assert np.array_equal(outs[0][4][k], outs[1][4][k])
    for (rank, rgb, xb, psnr_all, w), r in zip(outs, ref):
        assert np.max(np.abs(rgb - r[0])) < tol and np.max(np.abs(xb - r[1])) < tol
        assert np.max(np.abs(psnr_all - np.array(r[4]))) < 0.05
        for k in KEYS:
            # 4 Adam steps of lr 2e-6: an update direction that flips under rounding moves a weight by at most 2 lr per step
            assert np.max(np.abs(w[k] - w_ref[rank][k])) <= 4 * 2 * 2e-6 * 1.01
            assert np.mean(np.abs(w[k] - w_ref[rank][k])) < (0.1 if impl == "ref" else 0.5) * 2e-6
