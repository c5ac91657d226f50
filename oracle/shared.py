"""MODIFIED oracle for BASELINE config 4: measurement groups fine-tune ONE shared set of denoiser weights.

Test infrastructure.  The reference fine-tunes sequentially, group after group, with the weights carried over
(two_stage_ADMM_Online_FFD_Warm.py:241-275, ``reuse_model``).  The multi-GPU configuration the benchmark measures runs
the groups in lock-step instead, every Adam step taken on the gradient of the loss AVERAGED over the groups
 - a semantic change, so its oracle is a modification of the restated reference loop:

* each group runs ``oracle.admm.twoStageAdmm_denoise_bayer`` unchanged, on its own thread, with its own copy of the model
  (identical initial weights) and its own ``RandomState(42)`` noise stream - exactly what a rank of the product does
  (every rank seeds 42, utilspy.py:22-25);
* the only coupling is ``grad_hook``: after ``loss.backward()`` and before ``optimizer.step()`` the threads meet at a
  barrier and replace their gradients by the mean over groups (what NCCL's AVG all-reduce of the flat bucket does), so
  all copies take identical Adam steps.
"""
import copy
import threading

import numpy as np
import torch


class _MeanGrad:
    def __init__(self, n):
        self.n = n
        self.barrier = threading.Barrier(n)
        self.slots = [None] * n

    def hook(self, rank):
        def fn(params):
            self.slots[rank] = [p.grad.detach().clone() for p in params]
            self.barrier.wait()
            mean = [sum(self.slots[r][i] for r in range(self.n)) / self.n for i in range(len(params))]
            self.barrier.wait()                       # everybody has read all slots before anyone overwrites its own
            for p, g in zip(params, mean):
                p.grad.copy_(g)
        return fn


def shared_weight_runs(cases, model, denoiser, iter_max, sigma, **kw):
    """cases: list of (meas, mask, warm, orig).  Returns the list of the per-group result tuples of
    ``twoStageAdmm_denoise_bayer`` (each with ITS model copy, all copies equal at the end)."""
    from . import admm
    n = len(cases)
    sync = _MeanGrad(n)
    out, err = [None] * n, [None] * n
    torch.set_num_threads(max(1, torch.get_num_threads() // n))

    def run(r):
        try:
            meas, mask, warm, orig = cases[r]
            out[r] = admm.twoStageAdmm_denoise_bayer(meas, mask, 1, 0.01, denoiser, iter_max, False, sigma,
                                                     x0_bayer=torch.from_numpy(warm), X_orig=orig,
                                                     model_denoise=copy.deepcopy(model), grad_hook=sync.hook(r),
                                                     rng=np.random.RandomState(42), **kw)
        except BaseException as e:                    # noqa: BLE001 - re-raised on the caller's thread
            err[r] = e
            sync.barrier.abort()
    ts = [threading.Thread(target=run, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out
