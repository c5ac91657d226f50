"""Single-node multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink only where the path
really exchanges data.

* Measurement groups / videos are independent work units: ``my_units`` deals them round-robin to
  the ranks, ``gather_units`` collects the numpy results on rank 0 (object gather; no collective touches the
  per-iteration data path).
* Shared-weight online fine-tuning (BASELINE config 4): every rank fine-tunes the SAME denoiser weights on its
  own group; ``grad_sync`` all-reduces (mean) the flat gradient bucket — one NCCL call of <= 10 MB per Adam step —
  so all ranks take identical steps.  This is a semantic change against the reference (which fine-tunes
  sequentially, group after group, ``reuse_model=True``) and is only used when requested.

On CPU (``gloo``) the same code paths run in the unit tests (tests/test_parallel_cpu.py).
"""
import os

import torch


class Context:
    def __init__(self, rank, world, local_rank, backend):
        self.rank, self.world, self.local_rank, self.backend = rank, world, local_rank, backend

    def my_units(self, n):
        """Indices of the independent units (measurement groups) this rank owns."""
        return list(range(self.rank, n, self.world))

    def gather_units(self, results):
        """dict{unit: payload} from every rank -> merged dict on rank 0 (others get {})."""
        if self.world == 1:
            return results
        import torch.distributed as dist
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(results, out, dst=0)
        merged = {}
        if self.rank == 0:
            for d in out:
                merged.update(d)
        return merged

    def grad_sync(self, flat_grad):
        """Mean all-reduce of the flat gradient bucket (identity on one rank)."""
        if self.world > 1:
            import torch.distributed as dist
            if self.backend == "nccl":
                dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
                flat_grad.div_(self.world)
        return flat_grad

    def broadcast_(self, tensor, src=0):
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(tensor, src)
        return tensor

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def finalize(self):
        if self.world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


def init(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and initialises the process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local_rank)
            dist.init_process_group(backend, **kw)
    return Context(rank, world, local_rank, backend)
