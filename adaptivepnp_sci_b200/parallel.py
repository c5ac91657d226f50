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

    def check_even_split(self, n, what="units"):
        """Lock-step collectives (the shared-weight gradient all-reduce) need every rank to own the same number of units:
        a rank without a unit would never enter the all-reduce and the others would hang in NCCL until the timeout."""
        if n % self.world != 0:
            raise ValueError("%d %s do not divide evenly over %d ranks: a shared-weight run needs equal shares "
                             "(every rank joins every gradient all-reduce)" % (n, what, self.world))

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


class TileContext:
    """Row-strip spatial tiling of ONE large frame over the ranks (BASELINE config 5).

    Rank r owns the contiguous rows [r*rows, (r+1)*rows) of every plane (rows % 4 == 0 keeps the Bayer phase, the
    stride-2 and the PixelShuffle phases aligned).  Pixel-wise kernels (projection, dual updates) need nothing from
    the neighbours; stencils do:

    * Malvar demosaic: 2 rows of the merged mosaic,
    * denoiser: ``halo`` rows of its RGB input (FastDVDnet two-stage cascade: receptive field -77..+73 rows -> 80;
      FFDNet-colour: -24..+25 -> 28), exchanged once per ADMM iteration; the strip is then denoised with its halo and
      cropped ("overlap-tile"), so zero padding is only ever applied at true image borders.

    ``exchange`` moves the halo rows with point-to-point sends between neighbouring ranks (NCCL over NVLink on GPUs,
    gloo in the CPU/one-GPU tests).  No collective is on the per-iteration path; the only reductions are the scalar
    PSNR sums, the fine-tune loss/gradient all-reduce (SUM: the loss is normalised by the pixel count of the whole frame)
    and the final gather of the result strips."""

    def __init__(self, ctx, H_total, W):
        if H_total % (4 * ctx.world) != 0:
            raise ValueError("tiled mode needs H divisible by 4*world_size (got H=%d, world=%d)" % (H_total, ctx.world))
        self.ctx = ctx
        self.rank, self.world = ctx.rank, ctx.world
        self.H_total, self.W = H_total, W
        self.rows = H_total // ctx.world
        self.r0 = self.rank * self.rows
        self.total_pixels = H_total * W

    def slice_rows(self, a):
        """numpy/torch [H_total, ...] -> this rank's rows."""
        return a[self.r0:self.r0 + self.rows]

    def halo_sizes(self, halo):
        if self.world > 1 and halo > self.rows:
            # a strip shorter than the denoiser's receptive field would need rows from rank +-2 as well; clamping the halo
            # would silently break the "tiled == un-tiled" guarantee
            raise ValueError("tiled mode: strips of %d rows are shorter than the %d-row halo the denoiser needs "
                             "(use fewer ranks for a %d-row frame)" % (self.rows, halo, self.H_total))
        return (halo if self.rank > 0 else 0), (halo if self.rank < self.world - 1 else 0)

    def exchange(self, own, halo):
        """own [B,C,rows,W] (contiguous, this rank's rows) -> (ext [B,C,top+rows+bot,W], top)."""
        import torch.distributed as dist
        top, bot = self.halo_sizes(halo)
        B, C, rows, W = own.shape
        ext = torch.empty((B, C, top + rows + bot, W), dtype=own.dtype, device=own.device)
        ext[:, :, top:top + rows].copy_(own)
        if self.world == 1:
            return ext, 0
        host = self.ctx.backend != "nccl"          # gloo moves CUDA halos through host memory (tests / CPU runs)
        ops, bufs = [], []
        if top:      # upper neighbour: send my first rows, receive its last rows
            snd = own[:, :, :top].contiguous()
            snd = snd.cpu() if host else snd
            rcv = torch.empty_like(snd)
            ops += [dist.P2POp(dist.isend, snd, self.rank - 1), dist.P2POp(dist.irecv, rcv, self.rank - 1)]
            bufs.append(("top", rcv, snd))
        if bot:
            snd = own[:, :, rows - bot:].contiguous()
            snd = snd.cpu() if host else snd
            rcv = torch.empty_like(snd)
            ops += [dist.P2POp(dist.isend, snd, self.rank + 1), dist.P2POp(dist.irecv, rcv, self.rank + 1)]
            bufs.append(("bot", rcv, snd))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for where, rcv, _ in bufs:
            if where == "top":
                ext[:, :, :top].copy_(rcv)
            else:
                ext[:, :, top + rows:].copy_(rcv)
        return ext, top

    def all_reduce_sum(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def gather_rows(self, strip):
        """strip [..., rows, W] on every rank -> full [..., H_total, W] on every rank."""
        if self.world == 1:
            return strip
        import torch.distributed as dist
        src = strip.contiguous()
        if self.ctx.backend != "nccl":
            src = src.cpu()
        parts = [torch.empty_like(src) for _ in range(self.world)]
        dist.all_gather(parts, src)
        return torch.cat(parts, dim=-2).to(strip.device)


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
