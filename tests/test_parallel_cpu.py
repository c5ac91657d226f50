"""CPU (gloo, world_size 2): the N>1 host logic — unit sharding, result gathering, shared-weight gradient sync."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from adaptivepnp_sci_b200 import parallel
    ctx = parallel.init(backend="gloo")
    units = ctx.my_units(5)
    # every rank "reconstructs" its own measurement groups
    results = {u: (np.full((2, 2), u, np.float32), np.array([10.0 + u])) for u in units}
    merged = ctx.gather_units(results)
    # shared-weight fine-tune: identical averaged gradient on every rank
    g = torch.full((7,), float(rank + 1))
    ctx.grad_sync(g)
    w = torch.zeros(3) + rank
    ctx.broadcast_(w, 0)
    ctx.barrier()
    q.put((rank, units, sorted(merged.keys()), g.tolist(), w.tolist()))
    ctx.finalize()


def test_two_rank_sharding_and_grad_sync():
    world, port = 2, _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, u0, m0, g0, w0), (r1, u1, m1, g1, w1) = out
    assert u0 == [0, 2, 4] and u1 == [1, 3]                    # round-robin, disjoint, complete
    assert m0 == [0, 1, 2, 3, 4] and m1 == []                  # gathered on rank 0 only
    assert g0 == g1 == [1.5] * 7                               # mean of the two ranks' gradients
    assert w0 == w1 == [0.0, 0.0, 0.0]


def test_single_rank_is_identity():
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    from adaptivepnp_sci_b200 import parallel
    ctx = parallel.init(backend="gloo")
    assert ctx.world == 1 and ctx.my_units(3) == [0, 1, 2]
    g = torch.ones(4)
    assert ctx.grad_sync(g) is g and ctx.gather_units({1: 2}) == {1: 2}


def _uneven_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from adaptivepnp_sci_b200 import parallel
    ctx = parallel.init(backend="gloo")
    try:
        ctx.check_even_split(3, "measurement groups")       # 3 groups over 2 ranks: rank 1 would skip an all-reduce
        q.put((rank, "accepted"))
    except ValueError as e:
        q.put((rank, str(e)))
    ctx.check_even_split(4, "measurement groups")           # 4 over 2 is fine
    ctx.finalize()


def test_shared_weights_need_an_even_split():
    """--share-weights with groups that do not divide over the ranks must fail loudly, not hang in NCCL."""
    world, port = 2, _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_uneven_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all("do not divide evenly" in msg for _, msg in out)


def test_halo_longer_than_strip_is_refused():
    """a strip shorter than the denoiser's halo must raise instead of silently clamping the halo."""
    import pytest
    from adaptivepnp_sci_b200 import parallel
    ctx = parallel.Context(rank=1, world=8, local_rank=1, backend="gloo")
    tile = parallel.TileContext(ctx, 512, 64)                # 64-row strips
    assert tile.halo_sizes(28) == (28, 28)
    with pytest.raises(ValueError, match="shorter than the 80-row halo"):
        tile.halo_sizes(80)
    one = parallel.TileContext(parallel.Context(0, 1, 0, "gloo"), 64, 64)
    assert one.halo_sizes(80) == (0, 0)


def _gather_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from adaptivepnp_sci_b200 import parallel
    ctx = parallel.init(backend="gloo")
    tile = parallel.TileContext(ctx, 16, 6)                  # 8-row strips
    full = torch.arange(2 * 16 * 6, dtype=torch.float32).view(2, 16, 6)
    strip = tile.slice_rows(full.permute(1, 2, 0).numpy())    # the solvers slice [H, W, ...] host arrays
    mine = torch.from_numpy(strip).permute(2, 0, 1).contiguous()
    everywhere = tile.gather_rows(mine)
    root = tile.gather_rows(mine, root_only=True)
    q.put((rank, bool(torch.equal(everywhere, full)), None if root is None else bool(torch.equal(root, full))))
    ctx.finalize()


def test_strip_gather_all_and_root_only():
    """TileContext.gather_rows: every rank gets the full frame, or (root_only) rank 0 alone and the others None."""
    world, port = 2, _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, True, True), (1, True, None)]
