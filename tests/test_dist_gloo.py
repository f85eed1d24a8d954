"""world_size-2 gloo test of the sharding + statistics reduction used by bench.py --gpus N."""
import math
import os
import socket

import pytest
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
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from hyres_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    B = 5
    g = torch.Generator().manual_seed(0)
    lik_y, lik_z = torch.rand(B, 7, generator=g) * 0.9 + 0.05, torch.rand(B, 3, generator=g) * 0.9 + 0.05
    se = torch.rand(B, generator=g)
    lo, hi = D.shard_range(B, r, w)
    stats = torch.tensor([lik_y[lo:hi].log2().sum(), lik_z[lo:hi].log2().sum(), se[lo:hi].sum(), float(hi - lo) * 10],
                         dtype=torch.float64)
    D.reduce_stats(stats)
    t = D.max_over_ranks(1.0 + r)
    q.put((r, stats.tolist(), t, (lo, hi)))
    torch.distributed.destroy_process_group()


def test_two_rank_stats_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    lik_y, lik_z = torch.rand(5, 7, generator=g) * 0.9 + 0.05, torch.rand(5, 3, generator=g) * 0.9 + 0.05
    se = torch.rand(5, generator=g)
    want = [lik_y.log2().sum().item(), lik_z.log2().sum().item(), se.sum().item(), 50.0]
    ranges = sorted(r[3] for r in res)
    assert ranges == [(0, 3), (3, 5)]
    for _, stats, t, _ in res:
        assert t == 2.0
        for a, b in zip(stats, want):
            assert math.isclose(a, b, rel_tol=1e-6)


def test_shard_range_and_tiles():
    from hyres_b200 import dist as D
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 4, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)
    tiles = D.tile_grid(1408, 2048, 2, 4)  # BASELINE.json configs[2]
    assert len(tiles) == 8 and all((h1 - h0, w1 - w0) == (704, 512) for h0, h1, w0, w1 in tiles)
    assert all(v % 32 == 0 for t in tiles for v in t)
    cover = torch.zeros(1408, 2048, dtype=torch.int32)
    for h0, h1, w0, w1 in tiles:
        cover[h0:h1, w0:w1] += 1
    assert (cover == 1).all()
    tiles = D.tile_grid(96, 160, 2, 2)
    assert [(t[1] - t[0], t[3] - t[2]) for t in tiles] == [(64, 96), (64, 64), (32, 96), (32, 64)]
    with pytest.raises(ValueError):
        D.tile_grid(100, 160, 2, 2)


def test_rd_from_stats_matches_loss(oracle):
    from hyres_b200 import dist as D
    g = torch.Generator().manual_seed(1)
    out = {"likelihoods": {"y": torch.rand(2, 4, 8, 8, generator=g) * 0.9 + 0.05,
                           "z": torch.rand(2, 4, 2, 2, generator=g) * 0.9 + 0.05},
           "x_hat": torch.rand(2, 3, 64, 64, generator=g), "jpeg_bpp_loss": torch.tensor(0.25)}
    x = torch.rand(2, 3, 64, 64, generator=g)
    want = oracle.RateDistortionLoss(lmbda=0.008)(out, x)
    stats = torch.tensor([out["likelihoods"]["y"].double().log2().sum(), out["likelihoods"]["z"].double().log2().sum(),
                          (out["x_hat"] - x).double().pow(2).sum(), 2 * 64 * 64.0])
    got = D.rd_from_stats(stats, 0.008, jpeg_bpp=0.25)
    for k in ("y_bpp_loss", "z_bpp_loss", "bpp_loss", "mse_loss", "loss"):
        assert math.isclose(float(got[k]), float(want[k]), rel_tol=1e-5), k


def _bucket_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hyres_b200.train import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(40, 300), torch.nn.Tanh(), torch.nn.Linear(300, 7))
    unused = torch.nn.Parameter(torch.zeros(5))  # receives no gradient: must not stall the reduction
    params = list(net.parameters()) + [unused]
    buckets = GradBuckets(params, bucket_bytes=4096)  # several small buckets
    out = []
    for step in range(2):
        for p in params:
            p.grad = None
        x = torch.full((3, 40), float(rank + 1 + step))
        net(x).square().sum().backward()
        local = [p.grad.numpy().copy() for p in net.parameters()]
        buckets.finish()
        out.append((local, [p.grad.numpy().copy() for p in net.parameters()]))
    assert unused.grad is None
    q.put((rank, out, len(buckets.buckets)))
    dist.destroy_process_group()


def test_gradient_buckets_average_over_two_ranks():
    """hyres_b200.train.GradBuckets (gradient all-reduce launched from autograd hooks; replaces nn.DataParallel,
    src/training.py:211-212): after finish() every rank holds the mean of the ranks' gradients, in several buckets,
    across consecutive steps, with a parameter that never receives a gradient."""
    import socket
    import torch
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        rank, out, nb = q.get(timeout=120)
        res[rank] = out
        assert nb >= 2
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        mean = [(a + b) / 2 for a, b in zip(res[0][step][0], res[1][step][0])]
        for r in range(2):
            for got, want in zip(res[r][step][1], mean):
                torch.testing.assert_close(torch.from_numpy(got), torch.from_numpy(want), rtol=1e-6, atol=1e-7)
