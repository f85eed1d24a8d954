"""Multi-GPU plumbing: one process per GPU, the batch (or the tile list of a large image) is
sharded by image across ranks; the data path has no collective.  The only exchange is the
all-reduce of the four rate/distortion sums (replaces nn.DataParallel's gather of outputs,
src/utils/dataset_utils.py:76-82 + src/losses/rd_loss.py:23-26,39).  Works on NCCL (GPU) and
gloo (CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* when launched by torchrun.
    Returns (rank, world_size, local_rank); a plain `python` launch gives (0, 1, 0)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_grid(H, W, rows, cols, multiple=32):
    """Split an HxW image into rows x cols tiles whose sides are multiples of `multiple`
    (the codec needs multiples of 32; cfg3: 1408x2048 -> 2x4 tiles of 704x512).
    Returns [(h0, h1, w0, w1)] in row-major order."""
    if H % multiple or W % multiple:
        raise ValueError("image sides must be multiples of %d" % multiple)

    def cuts(n, k):
        units = n // multiple
        if k > units:
            raise ValueError("more tiles than %d-pixel units" % multiple)
        out, acc = [0], 0
        for i in range(k):
            acc += units // k + (1 if i < units % k else 0)
            out.append(acc * multiple)
        return out

    hc, wc = cuts(H, rows), cuts(W, cols)
    return [(hc[i], hc[i + 1], wc[j], wc[j + 1]) for i in range(rows) for j in range(cols)]


def reduce_stats(stats):
    """Sum a small tensor of statistics (sum log2 lik_y, sum log2 lik_z, sum squared error, pixel
    count) over all ranks, in place.  No-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rd_from_stats(stats, lmbda, jpeg_bpp=0.0, channels=3):
    """RateDistortionLoss (src/losses/rd_loss.py:18-44, alpha = 0) from globally reduced sums:
    stats = [sum log2 lik_y, sum log2 lik_z, sum (x_hat - x)^2, pixel count]."""
    s = stats.double()
    npx = s[3]
    out = {"y_bpp_loss": -s[0] / npx, "z_bpp_loss": -s[1] / npx}
    out["residual_bpp_loss"] = out["y_bpp_loss"] + out["z_bpp_loss"]
    out["bpp_loss"] = out["residual_bpp_loss"] + jpeg_bpp
    out["mse_loss"] = s[2] / (npx * channels) * 255 ** 2
    out["loss"] = lmbda * out["mse_loss"] + out["bpp_loss"]
    return out
