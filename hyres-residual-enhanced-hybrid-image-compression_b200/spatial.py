"""Bit-exact spatial sharding of ``compress`` (SURVEY.md section 8e tier T-B, section 8f rank 3).

``dist.tile_grid`` (tier T-A) treats the tiles of a large image as independent images: each gets its own strings, and
the result differs from coding the whole image because every tile border is a zero-padded image border.  Here one
image is coded by several ranks and the strings are byte-identical to the single-GPU whole-image ``compress``:

  * every rank runs the analysis / hyper / context / parameter networks on its tiles *extended by a halo* that is
    wider than the receptive field of everything that decides an integer (g_a 50 px, z 106 px, the entropy
    parameters through h_s ~162 px, the second pass through the 5x5 context another 16 px: 178 px; the default halo
    is 256 px, a multiple of the 32-pixel hyper-latent grid).  Windows are clipped at the true image border, where
    the zero padding is the whole-image run's own;
  * inside the halo-free interior every convolution output is the same arithmetic as in the whole-image run: the
    fp32-equivalent trunk runs on the streaming kernel, whose per-element accumulation order does not depend on
    the tensor's extent, the tile position, the batch or the CTA count (tests/test_gpu_precise.py);
  * each rank writes the integers of its interiors into zero-initialised whole-image tensors in the coder's
    (B, C, h, w) order; one integer all-reduce (each element is written by exactly one rank) assembles them, and
    rank 0 runs the host entropy coder once over the whole image.

The JPEG stage is computed for the whole image on every rank (0.15 ms; its chroma up-sampling crosses block borders,
so windows would not reproduce it).
"""
import time

import torch
import torch.distributed as dist

from .dist import shard_range, tile_grid
from .models import _check_finite

DEFAULT_HALO = 256


def spatial_windows(H, W, rows, cols, halo=DEFAULT_HALO):
    """[(tile, window)] with tile = (h0, h1, w0, w1) and window = the tile grown by ``halo`` and clipped to the
    image; everything is a multiple of 32."""
    if halo % 32:
        raise ValueError("halo must be a multiple of 32 (the hyper-latent grid)")
    out = []
    for h0, h1, w0, w1 in tile_grid(H, W, rows, cols):
        out.append(((h0, h1, w0, w1), (max(0, h0 - halo), min(H, h1 + halo), max(0, w0 - halo), min(W, w1 + halo))))
    return out


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def encode_symbols_sharded(codec, residual, rows, cols, halo=DEFAULT_HALO, group=None, ranks=None):
    """The five integer streams of ``LightWeightCheckerboard.encode_symbols`` for the whole ``residual``
    ([B,3,H,W] fp32 on the device), computed tile by tile with halos.  Each rank of ``group`` takes a contiguous
    share of the tiles; ``ranks=(r, n)`` overrides the rank / world size (single-process emulation of n ranks:
    call once per r and add the results).  Returns whole-image int32 tensors (complete on every rank after the
    all-reduce)."""
    B, _, H, W = residual.shape
    M, N = codec.M, codec.N
    dev = residual.device
    out = {"sym_z": torch.zeros((B, N, H // 32, W // 32), dtype=torch.int32, device=dev)}
    for k in ("sym_a", "idx_a", "sym_na", "idx_na"):
        out[k] = torch.zeros((B, M, H // 8, W // 8), dtype=torch.int32, device=dev)
    rank, world = ranks if ranks is not None else _world(group)
    wins = spatial_windows(H, W, rows, cols, halo)
    lo, hi = shard_range(len(wins), rank, world)
    for (h0, h1, w0, w1), (H0, H1, W0, W1) in wins[lo:hi]:
        s = codec.encode_symbols(residual[:, :, H0:H1, W0:W1].contiguous())
        _check_finite(s)
        for k, d in (("sym_z", 32), ("sym_a", 8), ("idx_a", 8), ("sym_na", 8), ("idx_na", 8)):
            out[k][:, :, h0 // d:h1 // d, w0 // d:w1 // d] = s[k][:, :, (h0 - H0) // d:(h1 - H0) // d,
                                                                (w0 - W0) // d:(w1 - W0) // d]
    if ranks is None and world > 1:
        flat = torch.cat([t.reshape(-1) for t in out.values()])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)  # every element has exactly one writer
        off = 0
        for k, t in out.items():
            out[k] = flat[off:off + t.numel()].view(t.shape)
            off += t.numel()
    return out


def compress_sharded(model, x, rows=2, cols=4, halo=DEFAULT_HALO, group=None):
    """``ResidualJPEGCompression.compress`` (models/hyres.py:79-102) / ``LightWeightCheckerboard.compress`` of the
    image batch ``x`` with the residual codec's GPU work split over the ranks of ``group``.  Every rank passes the
    same ``x``; rank 0 returns the dict of ``compress`` (byte-identical strings), the others return None."""
    start = time.time()
    wrapper = hasattr(model, "residual_model")
    codec = model.residual_model if wrapper else model
    dev = next(model.parameters()).device
    xd = x.to(dev, torch.float32).contiguous()
    codec._check_input(xd)
    rank, world = _world(group)
    jpeg_buffers = None
    if wrapper:
        if rank == 0:
            jpeg_buffers, jpeg_decoded = model.jpeg.compress_device(xd)
        else:
            jpeg_decoded, _ = model.jpeg.forward_device(xd)
        residual = xd - jpeg_decoded  # fp32, the subtraction compress() fuses into the first layer
    else:
        residual = xd
    s = encode_symbols_sharded(codec, residual, rows, cols, halo, group)
    if rank != 0:
        return None
    gc, ebm = codec.gaussian_conditional, codec.entropy_bottleneck
    z_strings = ebm.encode_symbols(s["sym_z"], ebm._build_indexes(s["sym_z"].size()))
    anchor, non_anchor = gc.encode_symbol_groups([(s["sym_a"], s["idx_a"]), (s["sym_na"], s["idx_na"])])
    out = {"strings": [[anchor, non_anchor], z_strings], "shape": torch.Size(s["sym_z"].shape[-2:]),
           "time": time.time() - start}
    if wrapper:
        out["jpeg_buffers"] = jpeg_buffers
    return out
