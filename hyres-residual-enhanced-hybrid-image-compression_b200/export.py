"""Deployment export: the analogue of ``src/updata.py:36-78`` for the B200 path.

The reference's export step loads a training checkpoint, runs ``update(force=True)`` so that the CDF tables are
part of the state dict, and saves it.  ``export_model`` does the same and additionally stores every convolution's
*packed device operands* (K-major bf16 weight tiles -- three bf16 parts for the split-precision trunk --, the
tap-major copy of the 3-output-channel layers, the padded bias), keyed by a digest of the fp32 weights they were
packed from.  ``load_exported`` rebuilds the model and fills the layers from those operands instead of re-packing
(the state dict stays the authority: a layer whose weights no longer match its digest is packed from them).
"""
import torch

from . import ops
from .models import ResidualJPEGCompression

FORMAT = 1


def _build_engines(net):
    codec = net.residual_model
    codec._engine = None
    codec._precise = {}
    net._refine_engine = None
    codec.engine()
    if codec.codec_precision != "bf16":
        codec.precise(codec.codec_precision)
    if codec.precision not in ("bf16", codec.codec_precision):
        codec.precise(codec.precision)
    net.refine_engine()


def export_model(net, path=None, update=True):
    """net: ``ResidualJPEGCompression`` on a CUDA sm_100 device.  Returns the export dict (and saves it to ``path``):
    ``{"format", "state_dict", "packed", "codec_precision", "precision", "jpeg_quality"}``."""
    if next(net.parameters()).device.type != "cuda":
        raise RuntimeError("export_model packs on the device: move the model to a CUDA sm_100 device first")
    if update:
        net.update(force=True)
    ops.PACKED_COLLECT = {}
    try:
        _build_engines(net)
        packed = ops.PACKED_COLLECT
    finally:
        ops.PACKED_COLLECT = None
    blob = {"format": FORMAT, "state_dict": {k: v.detach().cpu() for k, v in net.state_dict().items()},
            "packed": packed, "codec_precision": net.residual_model.codec_precision,
            "precision": net.residual_model.precision, "jpeg_quality": net.jpeg.quality}
    if path is not None:
        torch.save(blob, path)
    return blob


def load_exported(blob_or_path, device="cuda"):
    """-> ``ResidualJPEGCompression`` in eval mode on ``device`` with its layers filled from the exported operands."""
    blob = torch.load(blob_or_path, weights_only=False) if isinstance(blob_or_path, (str, bytes)) or hasattr(
        blob_or_path, "read") else blob_or_path
    if blob.get("format") != FORMAT:
        raise ValueError(f"unsupported export format {blob.get('format')!r}")
    net = ResidualJPEGCompression.from_state_dict(blob["state_dict"], jpeg_quality=blob["jpeg_quality"])
    net.residual_model.codec_precision = blob["codec_precision"]
    net.residual_model.precision = blob["precision"]
    net = net.to(device).eval()
    ops.PACKED_CACHE = blob["packed"]
    try:
        _build_engines(net)
    finally:
        ops.PACKED_CACHE = None
    return net
