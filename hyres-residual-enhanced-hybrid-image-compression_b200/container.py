"""Byte container for the result of ``compress()``.

The reference keeps ``{"strings", "shape", "jpeg_buffers"}`` as a Python dict and only counts bytes
(src/inference.py:103-114); there is no file format.  This module gives the dict a self-describing
little-endian layout so that a compressed batch can leave the process and come back:

    magic "HYRS" | u16 version | u16 flags | u32 B | u32 shape_h | u32 shape_w
    then, per image b in 0..B-1:   u32 len | bytes     for each stream, in the order
        jpeg (only if flags & 1), anchor, non_anchor, z

Flag bits 1-2 record the arithmetic of the entropy-critical trunk that produced the strings, because the decoder
recomputes the CDF indexes and must land on the same integers: 0 = not recorded, 1 = "bf16" (decodable only by this
library's bf16 trunk), 2 = fp32-equivalent ("fp32x2" / "fp32x3" / "fp32h2": decodable by the fp32 reference decoder as well, up
to numerical ties).  ``pack(c, trunk=model.codec_precision)`` sets them, ``unpack`` returns them as ``"trunk"`` and
``decompress`` refuses a stream whose tag does not match the model's trunk.

``pack`` / ``unpack`` are exact inverses: ``decompress(unpack(pack(c)))`` equals ``decompress(c)`` bit for bit.
"""
import io
import struct

import torch

MAGIC = b"HYRS"
VERSION = 1
_FLAG_JPEG = 1
_TRUNK_SHIFT, _TRUNK_MASK = 1, 3
_TRUNK_CODE = {None: 0, "bf16": 1, "fp32x2": 2, "fp32x3": 2, "fp32h2": 2, "fp32": 2}
_TRUNK_NAME = {0: None, 1: "bf16", 2: "fp32"}


def pack(compressed, trunk=None):
    """compressed: the dict returned by ``ResidualJPEGCompression.compress`` (with ``jpeg_buffers``) or by
    ``LightWeightCheckerboard.compress`` (without) -> bytes.  ``trunk``: the model's ``codec_precision`` (recorded
    in the header; defaults to the dict's own ``"trunk"`` entry if it has one)."""
    if trunk is None:
        trunk = compressed.get("trunk")
    if trunk not in _TRUNK_CODE:
        raise ValueError(f"unknown trunk precision {trunk!r}")
    (anchor, non_anchor), z = compressed["strings"]
    B = len(z)
    if len(anchor) != B or len(non_anchor) != B:
        raise ValueError("anchor / non_anchor / z string lists must have one entry per image")
    jpeg = compressed.get("jpeg_buffers")
    if jpeg is not None and len(jpeg) != B:
        raise ValueError("jpeg_buffers must have one entry per image")
    h, w = (int(v) for v in compressed["shape"])
    flags = (_FLAG_JPEG if jpeg is not None else 0) | (_TRUNK_CODE[trunk] << _TRUNK_SHIFT)
    out = [MAGIC, struct.pack("<HHIII", VERSION, flags, B, h, w)]
    for b in range(B):
        streams = ([jpeg[b].getvalue()] if jpeg is not None else []) + [anchor[b], non_anchor[b], z[b]]
        for s in streams:
            s = bytes(s)
            out.append(struct.pack("<I", len(s)))
            out.append(s)
    return b"".join(out)


def unpack(data):
    """bytes -> dict accepted by ``decompress`` (``strings``, ``shape`` and, if present, ``jpeg_buffers``)."""
    data = bytes(data)
    if len(data) < 20 or data[:4] != MAGIC:
        raise ValueError("not a HYRS container")
    version, flags, B, h, w = struct.unpack_from("<HHIII", data, 4)
    if version != VERSION:
        raise ValueError(f"unsupported HYRS container version {version}")
    pos = 20
    has_jpeg = bool(flags & _FLAG_JPEG)
    jpeg, anchor, non_anchor, z = [], [], [], []

    def take():
        nonlocal pos
        if pos + 4 > len(data):
            raise ValueError("truncated HYRS container")
        (n,) = struct.unpack_from("<I", data, pos)
        pos += 4
        if pos + n > len(data):
            raise ValueError("truncated HYRS container")
        s = data[pos:pos + n]
        pos += n
        return s

    for _ in range(B):
        if has_jpeg:
            jpeg.append(io.BytesIO(take()))
        anchor.append(take())
        non_anchor.append(take())
        z.append(take())
    if pos != len(data):
        raise ValueError("trailing bytes after the last stream of a HYRS container")
    out = {"strings": [[anchor, non_anchor], z], "shape": torch.Size([h, w])}
    if has_jpeg:
        out["jpeg_buffers"] = jpeg
    trunk = _TRUNK_NAME.get((flags >> _TRUNK_SHIFT) & _TRUNK_MASK)
    if trunk is not None:
        out["trunk"] = trunk
    return out


def check_trunk(compressed, codec_precision):
    """Raise if a tagged stream was produced by a trunk whose integers this model's trunk cannot reproduce."""
    tag = compressed.get("trunk") if isinstance(compressed, dict) else None
    if tag is None:
        return
    mine = "bf16" if codec_precision == "bf16" else "fp32"
    if _TRUNK_NAME[_TRUNK_CODE[tag]] != mine:
        raise ValueError(f"the stream was coded with the {tag} trunk; this model decodes with {codec_precision} "
                         "(set model.codec_precision accordingly: the decoder must recompute the coder's CDF indexes)")
