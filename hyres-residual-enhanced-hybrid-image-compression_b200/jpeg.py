"""JPEG stage at the boundary of the hot path (models/utils/turbo_jpeg_compression.py:17-77).

The JPEG round trip itself is third-party CPU code on both sides (libturbojpeg); only the
residual subtraction / add-back it feeds is in the kernel scope.
PyTurboJPEG is used when importable; otherwise OpenCV's libjpeg-turbo build stands in with
the parameters PyTurboJPEG's defaults imply (RGB array handed over as BGR, 4:2:2).
"""
import io
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from torch import nn


def _load_backend():
    try:
        from turbojpeg import TurboJPEG  # noqa: F401
        return "turbojpeg"
    except Exception:
        pass
    try:
        import cv2  # noqa: F401
        return "cv2"
    except Exception as e:  # pragma: no cover
        raise OSError("neither PyTurboJPEG nor OpenCV is available for the JPEG stage") from e


_POOL = None


def _pool():
    """libjpeg-turbo releases the GIL (through cv2 / ctypes): the per-image loop of the reference
    (models/utils/turbo_jpeg_compression.py:24-37,44-57) runs on a small thread pool instead."""
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    return _POOL


class TurboJPEGCompression(nn.Module):
    def __init__(self, quality=25, lib_path=None):
        super().__init__()
        self.quality = quality
        self.backend = _load_backend()
        self._tj = None
        if self.backend == "turbojpeg":
            from turbojpeg import TurboJPEG
            try:
                self._tj = TurboJPEG(lib_path=lib_path) if lib_path else TurboJPEG()
            except Exception:
                self.backend = "cv2"
                import cv2  # noqa: F401

    def _encode_one(self, img_np):
        if self.backend == "turbojpeg":
            return self._tj.encode(img_np, quality=self.quality)
        import cv2
        ok, enc = cv2.imencode(".jpg", img_np, [cv2.IMWRITE_JPEG_QUALITY, int(self.quality),
                                               cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                               cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        if not ok:
            raise RuntimeError("JPEG encode failed")
        return enc.tobytes()

    def _decode_one(self, raw):
        if self.backend == "turbojpeg":
            return self._tj.decode(raw)
        import cv2
        return cv2.imdecode(np.frombuffer(raw, dtype=np.uint8), cv2.IMREAD_COLOR)

    def compress(self, x):
        x_cpu = x.cpu() if x.device.type != "cpu" else x
        x_cpu = torch.clamp(x_cpu, 0, 1)
        if x_cpu.size(1) == 1:
            x_cpu = x_cpu.repeat(1, 3, 1, 1)
        u8 = (x_cpu.permute(0, 2, 3, 1) * 255).byte().contiguous().numpy()  # .byte() truncates (reference behaviour)
        if u8.shape[0] == 1:
            datas = [self._encode_one(u8[0])]
        else:
            datas = list(_pool().map(self._encode_one, [u8[i] for i in range(u8.shape[0])]))
        return [io.BytesIO(d) for d in datas]

    def decompress(self, compressed_buffers, device):
        raws = [buf.getvalue() for buf in compressed_buffers]
        decs = [self._decode_one(raws[0])] if len(raws) == 1 else list(_pool().map(self._decode_one, raws))
        u8 = torch.from_numpy(np.stack(decs, axis=0))  # [N, H, W, 3] uint8
        dev = torch.device(device)
        if dev.type == "cuda":
            # 1 byte per sample over PCIe from pinned memory; the float conversion (u8 / 255.0, the same fp32
            # division as on the host) and the NHWC -> NCHW permute run on the device
            u8 = u8.pin_memory().to(dev, non_blocking=True)
        return (u8.permute(0, 3, 1, 2).float() / 255.0).contiguous()

    def forward(self, x):
        device = x.device
        bufs = self.compress(x)
        N, _, H, W = x.size()
        bits = sum(len(b.getvalue()) * 8 for b in bufs)
        return self.decompress(bufs, device), bits / (N * H * W)
