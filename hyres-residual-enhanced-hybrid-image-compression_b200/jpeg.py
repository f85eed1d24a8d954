"""JPEG stage at the boundary of the hot path (models/utils/turbo_jpeg_compression.py:17-77).

The JPEG round trip itself is third-party CPU code on both sides (libturbojpeg); only the
residual subtraction / add-back it feeds is in the kernel scope.
PyTurboJPEG is used when importable; otherwise OpenCV's libjpeg-turbo build stands in with
the parameters PyTurboJPEG's defaults imply (RGB array handed over as BGR, 4:2:2).
"""
import io

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

    def compress(self, x):
        x_cpu = x.cpu() if x.device.type != "cpu" else x
        bufs = []
        for i in range(x_cpu.size(0)):
            img = torch.clamp(x_cpu[i], 0, 1)
            if img.size(0) == 1:
                img = img.repeat(3, 1, 1)
            img_np = (img.permute(1, 2, 0) * 255).byte().numpy()  # .byte() truncates (reference behaviour)
            if self.backend == "turbojpeg":
                data = self._tj.encode(img_np, quality=self.quality)
            else:
                import cv2
                ok, enc = cv2.imencode(".jpg", img_np, [cv2.IMWRITE_JPEG_QUALITY, int(self.quality),
                                                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
                if not ok:
                    raise RuntimeError("JPEG encode failed")
                data = enc.tobytes()
            bufs.append(io.BytesIO(data))
        return bufs

    def decompress(self, compressed_buffers, device):
        imgs = []
        for buf in compressed_buffers:
            raw = buf.getvalue()
            if self.backend == "turbojpeg":
                dec = self._tj.decode(raw)
            else:
                import cv2
                dec = cv2.imdecode(np.frombuffer(raw, dtype=np.uint8), cv2.IMREAD_COLOR)
            imgs.append(torch.from_numpy(dec).float().permute(2, 0, 1) / 255.0)
        return torch.stack(imgs, dim=0).to(device)

    def forward(self, x):
        device = x.device
        bufs = self.compress(x)
        N, _, H, W = x.size()
        bits = sum(len(b.getvalue()) * 8 for b in bufs)
        return self.decompress(bufs, device), bits / (N * H * W)
