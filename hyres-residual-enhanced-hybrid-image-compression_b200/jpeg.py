"""JPEG stage of the hot path (models/utils/turbo_jpeg_compression.py:17-77).

The reference round-trips every image through libjpeg-turbo on the CPU, one image at a time.  For CUDA inputs the
encode side runs on the device instead (``csrc/jpeg.cu``: colour conversion, 4:2:2 down-sampling, ISLOW DCT,
quantisation, Huffman coding, and the decoder's reconstruction), bit-exact with libjpeg-turbo: ``forward`` never
leaves the GPU, and ``compress`` only copies the entropy-coded scans to the host to wrap them into JPEG files that
are byte-identical to the library's (``tests/test_gpu_jpeg.py``).  Decoding JPEG *files* (``decompress``) and CPU
inputs stay on libjpeg-turbo itself: PyTurboJPEG when importable, otherwise OpenCV's build of the same library
with the parameters PyTurboJPEG's defaults imply (RGB array handed over as BGR, 4:2:2).
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

    # ---- device stage (CUDA inputs) ----
    @staticmethod
    def _device_ok(x):
        return x.is_cuda and x.dim() == 4 and x.size(1) in (1, 3) and x.size(2) % 8 == 0 and x.size(3) % 16 == 0

    @staticmethod
    def _rgb(x):
        x = x.float()
        if x.size(1) == 1:
            x = x.repeat(1, 3, 1, 1)
        return x.contiguous()

    def forward_device(self, x):
        """CUDA ``[B,3,H,W]`` -> (decoded fp32 ``[B,3,H,W]``, bpp as a 0-d fp32 CUDA tensor); no host sync."""
        from . import ops
        r = ops.jpeg_forward(self._rgb(x), self.quality)
        return r["decoded"], r["bpp"]

    def compress_device(self, x):
        """CUDA ``[B,3,H,W]`` -> (list of ``io.BytesIO`` JPEG files, decoded fp32 ``[B,3,H,W]`` on the device):
        the files are what libjpeg-turbo writes for these images, the pixels what it decodes from them."""
        from . import ops
        xr = self._rgb(x)
        r = ops.jpeg_forward(xr, self.quality, want_sizes=False, want_scan=True)
        nbits = r["nbits"].cpu()
        nwords = int((int(nbits.max()) + 31) // 32)
        words = r["words"][:, :max(nwords, 1)].cpu().numpy()
        H, W = xr.shape[2:]
        return [io.BytesIO(ops.jpeg_assemble(words[i], int(nbits[i]), H, W, self.quality))
                for i in range(xr.size(0))], r["decoded"]

    def compress(self, x):
        if self._device_ok(x):
            return self.compress_device(x)[0]
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
            # 1 byte per sample over PCIe from pinned memory; the NHWC -> NCHW permute runs on the device and the
            # division (torch's CUDA division by a Python scalar multiplies by the reciprocal, which is 1 ulp off
            # for some bytes) goes through a 256-entry table computed with the host's IEEE division
            u8 = u8.pin_memory().to(dev, non_blocking=True)
            lut = (torch.arange(256, dtype=torch.float32) / 255.0).to(dev)
            return lut[u8.permute(0, 3, 1, 2).long()].contiguous()
        return (u8.permute(0, 3, 1, 2).float() / 255.0).contiguous()

    def forward(self, x):
        if self._device_ok(x):
            dec, bpp = self.forward_device(x)
            return dec, float(bpp)  # the reference returns a Python float (one host sync)
        device = x.device
        bufs = self.compress(x)
        N, _, H, W = x.size()
        bits = sum(len(b.getvalue()) * 8 for b in bufs)
        return self.decompress(bufs, device), bits / (N * H * W)
