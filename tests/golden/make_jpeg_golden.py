"""Makes tests/golden/jpeg_golden.npz: JPEG files and decoded pixels produced by libjpeg-turbo itself (OpenCV's bundled
build, the library PyTurboJPEG wraps) with the parameters the reference's JPEG stage implies
(models/utils/turbo_jpeg_compression.py:32-35,52: RGB array read as BGR, 4:2:2, baseline Huffman).

    python tests/golden/make_jpeg_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(64, 96, 1, "smooth"), (64, 96, 25, "smooth"), (32, 32, 75, "noise"), (96, 64, 50, "smooth"),
         (64, 64, 100, "noise"), (128, 160, 1, "noise"), (32, 48, 1, "flat")]


def image(rng, H, W, kind):
    if kind == "noise":
        return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if kind == "flat":
        return np.full((H, W, 3), 255, dtype=np.uint8)
    base = rng.random((H // 8, W // 8, 3)).astype(np.float32)
    img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC) + 0.02 * rng.standard_normal((H, W, 3)).astype(np.float32)
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(1926)
    out = {"n": np.int64(len(CASES)), "libjpeg": np.bytes_(
        [l.strip() for l in cv2.getBuildInformation().splitlines() if "JPEG:" in l][0].encode())}
    for i, (H, W, q, kind) in enumerate(CASES):
        img = image(rng, H, W, kind)
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                            cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        assert ok
        out[f"img{i}"] = img
        out[f"q{i}"] = np.int64(q)
        out[f"file{i}"] = np.frombuffer(enc.tobytes(), dtype=np.uint8)
        out[f"dec{i}"] = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), **out)
    print("wrote jpeg_golden.npz", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
