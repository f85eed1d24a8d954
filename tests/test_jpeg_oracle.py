"""The JPEG oracle (oracle/jpeg_oracle.py) pinned against libjpeg-turbo itself: byte-identical files and
pixel-identical decodes, on committed fixtures (tests/golden/jpeg_golden.npz, made by make_jpeg_golden.py with OpenCV's
bundled libjpeg-turbo) and live against cv2 when it is importable."""
import os

import numpy as np
import pytest

from oracle import jpeg_oracle as J

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg_golden.npz"))
N = int(G["n"])


@pytest.mark.parametrize("i", range(N))
def test_encode_is_byte_identical_to_libjpeg_turbo(i):
    img, q = G[f"img{i}"], int(G[f"q{i}"])
    assert J.encode(img, q) == G[f"file{i}"].tobytes()


@pytest.mark.parametrize("i", range(N))
def test_round_trip_pixels_match_libjpeg_turbo(i):
    img, q = G[f"img{i}"], int(G[f"q{i}"])
    assert np.array_equal(J.roundtrip(img, q), G[f"dec{i}"])


def test_header_is_a_fixed_623_bytes():
    assert len(J.header(512, 768, 1)) == 623
    assert J.header(64, 96, 1) == G["file0"].tobytes()[:623]


def test_quality_one_tables_saturate_at_255():
    ql, qc = J.quant_tables(1)
    assert int(ql.min()) == 255 and int(qc.min()) == 255


def test_stage_forward_matches_the_product_boundary_arithmetic():
    rng = np.random.default_rng(3)
    x = rng.random((2, 3, 32, 48), dtype=np.float32)
    dec, bpp, sizes = J.stage_forward(x, 25)
    u8 = (x.transpose(0, 2, 3, 1) * np.float32(255)).astype(np.uint8)
    for b in range(2):
        assert len(J.encode(u8[b], 25)) == sizes[b]
        assert np.array_equal((dec[b] * 255).round().astype(np.uint8).transpose(1, 2, 0), J.roundtrip(u8[b], 25))
    assert bpp == pytest.approx(8.0 * sum(sizes) / (2 * 32 * 48))


def test_rejects_sizes_that_would_need_edge_padding():
    with pytest.raises(ValueError):
        J.coefficients(np.zeros((20, 32, 3), np.uint8), 1)


def test_live_against_cv2_on_random_cases():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for _ in range(6):
        H, W, q = 8 * int(rng.integers(1, 9)), 16 * int(rng.integers(1, 6)), int(rng.integers(1, 101))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if rng.random() < 0.5:
            img = (img // 32) * 32  # long zero runs / ZRL
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                            cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        assert ok and J.encode(img, q) == enc.tobytes(), (H, W, q)
        assert np.array_equal(J.roundtrip(img, q), cv2.imdecode(enc, cv2.IMREAD_COLOR)), (H, W, q)


@pytest.mark.parametrize("i", range(N))
def test_host_assembly_of_the_c_abi_writes_libjpeg_turbos_file(build_lib, i):
    """hyres_jpeg_assemble (host half of the device JPEG stage: markers, byte stuffing, padding, EOI) fed the oracle's
    raw scan bits reproduces the library's file; needs no GPU."""
    from hyres_b200 import ops
    img, q = G[f"img{i}"], int(G[f"q{i}"])
    coefs, _ = J.coefficients(img, q)
    words, nbits = J.scan_bits(coefs)
    mine = ops.jpeg_assemble(words.view(np.int32), nbits, img.shape[0], img.shape[1], q)
    assert mine == G[f"file{i}"].tobytes()
    assert mine == J.encode(img, q)


def test_host_assembly_rejects_bad_arguments(build_lib):
    from hyres_b200 import ops
    from hyres_b200._lib import HyresError
    with pytest.raises(HyresError):
        ops.jpeg_assemble(np.zeros(4, np.int32), 64, 70000, 32, 1)
