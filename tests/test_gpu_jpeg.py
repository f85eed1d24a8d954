"""The device JPEG stage (csrc/jpeg.cu) through the C-ABI against libjpeg-turbo itself.

Bar: bit-exact.  Decoded pixels, file sizes and the complete JPEG files are compared with (a) committed fixtures
written by libjpeg-turbo 3.1.2 (tests/golden/jpeg_golden.npz), (b) the numpy oracle (oracle/jpeg_oracle.py, itself
pinned to the library), and (c) cv2's libjpeg-turbo live at BASELINE.json's full batch size when cv2 is importable."""
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg_golden.npz"))
N = int(G["n"])


def _as_float(img_u8):
    """u8 [H,W,3] -> fp32 [1,3,H,W] whose reference-style ``(x * 255).byte()`` truncation gives the bytes back."""
    return ((torch.from_numpy(img_u8.astype(np.float32)) + 0.5) / 255.0).permute(2, 0, 1)[None].contiguous()


@pytest.fixture(scope="module")
def stage(build_lib):
    import hyres_b200
    return hyres_b200


@pytest.mark.parametrize("i", range(N))
def test_fixture_pixels_sizes_and_files_are_libjpeg_turbos(stage, i):
    from hyres_b200 import ops
    img, q = G[f"img{i}"], int(G[f"q{i}"])
    ref_file, ref_dec = G[f"file{i}"].tobytes(), G[f"dec{i}"]
    x = _as_float(img).cuda()
    r = ops.jpeg_forward(x, q, want_scan=True)
    dec = r["decoded"].cpu()
    want = (torch.from_numpy(ref_dec).permute(2, 0, 1).float() / 255.0)[None]
    assert torch.equal(dec, want), "decoded pixels differ from libjpeg-turbo's"
    assert int(r["sizes"][0]) == len(ref_file), "file size differs from libjpeg-turbo's"
    nbits = int(r["nbits"][0])
    mine = ops.jpeg_assemble(r["words"][0].cpu().numpy(), nbits, img.shape[0], img.shape[1], q)
    assert mine == ref_file, "JPEG file is not byte-identical to libjpeg-turbo's"


def test_stage_module_api_matches_the_reference_stage(stage):
    """TurboJPEGCompression.forward / compress on CUDA tensors (models/utils/turbo_jpeg_compression.py:17-77)."""
    from oracle import jpeg_oracle as J
    tj = stage.TurboJPEGCompression(quality=1)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(3, 3, 64, 96, generator=g)
    x[0] = -0.2 + 1.4 * x[0]  # exercises the clamp
    dec, bpp = tj(x.cuda())
    odec, obpp, osizes = J.stage_forward(x.numpy(), 1)
    assert isinstance(bpp, float) and bpp == pytest.approx(obpp, rel=1e-6)
    assert torch.equal(dec.cpu(), torch.from_numpy(odec))
    bufs = tj.compress(x.cuda())
    assert all(isinstance(b, io.BytesIO) for b in bufs) and [len(b.getvalue()) for b in bufs] == osizes
    u8 = (np.clip(x.numpy(), 0, 1).transpose(0, 2, 3, 1) * np.float32(255)).astype(np.uint8)
    for b in range(3):
        assert bufs[b].getvalue() == J.encode(u8[b], 1)
    # the CPU decoder of the stage (libjpeg-turbo) returns the pixels the device reconstruction predicted
    assert torch.equal(tj.decompress(bufs, "cuda").cpu(), dec.cpu())
    # grayscale input is repeated to three channels like the reference does
    d1, _ = tj(x[:, :1].cuda())
    d3, _ = tj(x[:, :1].repeat(1, 3, 1, 1).cuda())
    assert torch.equal(d1, d3)


@pytest.mark.parametrize("q", [1, 25, 90])
def test_full_batch_against_libjpeg_turbo_live(stage, q):
    """configs[1] size (16 x 768 x 512): every image's file size and decoded pixels equal cv2's libjpeg-turbo."""
    cv2 = pytest.importorskip("cv2")
    from hyres_b200 import ops, synthetic
    x = synthetic.synthetic_image(16, 512, 768, seed=1926)
    r = ops.jpeg_forward(x.cuda(), q, want_scan=True)
    dec = r["decoded"].cpu()
    sizes = r["sizes"].cpu().tolist()
    u8 = (x.clamp(0, 1).permute(0, 2, 3, 1) * 255).byte().numpy()
    for b in range(16):
        ok, enc = cv2.imencode(".jpg", u8[b], [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                              cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        assert ok and sizes[b] == len(enc), (b, sizes[b], len(enc))
        ref = torch.from_numpy(cv2.imdecode(enc, cv2.IMREAD_COLOR)).permute(2, 0, 1).float() / 255.0
        assert torch.equal(dec[b], ref), b
        if b in (0, 15):
            mine = ops.jpeg_assemble(r["words"][b].cpu().numpy(), int(r["nbits"][b]), 512, 768, q)
            assert mine == enc.tobytes()


def test_model_forward_runs_the_stage_on_the_device(stage, oracle_net):
    """ResidualJPEGCompression.forward without an injected JPEG result: jpeg_decoded / jpeg_bpp_loss / residual are
    what the reference's CPU stage produces (oracle on the same input), bit for bit."""
    from oracle import jpeg_oracle as J
    import hyres_b200
    pnet = hyres_b200.ResidualJPEGCompression()
    pnet.load_state_dict(oracle_net.state_dict())
    pnet = pnet.cuda().eval()
    from hyres_b200 import synthetic
    x = synthetic.synthetic_image(2, 64, 96, seed=9)
    odec, obpp, _ = J.stage_forward(x.numpy(), 1)
    with torch.no_grad():
        out = pnet(x.cuda())
        inj = pnet(x.cuda(), jpeg=(torch.from_numpy(odec), obpp))
    assert torch.equal(out["jpeg_decoded"].cpu(), torch.from_numpy(odec))
    assert float(out["jpeg_bpp_loss"]) == pytest.approx(obpp, rel=1e-6)
    assert torch.equal(out["residual"].cpu(), x - torch.from_numpy(odec))
    assert torch.equal(out["x_hat"], inj["x_hat"])
    # compress: device-made jpeg_buffers are the library's files; decompress reads them back
    with torch.no_grad():
        c = pnet.compress(x.cuda())
        d = pnet.decompress(c)
    u8 = (x.clamp(0, 1).permute(0, 2, 3, 1) * 255).byte().numpy()
    assert [b.getvalue() for b in c["jpeg_buffers"]] == [J.encode(u8[i], 1) for i in range(2)]
    assert d["x_hat"].shape == x.shape


def test_unsupported_sizes_fail_loudly(stage):
    from hyres_b200 import ops
    with pytest.raises(ValueError):
        ops.jpeg_forward(torch.zeros(1, 3, 20, 32, device="cuda"), 1)
    with pytest.raises(ValueError):
        ops.jpeg_forward(torch.zeros(1, 3, 32, 40, device="cuda"), 1)


def test_extreme_inputs(stage):
    """All-white, all-black and a 0/1 checker at quality 100 (largest coefficients, 0xFF stuffing paths)."""
    from oracle import jpeg_oracle as J
    from hyres_b200 import ops
    H, W = 32, 48
    chk = ((np.add.outer(np.arange(H), np.arange(W)) & 1) * 255).astype(np.uint8)
    imgs = [np.full((H, W, 3), 255, np.uint8), np.zeros((H, W, 3), np.uint8), np.stack([chk, 255 - chk, chk], -1)]
    for img in imgs:
        for q in (1, 100):
            r = ops.jpeg_forward(_as_float(img).cuda(), q, want_scan=True)
            ref = J.encode(img, q)
            assert int(r["sizes"][0]) == len(ref)
            assert ops.jpeg_assemble(r["words"][0].cpu().numpy(), int(r["nbits"][0]), H, W, q) == ref
            want = torch.from_numpy(J.roundtrip(img, q)).permute(2, 0, 1).float()[None] / 255.0
            assert torch.equal(r["decoded"].cpu(), want)
