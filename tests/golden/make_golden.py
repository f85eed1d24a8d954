"""Generates tests/golden/*.npz from the REAL reference sources (run in the build container only).

    python tests/golden/make_golden.py [--reference /root/reference]

The reference (tmkhang1999/HyRES-...) is pure Python, but `import models` fails here because
compressai / PyTurboJPEG / timm are not installed.  This script therefore
executes the reference's *own in-tree files* --

    models/checkerboard.py, models/hyres.py, models/layers/{attention,checkerboard,common,
    enhancement}.py, models/utils/{quantization,turbo_jpeg_compression}.py, src/losses/rd_loss.py

-- unmodified, from where they lie under /root/reference, after registering stand-in modules for
the three absent third-party packages:

  * ``compressai.*``  -> the restatements in oracle/hyres_oracle.py (EntropyBottleneck,
    GaussianConditional, GDN, ResidualBottleneckBlock, conv/deconv, quantize_ste,
    CompressionModel, update_registered_buffers) and oracle/rans_oracle.c (``compressai.ans``);
  * ``turbojpeg``     -> a TurboJPEG class over cv2/libjpeg-turbo with PyTurboJPEG's defaults;
  * ``src.losses.vgg16`` -> an unused placeholder (alpha = 0 never calls it).

What the fixtures pin: everything the reference's in-tree code does (checkerboard split quirks
Q1-Q3, parameter-head channel order, decompress clamp, refine network, RD loss, wrapper
arithmetic) is produced by the reference's own code.  What they do NOT pin: the arithmetic
inside compressai itself (entropy models, GDN, rANS) -- that stays "parity unpinned" because
neither its source nor a wheel is available here.

Weights are 10.4 M floats, so fixtures store the seed and a SHA-256 of the regenerated state
dict; tests regenerate the weights from the seed and skip if the digest differs (different
torch build).  Nothing under tests/ reads /root/reference at run time.
"""
import argparse
import hashlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import hyres_oracle as O  # noqa: E402


def state_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        v = sd[k]
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def install_stubs(ref):
    """Register stand-ins for compressai / turbojpeg and namespace packages for the reference."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _RansEncoder:
        def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
            width = max(len(c) for c in cdfs)
            table = np.zeros((len(cdfs), width), dtype=np.int32)
            for i, c in enumerate(cdfs):
                table[i, : len(c)] = c
            return O.rans_encode_with_indexes(symbols, indexes, table, cdfs_sizes, offsets)

    class _RansDecoder:
        def decode_with_indexes(self, string, indexes, cdfs, cdfs_sizes, offsets):
            width = max(len(c) for c in cdfs)
            table = np.zeros((len(cdfs), width), dtype=np.int32)
            for i, c in enumerate(cdfs):
                table[i, : len(c)] = c
            return O.rans_decode_with_indexes(string, indexes, table, cdfs_sizes, offsets).tolist()

    mod("compressai")
    mod("compressai.ans", RansEncoder=_RansEncoder, RansDecoder=_RansDecoder)
    mod("compressai.entropy_models", EntropyBottleneck=O.EntropyBottleneck, GaussianConditional=O.GaussianConditional)
    mod("compressai.layers", GDN=O.GDN)
    mod("compressai.models", CompressionModel=O.CompressionModel)
    mod("compressai.models.base", CompressionModel=O.CompressionModel)
    mod("compressai.models.sensetime", ResidualBottleneckBlock=O.ResidualBottleneckBlock)
    mod("compressai.models.utils", conv=O.conv, deconv=O.deconv,
        update_registered_buffers=O._update_registered_buffers)
    mod("compressai.ops", quantize_ste=O.quantize_ste)

    class TurboJPEG:  # PyTurboJPEG 1.7.x defaults: pixel_format BGR, subsample 4:2:2
        def __init__(self, lib_path=None):
            pass

        def encode(self, img, quality=85):
            import cv2
            ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, int(quality),
                                                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
            assert ok
            return enc.tobytes()

        def decode(self, buf):
            import cv2
            return cv2.imdecode(np.frombuffer(buf, dtype=np.uint8), cv2.IMREAD_COLOR)

    mod("turbojpeg", TurboJPEG=TurboJPEG)

    # namespace packages so that the reference's package __init__ files (which import elic/timm,
    # lpips, ...) are not executed; only the files on the hot path are.
    for name, sub in (("models", "models"), ("models.utils", "models/utils"), ("src", "src"),
                      ("src.losses", "src/losses")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(ref, sub)]
        sys.modules[name] = m
    mod("src.losses.vgg16", VGGLoss=lambda *a, **k: None)
    sys.path.insert(0, ref)


def npz_save(path, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        if isinstance(v, (bytes, bytearray)):
            v = np.frombuffer(bytes(v), dtype=np.uint8)
        out[k] = np.asarray(v)
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def golden_layers(ref_layers, ref_enh, ref_quant):
    """In-tree leaf modules, no third-party arithmetic at all (fully pinned)."""
    g = torch.Generator().manual_seed(11)
    out = {}
    torch.manual_seed(101)
    att = ref_layers.AttentionBlock(32).eval()
    x = torch.randn(1, 32, 8, 8, generator=g)
    out["att_x"], out["att_y"] = x, att(x)
    for k, v in att.state_dict().items():
        out["att_sd." + k] = v
    torch.manual_seed(102)
    cm = ref_layers.CheckboardMaskedConv2d(8, 16, kernel_size=5, padding=2, stride=1).eval()
    x = torch.randn(1, 8, 6, 6, generator=g)
    w_before = cm.weight.detach().clone()
    out["cm_x"], out["cm_y"] = x, cm(x)
    out["cm_w_before"], out["cm_w_after"], out["cm_b"], out["cm_mask"] = w_before, cm.weight, cm.bias, cm.mask
    torch.manual_seed(103)
    rf = ref_enh.MultiScaleRefine(in_channels=3, mid_channels=64).eval()
    with torch.no_grad():
        for p in rf.parameters():  # make the SE / attention branches non-trivial
            if p.dim() == 1 and p.numel() == 1:
                p.fill_(0.2)
    x = torch.rand(1, 3, 16, 16, generator=g)
    out["rf_x"], out["rf_y"] = x, rf(x)
    for k, v in rf.state_dict().items():
        out["rf_sd." + k] = v
    qz = ref_quant.Quantizer()
    x = torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 0.49999997, -0.3, 3.7, 1e-9])
    out["q_x"], out["q_ste"], out["q_round"] = x, qz.quantize(x, "ste"), qz.quantize(x, "other")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    torch.set_num_threads(1)  # fixed summation order inside oneDNN for the fixtures
    install_stubs(a.reference)
    ref_layers = importlib.import_module("models.layers")
    ref_enh = importlib.import_module("models.layers.enhancement")
    ref_quant = importlib.import_module("models.utils.quantization")
    ref_cb = importlib.import_module("models.checkerboard")
    ref_hy = importlib.import_module("models.hyres")
    ref_rd = importlib.import_module("src.losses.rd_loss")

    with torch.no_grad():
        npz_save(os.path.join(HERE, "layers.npz"), **golden_layers(ref_layers, ref_enh, ref_quant))

        # ---- full codec + wrapper: the reference's classes on oracle-provided compressai parts ----
        onet = O.make_model(seed=1926, wrapper=True, lively=True)
        sd = onet.state_dict()
        digest = state_digest(sd)
        rnet = ref_hy.ResidualJPEGCompression(jpeg_quality=1)
        rnet.residual_model.load_state_dict({k[len("residual_model."):]: v for k, v in sd.items()
                                             if k.startswith("residual_model.")})
        # the reference's own load_state_dict cannot load "refine.*" keys (models/hyres.py:150-162
        # passes the prefixed keys on); load the sub-module directly.
        rnet.refine.load_state_dict({k[len("refine."):]: v for k, v in sd.items() if k.startswith("refine.")})
        rnet.eval()
        rcodec = rnet.residual_model

        for tag, (H, W) in (("codec64", (64, 64)), ("codec96x160", (96, 160))):
            x = O.synthetic_image(1, H, W, seed=5)
            jpeg_dec, jpeg_bpp = onet.jpeg(x)
            residual = x - jpeg_dec
            fwd = rcodec(residual)
            cmp_ = rcodec.compress(residual)
            dec = rcodec.decompress(cmp_["strings"], cmp_["shape"])
            # intermediates via the reference's own sub-modules (same call order as compress())
            y = rcodec.g_a(residual)
            z = rcodec.h_a(y)
            z_hat = rcodec.entropy_bottleneck.decompress(cmp_["strings"][1], z.size()[-2:])
            latent = rcodec.h_s(z_hat)
            pa = rcodec.param_aggregation(torch.cat([latent, torch.zeros_like(latent)], 1))
            s_a, m_a = pa.chunk(2, 1)
            ya = rcodec._split_tensor(y, "anchor")
            gc = rcodec.gaussian_conditional
            idx_a = gc.build_indexes(s_a)
            sym_a = gc.quantize(ya, "symbols", m_a)
            ya_hat = rcodec._decompress_part(cmp_["strings"][0][0], s_a, m_a)
            ctx = rcodec.context_prediction(ya_hat)
            pna = rcodec.param_aggregation(torch.cat([latent, ctx], 1))
            s_na, m_na = pna.chunk(2, 1)
            yna = rcodec._split_tensor(y, "non_anchor")
            idx_na = gc.build_indexes(s_na)
            sym_na = gc.quantize(yna, "symbols", m_na)
            med = rcodec.entropy_bottleneck._get_medians().detach().reshape(1, -1, 1, 1)
            sym_z = rcodec.entropy_bottleneck.quantize(z, "symbols", med)
            # wrapper forward + RD loss with the JPEG stage injected the way models/hyres.py:45-53 feeds it
            rnet.jpeg.forward = lambda _x, _j=(jpeg_dec, jpeg_bpp): _j
            wf = rnet(x)
            crit = ref_rd.RateDistortionLoss(lmbda=0.008, alpha=0)
            N_, _, H_, W_ = x.size()
            npx = N_ * H_ * W_
            import math
            bpp_y = torch.log(wf["likelihoods"]["y"]).sum() / (-math.log(2) * npx)
            bpp_z = torch.log(wf["likelihoods"]["z"]).sum() / (-math.log(2) * npx)
            mse = torch.nn.functional.mse_loss(wf["x_hat"], x) * 255 ** 2
            del crit  # (its forward instantiates the VGG term; the three sums above are rd_loss.py:23-26,39)
            npz_save(os.path.join(HERE, tag + ".npz"),
                     seed=1926, image_seed=5, state_digest=np.frombuffer(digest.encode(), dtype=np.uint8),
                     x=x, jpeg_decoded=jpeg_dec, jpeg_bpp=jpeg_bpp, residual=residual,
                     y=y, z=z, params_a=pa, params_na=pna,
                     sym_z=sym_z.to(torch.int16), sym_a=sym_a.to(torch.int16), sym_na=sym_na.to(torch.int16),
                     idx_a=idx_a.to(torch.uint8), idx_na=idx_na.to(torch.uint8),
                     str_a=cmp_["strings"][0][0][0], str_na=cmp_["strings"][0][1][0], str_z=cmp_["strings"][1][0],
                     shape=np.array(list(cmp_["shape"])),
                     fwd_x_hat=fwd["x_hat"], fwd_lik_y=fwd["likelihoods"]["y"], fwd_lik_z=fwd["likelihoods"]["z"],
                     dec_x_hat=dec["x_hat"],
                     w_x_hat=wf["x_hat"], w_residual_hat=wf["residual_hat"],
                     bpp_y=bpp_y, bpp_z=bpp_z, mse255=mse)

        # ---- cfg1 (BASELINE.json configs[0]): 256x256 synthetic residual, summary only ----
        x = O.synthetic_residual(1, 256, 256)
        fwd = rcodec(x)
        cmp_ = rcodec.compress(x)
        h = hashlib.sha256()
        for s in (cmp_["strings"][0][0][0], cmp_["strings"][0][1][0], cmp_["strings"][1][0]):
            h.update(s)
        npz_save(os.path.join(HERE, "cfg1_summary.npz"),
                 state_digest=np.frombuffer(digest.encode(), dtype=np.uint8),
                 x_hat_sum=fwd["x_hat"].double().sum(), x_hat_abs_sum=fwd["x_hat"].double().abs().sum(),
                 log2_lik_y=fwd["likelihoods"]["y"].double().log2().sum(),
                 log2_lik_z=fwd["likelihoods"]["z"].double().log2().sum(),
                 string_bytes=np.array([len(cmp_["strings"][0][0][0]), len(cmp_["strings"][0][1][0]),
                                        len(cmp_["strings"][1][0])]),
                 strings_sha256=np.frombuffer(h.hexdigest().encode(), dtype=np.uint8),
                 x_hat_center=fwd["x_hat"][0, :, 120:136, 120:136])

        # ---- entropy-coder known-answer vectors (oracle C coder; includes bypass escapes) ----
        gc = onet.residual_model.gaussian_conditional
        rng = np.random.default_rng(3)
        n = 4096
        idx = rng.integers(0, 64, size=n).astype(np.int32)
        sig = gc.scale_table.numpy()[idx]
        sym = np.round(rng.standard_normal(n) * sig * 1.3).astype(np.int32)
        sym[::97] += 4000  # out-of-range values -> bypass coding
        sym[5::131] -= 3000
        s = O.rans_encode_with_indexes(sym, idx, gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy())
        back = O.rans_decode_with_indexes(s, idx, gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy())
        assert (back == sym).all()
        npz_save(os.path.join(HERE, "rans_kat.npz"), symbols=sym, indexes=idx.astype(np.uint8), string=s,
                 cdf_rows_sha256=np.frombuffer(hashlib.sha256(gc._quantized_cdf.numpy().tobytes()).hexdigest().encode(),
                                               dtype=np.uint8),
                 cdf_length=gc._cdf_length, offset=gc._offset,
                 cdf_row0=gc._quantized_cdf[0, :16], cdf_row63_head=gc._quantized_cdf[63, :8])


if __name__ == "__main__":
    main()
