"""Training step of the B200 path (BASELINE.json configs[4]; src/utils/engine.py:29-90).

The reference trains with ``loss.backward()`` through cuDNN / ATen under fp16 autocast.  Here the same graph --
``ResidualJPEGCompression.forward`` (models/hyres.py:23-77, models/checkerboard.py:90-147) followed by
``RateDistortionLoss`` (src/losses/rd_loss.py:18-44) -- is built from ``torch.autograd.Function`` nodes whose
arithmetic runs in this library's kernels:

  * every convolution's forward is the tcgen05 implicit GEMM (``csrc/conv_tc.cu``), NHWC bf16 activations, fp32
    accumulation, weights re-packed on the device from the fp32 master parameters each step
    (``hyres_conv_update_device``);
  * its data gradient is the same kernel on the transposed problem: a stride-1 convolution with flipped taps and
    swapped channel roles, the k5 s2 transposed convolution for a k5 s2 convolution and vice versa (a transposed
    convolution is by definition the data gradient of the convolution with the same weight tensor);
  * its weight / bias gradient is the position-reduction GEMM ``dW[co, ci, r, s] = sum_pos g[pos, co] x[pos + tap, ci]``
    (``csrc/wgrad.cu`` when the layer shape is supported, ATen's convolution_backward otherwise).

The element-wise pieces between convolutions (GDN, gates, PReLU, the straight-through / noise quantisers, the erfc
likelihood with compressai's LowerBound gradient rule, the factorised prior, SE / bilinear / spatial attention of
MultiScaleRefine, the loss) are ordinary differentiable tensor expressions in the reference's own formulation,
which is what lets autograd differentiate them; they hold under 2 % of the step's FLOPs.

Data parallelism replaces ``nn.DataParallel`` (src/training.py:211-212): one process per GPU, gradients are
all-reduced over NCCL in buckets that are launched from autograd hooks while the backward pass is still running.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from ._lib import ACT_NONE, ACT_RELU, HYRES_CONV, HYRES_DECONV_K5S2
from .entropy import _LowerBoundFn

BF16 = torch.bfloat16


# --------------------------------------------------------------------------------------------------------------
# convolution node
# --------------------------------------------------------------------------------------------------------------
class TrainConv:
    """Forward / data-gradient / weight-gradient launches of one convolution geometry (channel counts are multiples
    of 8 here; 3-channel ends of the network are zero-padded to 8 by the caller)."""

    def __init__(self, kind, w_shape, stride=1, pad=0, dil=1, tap_mask=None):
        self.kind, self.w_shape, self.stride, self.pad, self.dil = kind, tuple(w_shape), stride, pad, dil
        self.tap_mask = tap_mask
        self.fwd = ops.ConvLayer(self.w_shape, None, kind=kind, stride=stride, pad=pad, dil=dil, tap_mask=tap_mask)
        self._dgrad = None
        self._unmasked = None
        if kind == HYRES_DECONV_K5S2:
            self.cin, self.cout = w_shape[0], w_shape[1]
        else:
            self.cout, self.cin = w_shape[0], w_shape[1]

    def dgrad_layer(self):
        if self._dgrad is None:
            R, S = self.w_shape[2:]
            if self.kind == HYRES_DECONV_K5S2:  # data gradient of a transposed conv = the conv with the same weight
                self._dgrad = ops.ConvLayer(self.w_shape, None, kind=HYRES_CONV, stride=2, pad=2, dil=1)
            elif self.stride == 2:              # ... and of a k5 s2 conv = the transposed conv with the same weight
                self._dgrad = ops.ConvLayer(self.w_shape, None, kind=HYRES_DECONV_K5S2)
            else:                               # stride 1 'same': flipped taps, channel roles swapped
                mask = None if self.tap_mask is None else self.tap_mask.flip(0, 1).contiguous()
                self._dgrad = ops.ConvLayer((self.cin, self.cout, R, S), None, kind=HYRES_CONV, stride=1, pad=self.pad,
                                            dil=self.dil, tap_mask=mask)
        return self._dgrad

    def wgrad_layer(self):
        """The layer whose tap table drives the weight-gradient kernel.  A masked convolution uses the unmasked
        geometry: the reference masks ``weight.data`` (models/layers/checkerboard.py:47), not the gradient, so dead
        taps do receive a gradient -- it never moves the forward pass (they are re-masked on every call) but it
        enters ``clip_grad_norm_`` (src/utils/engine.py:71) and therefore the size of every other update."""
        if self.tap_mask is None:
            return self.fwd
        if self._unmasked is None:
            self._unmasked = ops.ConvLayer(self.w_shape, None, kind=self.kind, stride=self.stride, pad=self.pad,
                                           dil=self.dil)
        return self._unmasked

    def dgrad_weight(self, w):
        if self.kind == HYRES_DECONV_K5S2 or self.stride == 2:
            return w
        return w.flip(2, 3).transpose(0, 1).contiguous()


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, tc, relu, out_f32):
        w = weight.detach().float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        tc.fwd.update_device(w, b)
        if out_f32:
            _, _, y = tc.fwd(x, act=ACT_RELU if relu else ACT_NONE, out_bf16=False, out_f32="nhwc")
        else:
            y, _, _ = tc.fwd(x, act=ACT_RELU if relu else ACT_NONE)
        ctx.tc, ctx.relu, ctx.has_bias = tc, relu, bias is not None
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, y = ctx.saved_tensors
        tc = ctx.tc
        if ctx.relu:
            g = g * (y > 0)
        g16 = g.to(BF16).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            layer = tc.dgrad_layer()
            layer.update_device(tc.dgrad_weight(w), None)
            dx, _, _ = layer(g16)
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = conv_wgrad(tc, x, g16, want_bias=ctx.has_bias)
        return dx, dw, db, None, None, None


_WGRAD_NATIVE = True


def conv_wgrad(tc, x, g16, want_bias=True):
    """-> (dW fp32 in the layer's PyTorch weight layout, db fp32 or None).  x: the layer's input (bf16 NHWC),
    g16: gradient of its output (bf16 NHWC)."""
    if _WGRAD_NATIVE and ops.wgrad_supported(tc):
        return ops.conv_wgrad(tc, x, g16, want_bias)
    xn, gn = x.permute(0, 3, 1, 2), g16.permute(0, 3, 1, 2)
    wq = torch.empty(tc.w_shape, dtype=BF16, device=x.device)
    transposed = tc.kind == HYRES_DECONV_K5S2
    _, dw, db = torch.ops.aten.convolution_backward(
        gn, xn, wq, [tc.cout], [tc.stride, tc.stride], [tc.pad, tc.pad], [tc.dil, tc.dil], transposed,
        [1, 1] if transposed else [0, 0], 1, [False, True, False])
    return dw.float(), (ops.colsum_bf16(g16) if want_bias else None)


def wgrad_native_active():
    return bool(_WGRAD_NATIVE and getattr(ops.L.lib(), "hyres_wgrad_supported", None) is not None)


def conv(x, weight, bias, tc, relu=False, out_f32=False):
    return _ConvFn.apply(x, weight, bias, tc, relu, out_f32)


def _lower_bound(x, bound):
    # torch.full is a fill kernel: no host-to-device copy, so the step stays CUDA-graph capturable
    return _LowerBoundFn.apply(x, torch.full((1,), float(bound), dtype=x.dtype, device=x.device))


def _ste_round(x):
    return (torch.round(x) - x).detach() + x


# --------------------------------------------------------------------------------------------------------------
# the training graph
# --------------------------------------------------------------------------------------------------------------
class TrainGraph:
    """Differentiable forward of ``ResidualJPEGCompression`` on the GPU (activations NHWC)."""

    def __init__(self, net):
        self.net = net
        self.codec = net.residual_model
        self._tc = {}
        # host copy of the scale lower bound (a device buffer): reading it per step would be a device-to-host sync
        self._scale_bound = float(self.codec.gaussian_conditional.scale_bound.cpu())

    # -- layer plumbing --
    def _node(self, key, kind, w_shape, stride=1, pad=0, dil=1, tap_mask=None):
        tc = self._tc.get(key)
        if tc is None:
            tc = self._tc[key] = TrainConv(kind, w_shape, stride, pad, dil, tap_mask)
        return tc

    def _conv(self, x, m, relu=False, out_f32=False, key=None):
        """nn.Conv2d / nn.ConvTranspose2d holder -> conv node (3-channel ends padded to 8 channels)."""
        w, b = m.weight, m.bias
        transposed = isinstance(m, torch.nn.ConvTranspose2d)
        kind = HYRES_DECONV_K5S2 if transposed else HYRES_CONV
        cin_dim, cout_dim = (0, 1) if transposed else (1, 0)
        cout = w.shape[cout_dim]
        pad_in, pad_out = (-w.shape[cin_dim]) % 8, (-cout) % 8
        if pad_in or pad_out:
            # F.pad lists the last dimension first: (S, S, R, R, dim 1 lo / hi, dim 0 lo / hi)
            w = F.pad(w, (0, 0, 0, 0, 0, pad_out, 0, pad_in) if transposed else (0, 0, 0, 0, 0, pad_in, 0, pad_out))
            if b is not None and pad_out:
                b = F.pad(b, (0, pad_out))
        if pad_in:
            x = F.pad(x, (0, pad_in))
        tc = self._node(key or id(m), kind, w.shape, 2 if transposed else m.stride[0], 2 if transposed else m.padding[0],
                        1 if transposed else m.dilation[0])
        y = conv(x.contiguous(), w, b, tc, relu, out_f32)
        return y[..., :cout] if pad_out else y

    def _gdn(self, x, m):
        gamma, beta = m.effective()  # non-negative reparametrisation: autograd carries the gradient to gamma / beta
        tc = self._node(id(m), HYRES_CONV, gamma.shape)
        xf = x.float()
        n = conv((xf * xf).to(BF16), gamma, beta, tc, False, True)
        return (xf * (torch.sqrt(n) if m.inverse else torch.rsqrt(n))).to(BF16)

    def _ru(self, x, c1, c2, c3, final_relu):
        t = self._conv(x, c1, relu=True)
        t = self._conv(t, c2, relu=True)
        t = self._conv(t, c3) + x
        return torch.relu(t) if final_relu else t

    def _rbb(self, x, m):
        return self._ru(x, m.conv1, m.conv2, m.conv3, False)

    def _attn(self, x, m, out_f32=False):
        a = x
        for u in m.conv_a:
            a = self._ru(a, u.conv[0], u.conv[2], u.conv[4], True)
        b = x
        for i in range(3):
            u = m.conv_b[i]
            b = self._ru(b, u.conv[0], u.conv[2], u.conv[4], True)
        b = self._conv(b, m.conv_b[3], out_f32=True)
        out = a.float() * torch.sigmoid(b) + x.float()
        return out if out_f32 else out.to(BF16)

    def _seq(self, x, seq, out_f32_last=False):
        """h_a / h_s style nn.Sequential of convs and ReLUs."""
        mods = list(seq)
        i = 0
        while i < len(mods):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
            last = i + (2 if relu else 1) >= len(mods)
            x = self._conv(x, mods[i], relu=relu, out_f32=out_f32_last and last)
            i += 2 if relu else 1
        return x

    # -- stages --
    def g_a(self, res_nhwc):
        ga = self.codec.g_a
        t = self._conv(res_nhwc, ga[0])
        t = self._gdn(t, ga[1])
        t = self._rbb(t, ga[2])
        t = self._attn(t, ga[3])
        t = self._conv(t, ga[4])
        t = self._gdn(t, ga[5])
        t = self._rbb(t, ga[6])
        t = self._conv(t, ga[7])
        return self._attn(t, ga[8], out_f32=True)  # y fp32 NHWC

    def g_s(self, y_hat16):
        gs = self.codec.g_s
        t = self._attn(y_hat16, gs[0])
        t = self._conv(t, gs[1])
        t = self._rbb(t, gs[2])
        t = self._gdn(t, gs[3])
        t = self._conv(t, gs[4])
        t = self._attn(t, gs[5])
        t = self._rbb(t, gs[6])
        t = self._gdn(t, gs[7])
        return self._conv(t, gs[8], out_f32=True)  # fp32 NHWC [B,H,W,3]

    def _head(self, latent16, ctx16):
        pa = self.codec.param_aggregation
        t = torch.cat([latent16, ctx16], dim=-1)
        return self._seq(t, pa, out_f32_last=True)  # fp32 NHWC [B,h,w,2M]: scales | means

    def _context(self, yq16):
        cp = self.codec.context_prediction
        tc = self._tc.get(id(cp))
        if tc is None:
            mask2d = (cp.mask[0, 0] != 0).to(torch.uint8).cpu()
            tc = self._node(id(cp), HYRES_CONV, cp.weight.shape, 1, cp.padding[0], 1, mask2d)
        return conv(yq16.contiguous(), cp.weight, cp.bias, tc, False, False)

    def _eb(self, z, noise_fn, training, noisequant):
        """EntropyBottleneck.forward (compressai) + models/checkerboard.py:96-101 on z fp32 NHWC."""
        eb = self.codec.entropy_bottleneck
        B, h, w, C = z.shape
        values = z.permute(3, 0, 1, 2).reshape(C, 1, -1)
        med = eb._get_medians()
        if training:
            outputs = values + noise_fn(values.shape, "z")
        else:
            outputs = torch.round(values.detach() - med) + med
        lower = eb._logits_cumulative(outputs - 0.5, stop_gradient=False)
        upper = eb._logits_cumulative(outputs + 0.5, stop_gradient=False)
        lik = _lower_bound(torch.sigmoid(upper) - torch.sigmoid(lower), eb.likelihood_bound)
        lik = lik.reshape(C, B, h, w).permute(1, 0, 2, 3)  # NCHW
        if noisequant:
            z_hat = outputs.reshape(C, B, h, w).permute(1, 2, 3, 0)
        else:
            m = med.reshape(1, 1, 1, C)
            z_hat = _ste_round(z - m) + m
        return z_hat, lik

    def _gc_likelihood(self, y, scales, means, noise_fn, training):
        """GaussianConditional.forward (compressai) via models/checkerboard.py:140-142; NHWC in, NCHW out."""
        gc = self.codec.gaussian_conditional
        outputs = y + noise_fn(y.shape, "y_lik") if training else torch.round(y.detach() - means) + means
        values = torch.abs(outputs - means)
        s = _lower_bound(scales, self._scale_bound)
        c = -(2 ** -0.5)
        upper = 0.5 * torch.erfc(c * ((0.5 - values) / s))
        lower = 0.5 * torch.erfc(c * ((-0.5 - values) / s))
        lik = _lower_bound(upper - lower, gc.likelihood_bound)
        return lik.permute(0, 3, 1, 2)

    def refine(self, x0):
        """MultiScaleRefine.forward (models/layers/enhancement.py:87-112): x0 fp32 NCHW -> refined fp32 NCHW."""
        r = self.net.refine

        def prelu(t, m):
            return F.prelu(t, m.weight.to(t.dtype))

        def down(t, s):
            return F.interpolate(t.permute(0, 3, 1, 2), scale_factor=s, mode="bilinear",
                                 align_corners=False).permute(0, 2, 3, 1).contiguous()

        def up(t, size):
            return F.interpolate(t.permute(0, 3, 1, 2), size=size, mode="bilinear",
                                 align_corners=False).permute(0, 2, 3, 1).contiguous()

        def block(t, seq):
            t = prelu(self._conv(t, seq[0]), seq[1])
            return prelu(self._conv(t, seq[2]), seq[3])

        x = x0.permute(0, 2, 3, 1).to(BF16)
        feat = prelu(self._conv(x, r.conv_in), r.act_in)
        pooled = feat.float().mean(dim=(1, 2))
        se = torch.sigmoid(F.linear(torch.relu(F.linear(pooled, r.se_block.fc[0].weight)), r.se_block.fc[2].weight))
        feat = (feat.float() * se[:, None, None, :]).to(BF16)
        size = feat.shape[1:3]
        f1 = block(feat, r.scale1)
        f2 = up(block(down(feat, 0.5), r.scale2), size)
        f3 = up(block(down(feat, 0.25), r.scale3), size)
        multi = torch.cat([f1, f2, f3], dim=-1)
        mf = multi.float()
        stats = torch.stack([mf.mean(dim=-1), mf.amax(dim=-1)], dim=1)  # [B,2,H,W]
        attn = torch.sigmoid(F.conv2d(stats, r.spatial_att.conv.weight.float(), None, padding=3))[:, 0, :, :, None]
        multi = (mf * attn).to(BF16)
        h = prelu(self._conv(multi, r.fusion[0]), r.fusion[1])
        return self._conv(h, r.fusion[2], out_f32=True).permute(0, 3, 1, 2)

    # -- whole forward --
    def forward(self, x, noisequant=False, jpeg=None, noise_fn=None, training=True):
        """x: fp32 NCHW on the device.  ``jpeg=(jpeg_decoded, jpeg_bpp)`` injects the JPEG stage (otherwise it runs
        on the device).  ``noise_fn(shape, tag)`` supplies the U(-1/2, 1/2) noise tensors (tests inject the oracle's).
        Returns the dict of ``ResidualJPEGCompression.forward`` with differentiable tensors."""
        net, codec = self.net, self.codec
        dev = x.device
        if noise_fn is None:
            def noise_fn(shape, tag):
                return torch.empty(shape, device=dev, dtype=torch.float32).uniform_(-0.5, 0.5)
        with torch.no_grad():
            if jpeg is None:
                jpeg_decoded, jpeg_bpp = net.jpeg.forward_device(x)
            else:
                jpeg_decoded, jpeg_bpp = jpeg
                jpeg_decoded = jpeg_decoded.to(dev, torch.float32)
                jpeg_bpp = torch.as_tensor(float(jpeg_bpp), dtype=torch.float32, device=dev)
            codec.context_prediction.weight.data *= codec.context_prediction.mask  # Q4
        residual = x - jpeg_decoded
        M = codec.M
        y = self.g_a(residual.permute(0, 2, 3, 1).to(BF16))
        z = self._seq(y.to(BF16), codec.h_a, out_f32_last=True)
        z_hat, z_lik = self._eb(z, noise_fn, training, noisequant)
        latent = self._seq(z_hat.to(BF16).contiguous(), codec.h_s)
        B, h, w, _ = y.shape
        ii = torch.arange(h, device=dev).view(1, h, 1, 1)
        jj = torch.arange(w, device=dev).view(1, 1, w, 1)
        anchor = ((ii + jj) % 2 == 0).to(y.dtype)
        y_a, y_na = y * anchor, y * (1 - anchor)
        pa = self._head(latent, torch.zeros_like(latent))
        s_a, m_a = pa[..., :M], pa[..., M:]
        ya_hat = y_a + noise_fn(y_a.shape, "y_a") if noisequant else _ste_round(y_a - m_a) + m_a
        ctx = self._context(ya_hat.to(BF16))
        pna = self._head(latent, ctx)
        s_na, m_na = pna[..., :M], pna[..., M:]
        yna_hat = y_na + noise_fn(y_na.shape, "y_na") if noisequant else _ste_round(y_na - m_na) + m_na
        y_hat = ya_hat + yna_hat
        r_hat = self.g_s(y_hat.to(BF16).contiguous()).permute(0, 3, 1, 2)
        y_lik = self._gc_likelihood(y, s_a + s_na, m_a + m_na, noise_fn, training)
        x0 = jpeg_decoded + r_hat
        refined = self.refine(x0)
        x_hat = torch.clamp(x0 + refined, 0, 1)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "jpeg_bpp_loss": jpeg_bpp,
                "jpeg_decoded": jpeg_decoded, "residual": residual, "residual_hat": r_hat}


def rd_loss(output, target, lmbda):
    """src/losses/rd_loss.py:18-44 without the VGG term (alpha = 0)."""
    N, _, H, W = target.shape
    npx = N * H * W
    out = {"y_bpp_loss": torch.log(output["likelihoods"]["y"]).sum() / (-math.log(2) * npx),
           "z_bpp_loss": torch.log(output["likelihoods"]["z"]).sum() / (-math.log(2) * npx)}
    out["residual_bpp_loss"] = out["y_bpp_loss"] + out["z_bpp_loss"]
    out["bpp_loss"] = out["residual_bpp_loss"] + output["jpeg_bpp_loss"]
    out["mse_loss"] = F.mse_loss(output["x_hat"], target) * 255 ** 2
    out["loss"] = lmbda * out["mse_loss"] + out["bpp_loss"]
    return out


# --------------------------------------------------------------------------------------------------------------
# gradient all-reduce, launched from autograd hooks while backward is running
# --------------------------------------------------------------------------------------------------------------
class GradBuckets:
    """Flat fp32 buckets in reverse registration order (the order gradients become ready).  A parameter's
    post-accumulate hook copies its gradient into the bucket; the bucket's asynchronous all-reduce starts when
    its last gradient has arrived.  ``finish`` waits, averages and scatters the results back into ``.grad``."""

    def __init__(self, params, bucket_bytes=8 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = []  # (flat tensor, [(param, offset, numel)])
        cur, cur_n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self._close(cur)
                cur, cur_n = [], 0
        if cur:
            self._close(cur)
        self._where = {}
        for bi, (_, items) in enumerate(self.buckets):
            for p, off, n in items:
                self._where[p] = (bi, off, n)
        self._pending = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._hooks = ([p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
                       if self.world > 1 else [])
        self.reset()

    def _close(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        items, off = [], 0
        for p in ps:
            items.append((p, off, p.numel()))
            off += p.numel()
        self.buckets.append((flat, items))

    def reset(self):
        for bi, (_, items) in enumerate(self.buckets):
            self._pending[bi] = len(items)
            self._works[bi] = None

    def _on_grad(self, p):
        bi, off, n = self._where[p]
        flat = self.buckets[bi][0]
        flat[off:off + n].copy_(p.grad.reshape(-1))
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.world > 1:
            self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for the reductions and write the averaged gradients back.  Parameters that received no gradient this
        step contribute zeros (every rank must reduce every bucket)."""
        if self.world > 1:
            for bi, (flat, items) in enumerate(self.buckets):
                if self._works[bi] is None:
                    for p, off, n in items:
                        if p.grad is None:
                            flat[off:off + n].zero_()
                    self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            for bi, (flat, items) in enumerate(self.buckets):
                self._works[bi].wait()
                flat.div_(self.world)
                for p, off, n in items:
                    if p.grad is not None:
                        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()


class Trainer:
    """One optimisation step as ``train_one_epoch`` performs it (src/utils/engine.py:29-90): forward, RD loss,
    backward, gradient clipping, Adam step, then the auxiliary (quantile) loss and its own Adam step."""

    def __init__(self, net, lmbda=0.008, lr=1e-4, aux_lr=1e-3, clip_max_norm=1.0, bucket_bytes=8 << 20,
                 capturable=False):
        """``capturable``: build the Adam optimisers with device-side step counters so that ``capture`` can record the
        whole step into a CUDA graph."""
        self.net, self.lmbda, self.clip = net, lmbda, clip_max_norm
        self._graph = None
        self.graph = TrainGraph(net)
        named = dict(net.named_parameters())
        main = sorted(n for n, p in named.items() if not n.endswith(".quantiles") and p.requires_grad)
        aux = sorted(n for n, p in named.items() if n.endswith(".quantiles") and p.requires_grad)
        self.main_params = [named[n] for n in main]
        self.aux_params = [named[n] for n in aux]
        # src/utils/optimizers.py:27-34
        self.optimizer = torch.optim.Adam(self.main_params, lr=lr, betas=(0.9, 0.999), capturable=capturable)
        self.aux_optimizer = torch.optim.Adam(self.aux_params, lr=aux_lr, betas=(0.9, 0.999), capturable=capturable)
        self.capturable = capturable
        self.buckets = GradBuckets(self.main_params, bucket_bytes)
        self.aux_buckets = GradBuckets(self.aux_params, bucket_bytes)

    def capture(self, x, noisequant=True, warmup=3):
        """Record one whole optimisation step (forward, backward, gradient all-reduce, clipping, both Adam steps) on a
        static input buffer into a CUDA graph; ``step`` then replays it for inputs of that shape.  The step issues
        ~2 500 kernel launches for ~15 ms of device work, so launched eagerly it is bound by the host (measured 53 ms
        against 17 ms replayed at 16 x 256 x 256).  The warm-up steps run eagerly first: they create every packed layer,
        tap table and optimiser state the captured step reuses."""
        if not self.capturable:
            raise RuntimeError("Trainer(capturable=True) is required for capture()")
        self._graph = None
        static_x = x.detach().clone()
        # warm up on the stream the capture will use: cuBLAS / cuDNN (the factorised prior's small matmuls, the 7x7
        # attention conv) set up per-stream workspaces on first use, which must not happen while capturing
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step_impl(static_x, noisequant, None, None)
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            res = self._step_impl(static_x, noisequant, None, None)
        self._graph = (g, static_x, res, noisequant)
        return res

    def step(self, x, noisequant=True, jpeg=None, noise_fn=None):
        """-> dict of 0-d device tensors (loss, bpp_loss, mse_loss, ..., aux_loss) of this step."""
        if (self._graph is not None and jpeg is None and noise_fn is None and self._graph[3] == noisequant
                and tuple(x.shape) == tuple(self._graph[1].shape)):
            g, static_x, res, _ = self._graph
            static_x.copy_(x, non_blocking=True)
            g.replay()
            return res
        return self._step_impl(x, noisequant, jpeg, noise_fn)

    def _step_impl(self, x, noisequant, jpeg, noise_fn):
        net = self.net
        net.train()
        out = self.graph.forward(x, noisequant=noisequant, jpeg=jpeg, noise_fn=noise_fn, training=True)
        crit = rd_loss(out, x, self.lmbda)
        self.optimizer.zero_grad(set_to_none=True)
        self.aux_optimizer.zero_grad(set_to_none=True)
        crit["loss"].backward()
        self.buckets.finish()
        if self.clip > 0:
            torch.nn.utils.clip_grad_norm_(net.parameters(), self.clip)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        aux = net.aux_loss()
        aux.backward()
        self.aux_buckets.finish()
        self.aux_optimizer.step()
        self.aux_optimizer.zero_grad(set_to_none=True)
        res = {k: v.detach() for k, v in crit.items()}
        res["aux_loss"] = aux.detach()
        return res
