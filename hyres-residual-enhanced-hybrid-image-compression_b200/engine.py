"""Compiles the parameter holders into the fused sm_100a launch sequence.

One ``ops.ConvLayer`` per convolution of the reference graph (models/checkerboard.py:35-88,
models/layers/attention.py, models/layers/enhancement.py:60-85), with the surrounding
element-wise work folded into conv epilogues:

  * ReLU / PReLU / clamp, bias                      -> activation field of the epilogue
  * ResidualUnit / ResidualBottleneckBlock skips    -> HYRES_EPI_ADD (+ReLU)
  * AttentionBlock ``a*sigmoid(b)+x``               -> HYRES_EPI_GATE on conv_b.3
  * GDN / IGDN                                      -> one 1x1 GEMM launch that squares its own input tile
                                                       on chip (x0_square) and ends in HYRES_EPI_(I)GDN
  * SpatialAttention multiply                       -> HYRES_EPI_PIXSCALE on fusion.0
  * g_a.0 / refine.conv_in (3 input channels)       -> im2col + 1x1 GEMM
  * ``torch.cat([latent, ctx])``                    -> two-input conv (never materialised); the
                                                       anchor pass drops the all-zero half (K=384)

Activations are NHWC bf16; y, z, the entropy parameters and the final images stay fp32.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .ops import (ACT_CLAMP01, ACT_NONE, ACT_PRELU, ACT_RELU, EPI_ADD, EPI_GATE, EPI_GDN, EPI_IGDN, EPI_LINEAR,
                  EPI_PIXSCALE, HYRES_CONV, HYRES_DECONV_K5S2)


class _Bound:
    """A ConvLayer bound to the callable that extracts its (weight, bias) from the holders."""

    def __init__(self, getter, **geom):
        self.getter = getter
        w, b = getter()
        self.layer = ops.ConvLayer(w, b, **geom)

    def refresh(self):
        w, b = self.getter()
        self.layer.update(w, b)

    def __call__(self, *a, **k):
        return self.layer(*a, **k)


def _wb(m):
    return lambda: (m.weight, m.bias)


def _conv(m, mt=0):
    """nn.Conv2d holder -> bound layer (stride 1 'same' or stride 2 k5)."""
    return _Bound(_wb(m), kind=HYRES_CONV, stride=m.stride[0], pad=m.padding[0], dil=m.dilation[0])


def _deconv(m):
    return _Bound(_wb(m), kind=HYRES_DECONV_K5S2)


def _gdn(m):
    def get():
        g, b = m.effective()
        return g, b
    return _Bound(get, kind=HYRES_CONV)


_NO_BRANCH_STREAMS = bool(os.environ.get("HYRES_NO_BRANCH_STREAMS"))
_PRECISE_UNFUSED = bool(os.environ.get("HYRES_PRECISE_UNFUSED"))
_side = {}


def _side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=dev)
    return _side[key]


class _RU:
    """1x1 -> ReLU -> 3x3 -> ReLU -> 1x1 (+skip [+ReLU]): ResidualUnit and ResidualBottleneckBlock."""

    def __init__(self, c1, c2, c3, final_relu):
        self.c1, self.c2, self.c3 = _conv(c1), _conv(c2), _conv(c3)
        self.final_act = ACT_RELU if final_relu else ACT_NONE
        self.fused = ops.ru_supported(self.c1.layer, self.c2.layer, self.c3.layer)

    def layers(self):
        return [self.c1, self.c2, self.c3]

    def __call__(self, x, out_sq=False, cta_limit=0):
        if self.fused and not out_sq:
            return ops.ru_fused(x, self.c1.layer, self.c2.layer, self.c3.layer, self.final_act == ACT_RELU)
        a, _, _ = self.c1(x, act=ACT_RELU, cta_limit=cta_limit)
        b, _, _ = self.c2(a, act=ACT_RELU, cta_limit=cta_limit)
        o, sq, _ = self.c3(b, epi=EPI_ADD, aux0=x, act=self.final_act, out_sq=out_sq, cta_limit=cta_limit)
        return (o, sq) if out_sq else o


def _ru_from_unit(u):
    return _RU(u.conv[0], u.conv[2], u.conv[4], final_relu=True)


def _ru_from_rbb(r):
    return _RU(r.conv1, r.conv2, r.conv3, final_relu=False)


class _Attn:
    def __init__(self, m):
        self.a = [_ru_from_unit(u) for u in m.conv_a]
        self.b = [_ru_from_unit(m.conv_b[i]) for i in range(3)]
        self.gate = _conv(m.conv_b[3])

    def layers(self):
        out = []
        for r in self.a + self.b:
            out += r.layers()
        return out + [self.gate]

    def __call__(self, x, out_f32=None):
        # under CUDA-graph capture only: replayed, the two branches overlap (1-2 % of a step); launched eagerly the
        # host issues the 18 small kernels more slowly than they run and the split grids just halve each layer's SMs
        if not self.a[0].fused and not _NO_BRANCH_STREAMS and torch.cuda.is_current_stream_capturing():
            return self._two_streams(x, out_f32)
        a = x
        for r in self.a:
            a = r(a)
        b = x
        for r in self.b:
            b = r(b)
        o16, _, o32 = self.gate(b, epi=EPI_GATE, aux0=x, aux1=a, out_f32=out_f32)
        return (o16, o32) if out_f32 else o16

    def _two_streams(self, x, out_f32):
        """The C = 192 blocks (1/8 resolution: 768 tiles a layer, 5 per SM, latency-bound launches of 15-40 us): the
        two independent branches run side by side on two streams, each layer on half of the SMs, so one branch's
        prologues / tails / tile-count rounding hide under the other's tiles.  Fork and join are events, so the step
        stays CUDA-graph capturable."""
        cur = torch.cuda.current_stream(x.device)
        side = _side_stream(x.device)
        half = max(1, ops.sm_count() // 2)
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        a = b = x
        # launches alternate between the streams layer by layer, so both always have queued work
        for ra, rb in zip(self.a, self.b):
            with torch.cuda.stream(side):
                ta, _, _ = ra.c1(a, act=ACT_RELU, cta_limit=half)
            tb, _, _ = rb.c1(b, act=ACT_RELU, cta_limit=half)
            with torch.cuda.stream(side):
                ua, _, _ = ra.c2(ta, act=ACT_RELU, cta_limit=half)
            ub, _, _ = rb.c2(tb, act=ACT_RELU, cta_limit=half)
            with torch.cuda.stream(side):
                a, _, _ = ra.c3(ua, epi=EPI_ADD, aux0=a, act=ra.final_act, cta_limit=half)
            b, _, _ = rb.c3(ub, epi=EPI_ADD, aux0=b, act=rb.final_act, cta_limit=half)
        a.record_stream(cur)  # allocated on the side stream, consumed by the gate on the main stream
        join = torch.cuda.Event()
        join.record(side)
        cur.wait_event(join)
        o16, _, o32 = self.gate(b, epi=EPI_GATE, aux0=x, aux1=a, out_f32=out_f32)
        return (o16, o32) if out_f32 else o16


class CodecEngine:
    """g_a / g_s / h_a / h_s / context / parameter head of ``LightWeightCheckerboard``."""

    def __init__(self, model):
        self.model = model
        N, M = model.N, model.M
        if (N, M) != (128, 192):
            # geometry is generic in the kernels, but the im2col width / TMEM budgets were sized for this config
            if N % 64 or M % 64:
                raise ValueError("N and M must be multiples of 64")
        ga, gs = model.g_a, model.g_s

        def ga0():
            w = ga[0].weight  # [N,3,5,5] -> [N,128,1,1], k = (r*5+s)*3 + c
            w2 = torch.zeros(w.shape[0], 128, 1, 1, dtype=torch.float32)
            w2[:, :75, 0, 0] = w.detach().float().cpu().permute(0, 2, 3, 1).reshape(w.shape[0], 75)
            return w2, ga[0].bias

        self.ga0 = _Bound(ga0, kind=HYRES_CONV)
        self.ga1 = _gdn(ga[1])
        self.ga2 = _ru_from_rbb(ga[2])
        self.ga3 = _Attn(ga[3])
        self.ga4 = _conv(ga[4])
        self.ga5 = _gdn(ga[5])
        self.ga6 = _ru_from_rbb(ga[6])
        self.ga7 = _conv(ga[7])
        self.ga8 = _Attn(ga[8])

        self.gs0 = _Attn(gs[0])
        self.gs1 = _deconv(gs[1])
        self.gs2 = _ru_from_rbb(gs[2])
        self.gs3 = _gdn(gs[3])
        self.gs4 = _deconv(gs[4])
        self.gs5 = _Attn(gs[5])
        self.gs6 = _ru_from_rbb(gs[6])
        self.gs7 = _gdn(gs[7])
        self.gs8 = _deconv(gs[8])

        self.ha = [_conv(model.h_a[0]), _conv(model.h_a[2]), _conv(model.h_a[4])]
        self.hs = [_deconv(model.h_s[0]), _deconv(model.h_s[2]), _conv(model.h_s[4])]

        cp = model.context_prediction
        mask = (cp.mask[0, 0] != 0).to(torch.uint8)
        self.ctx = _Bound(_wb(cp), kind=HYRES_CONV, stride=1, pad=cp.padding[0], dil=1, tap_mask=mask)
        pa = model.param_aggregation
        self.head0_anchor = _Bound(_wb(pa[0]), kind=HYRES_CONV, cin0=2 * M, cin1=0)
        self.head0_full = _Bound(_wb(pa[0]), kind=HYRES_CONV, cin0=2 * M, cin1=2 * M)
        self.head1 = _conv(pa[2])
        self.head2 = _conv(pa[4])
        self._versions = None
        self._eb_cache = None

    # -- weight tracking --
    def _all_bound(self):
        out = [self.ga0, self.ga1, self.ga4, self.ga5, self.ga7, self.gs1, self.gs3, self.gs4, self.gs7, self.gs8,
               self.ctx, self.head0_anchor, self.head0_full, self.head1, self.head2]
        out += self.ha + self.hs
        for blk in (self.ga2, self.ga3, self.ga6, self.ga8, self.gs0, self.gs2, self.gs5, self.gs6):
            out += blk.layers()
        return out

    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def sync(self, force=False):
        """Re-pack weights if any parameter changed since the last launch."""
        key = self._version_key()
        if force or (self._versions is not None and key != self._versions):
            for b in self._all_bound():
                b.refresh()
            self._eb_cache = None
        self._versions = key

    def eb_params(self):
        if self._eb_cache is None:
            eb = self.model.entropy_bottleneck
            self._eb_cache = (eb.kernel_params(), eb.medians_vector())
        return self._eb_cache

    # -- stages (NHWC bf16 in / out unless noted) --
    def g_a(self, x, jpeg=None, want_residual=True):
        """x (and optionally jpeg): fp32 NCHW [B,3,H,W]; g_a runs on residual = x - jpeg.
        -> (y bf16 NHWC, y fp32 NHWC, residual fp32 NCHW or x)."""
        residual, t = ops.conv3ch(self.ga0.layer, 5, 2, x, jpeg, sign=-1, want_sum=want_residual)
        t, _, _ = self.ga1(t, epi=EPI_GDN, aux0=t, x0_square=True)
        t = self.ga2(t)
        t = self.ga3(t)
        t, _, _ = self.ga4(t)
        t, _, _ = self.ga5(t, epi=EPI_GDN, aux0=t, x0_square=True)
        t = self.ga6(t)
        t, _, _ = self.ga7(t)
        y16, y32 = self.ga8(t, out_f32="nhwc")
        return y16, y32, residual

    def h_a(self, y16):
        t, _, _ = self.ha[0](y16, act=ACT_RELU)
        t, _, _ = self.ha[1](t, act=ACT_RELU)
        _, _, z = self.ha[2](t, out_bf16=False, out_f32="nhwc")
        return z

    def h_s(self, zhat16):
        t, _, _ = self.hs[0](zhat16, act=ACT_RELU)
        t, _, _ = self.hs[1](t, act=ACT_RELU)
        t, _, _ = self.hs[2](t)
        return t

    def head(self, latent16, ctx16=None):
        if ctx16 is None:
            t, _, _ = self.head0_anchor(latent16, act=ACT_RELU)
        else:
            t, _, _ = self.head0_full(latent16, ctx16, act=ACT_RELU)
        t, _, _ = self.head1(t, act=ACT_RELU)
        _, _, p = self.head2(t, out_bf16=False, out_f32="nhwc")
        return p  # fp32 NHWC [B,h,w,2M]: scales | means

    def context(self, yq16):
        t, _, _ = self.ctx(yq16)
        return t

    def g_s(self, y_hat16, clamp=False):
        t = self.gs0(y_hat16)
        t, _, _ = self.gs1(t)
        t = self.gs2(t)
        t, _, _ = self.gs3(t, epi=EPI_IGDN, aux0=t, x0_square=True)
        t, _, _ = self.gs4(t)
        t = self.gs5(t)
        t = self.gs6(t)
        t, _, _ = self.gs7(t, epi=EPI_IGDN, aux0=t, x0_square=True)
        _, _, x = self.gs8(t, out_bf16=False, out_f32="nchw", act=ACT_CLAMP01 if clamp else ACT_NONE)
        return x  # fp32 NCHW [B,3,H,W]


class _PT:
    """An activation of the split-precision trunk: fp32 NHWC tensor + its bf16 parts ([..., nsplit*C])."""
    __slots__ = ("f32", "sp")

    def __init__(self, f32, sp):
        self.f32, self.sp = f32, sp


class PreciseTrunk:
    """The entropy-critical half of ``LightWeightCheckerboard`` -- g_a, h_a, h_s, context_prediction,
    param_aggregation (models/checkerboard.py:35-45,61-88) -- at fp32-equivalent precision.

    These layers decide the integer symbols and CDF indexes (models/checkerboard.py:159-165), which must equal the
    fp32 reference's; a bf16 trunk flips a few per cent of them.  Here activations stay fp32 in HBM, every
    convolution is a split-bf16 product on the tensor cores (``ops.ConvLayer(nsplit=3)``: 6 bf16 MMAs per MAC, fp32
    accumulation in TMEM, error ~2^-22 relative) and the element-wise work between two convolutions (ReLU, skip,
    gate, GDN) is IEEE fp32 (``ops.split_f32``).  ``compress`` / ``decompress`` always run this trunk; ``forward``
    uses it when the model's ``precision`` is "fp32x3"."""

    def __init__(self, model, nsplit=3):
        """``nsplit``: the format code of the trunk's activations / weights: 3 (three bf16 parts, 6 tensor-core
        products per MAC), 2 (two bf16 parts, 3 products, ~2^-16) or 2 | SPLIT_F16 (two IEEE half parts: 3 products
        at fp32-equivalent accuracy).  The x^2 operand of a GDN keeps three bf16 parts in the half format too
        (a square leaves the half range at |x| >= 256), so only the three gamma GEMMs pay 6 products there."""
        self.model, self.P = model, nsplit
        M = model.M
        P = nsplit
        self.Psq = 3 if (nsplit & ops.SPLIT_F16) else nsplit  # format of the GDN operand x^2

        def cv(m):
            return _Bound(_wb(m), kind=HYRES_CONV, stride=m.stride[0], pad=m.padding[0], dil=m.dilation[0], nsplit=P)

        def dc(m):
            return _Bound(_wb(m), kind=HYRES_DECONV_K5S2, nsplit=P)

        def gd(m):
            return _Bound(m.effective, kind=HYRES_CONV, nsplit=self.Psq)

        def ru(c1, c2, c3):
            return [cv(c1), cv(c2), cv(c3)]

        def ru_unit(u):
            return ru(u.conv[0], u.conv[2], u.conv[4])

        def rbb(r):
            return ru(r.conv1, r.conv2, r.conv3)

        def attn(m):
            return dict(a=[ru_unit(u) for u in m.conv_a], b=[ru_unit(m.conv_b[i]) for i in range(3)],
                        gate=cv(m.conv_b[3]))

        ga = model.g_a

        def ga0():
            w = ga[0].weight  # [N,3,5,5] -> [N,128,1,1], k = (r*5+s)*3 + c
            w2 = torch.zeros(w.shape[0], 128, 1, 1, dtype=torch.float32)
            w2[:, :75, 0, 0] = w.detach().float().cpu().permute(0, 2, 3, 1).reshape(w.shape[0], 75)
            return w2, ga[0].bias

        self.ga0 = _Bound(ga0, kind=HYRES_CONV, nsplit=P)
        self.ga1, self.ga2, self.ga3, self.ga4 = gd(ga[1]), rbb(ga[2]), attn(ga[3]), cv(ga[4])
        self.ga5, self.ga6, self.ga7, self.ga8 = gd(ga[5]), rbb(ga[6]), cv(ga[7]), attn(ga[8])
        self.ha = [cv(model.h_a[0]), cv(model.h_a[2]), cv(model.h_a[4])]
        self.hs = [dc(model.h_s[0]), dc(model.h_s[2]), cv(model.h_s[4])]
        cp = model.context_prediction
        mask = (cp.mask[0, 0] != 0).to(torch.uint8)
        self.ctx = _Bound(_wb(cp), kind=HYRES_CONV, stride=1, pad=cp.padding[0], dil=1, tap_mask=mask, nsplit=P)
        pa = model.param_aggregation
        self.head0_anchor = _Bound(_wb(pa[0]), kind=HYRES_CONV, cin0=2 * M, cin1=0, nsplit=P)
        self.head0_full = _Bound(_wb(pa[0]), kind=HYRES_CONV, cin0=2 * M, cin1=2 * M, nsplit=P)
        self.head1, self.head2 = cv(pa[2]), cv(pa[4])
        self._versions = None

    def _all_bound(self):
        out = [self.ga0, self.ga1, self.ga4, self.ga5, self.ga7, self.ctx, self.head0_anchor, self.head0_full,
               self.head1, self.head2] + self.ha + self.hs + self.ga2 + self.ga6
        for blk in (self.ga3, self.ga8):
            for r in blk["a"] + blk["b"]:
                out += r
            out.append(blk["gate"])
        return out

    def sync(self, force=False):
        key = tuple((p.data_ptr(), p._version) for p in self.model.parameters())
        if force or (self._versions is not None and key != self._versions):
            for b in self._all_bound():
                b.refresh()
        self._versions = key

    # -- building blocks --
    def _conv(self, layer, x_sp, x1_sp=None, relu=False, mode=ops.SPLIT_COPY, aux0=None, aux1=None, want_f32=False,
              want_split=True, square=False):
        """bf16 parts in -> _PT(fp32 NHWC or None, bf16 parts or None); bias, the fp32 element-wise stage that
        follows the convolution (skip / gate / GDN), ReLU and the split of the result all run in the epilogue."""
        if _PRECISE_UNFUSED:  # experiment: convolution -> fp32, element-wise stage + split as a second kernel
            _, _, o = layer(x_sp, x1_sp, act=ACT_RELU if (relu and mode == ops.SPLIT_COPY) else ACT_NONE,
                            out_bf16=False, out_f32="nhwc")
            if mode == ops.SPLIT_COPY and not square:
                f, sp = ops.split_f32(o, nsplit=self.P, want_f32=True, want_split=want_split)
                return _PT(f, sp)
            if mode == ops.SPLIT_COPY:  # square: parts of o*o, fp32 o
                _, sp = ops.split_f32(o, mode=ops.SPLIT_SQUARE, nsplit=self.Psq)
                return _PT(o, sp)
            f, sp = ops.split_f32(o, mode=mode, aux0=aux0, aux1=aux1, relu=relu, nsplit=self.P, want_f32=True,
                                  want_split=want_split)
            return _PT(f, sp)
        _, sp, o = layer(x_sp, x1_sp, act=ACT_RELU if relu else ACT_NONE, out_bf16=False,
                         out_f32="nhwc" if want_f32 else None, split_mode=mode, aux0_f32=aux0, aux1_f32=aux1,
                         out_split=True if want_split else None, split_square=square,
                         out_code=self.Psq if square else self.P)
        return _PT(o, sp)

    def split(self, x32, want_f32=True, **kw):
        f, sp = ops.split_f32(x32, nsplit=self.P, want_f32=want_f32, **kw)
        return _PT(f, sp)

    def _ru(self, x, layers, final_relu):
        c1, c2, c3 = layers
        a = self._conv(c1, x.sp, relu=True)
        b = self._conv(c2, a.sp, relu=True)
        return self._conv(c3, b.sp, mode=ops.SPLIT_ADD, aux0=x.f32, relu=final_relu, want_f32=True)

    def _attn(self, x, blk):
        a = x
        for r in blk["a"]:
            a = self._ru(a, r, True)
        b = x
        for r in blk["b"]:
            b = self._ru(b, r, True)
        return self._conv(blk["gate"], b.sp, mode=ops.SPLIT_GATE, aux0=x.f32, aux1=a.f32, want_f32=True)

    def _conv_gdn(self, conv, gdn, x_sp, inverse=False):
        """conv -> (I)GDN: the conv writes its fp32 result and the parts of its square, the 1x1 gamma GEMM over
        those ends in x * rsqrt(.) (or x * sqrt(.))."""
        t = self._conv(conv, x_sp, want_f32=True, square=True)
        return self._conv(gdn, t.sp, mode=ops.SPLIT_IGDN if inverse else ops.SPLIT_GDN, aux0=t.f32, want_f32=True)

    # -- stages --
    def g_a(self, x, jpeg=None):
        """x (and optionally jpeg): fp32 NCHW -> (y as _PT [B,H/8,W/8,M], residual fp32 NCHW or x)."""
        residual, a = ops.residual_im2col5s2_split(x, jpeg, nsplit=self.P)
        t = self._conv_gdn(self.ga0, self.ga1, a)
        t = self._ru(t, self.ga2, False)
        t = self._attn(t, self.ga3)
        t = self._conv_gdn(self.ga4, self.ga5, t.sp)
        t = self._ru(t, self.ga6, False)
        t = self._conv(self.ga7, t.sp, want_f32=True)
        return self._attn(t, self.ga8), residual

    def h_a(self, y):
        t = self._conv(self.ha[0], y.sp, relu=True)
        t = self._conv(self.ha[1], t.sp, relu=True)
        return self._conv(self.ha[2], t.sp, want_f32=True, want_split=False).f32  # z fp32 NHWC

    def h_s(self, zhat32):
        t = self.split(zhat32, want_f32=False)
        t = self._conv(self.hs[0], t.sp, relu=True)
        t = self._conv(self.hs[1], t.sp, relu=True)
        return self._conv(self.hs[2], t.sp)  # latent parts [B,h,w,P*2M]

    def head(self, latent, ctx=None):
        if ctx is None:
            t = self._conv(self.head0_anchor, latent.sp, relu=True)
        else:
            t = self._conv(self.head0_full, latent.sp, ctx.sp, relu=True)
        t = self._conv(self.head1, t.sp, relu=True)
        return self._conv(self.head2, t.sp, want_f32=True, want_split=False).f32  # fp32 NHWC [B,h,w,2M]: scales | means

    def context(self, yq32):
        t = self.split(yq32, want_f32=False)
        return self._conv(self.ctx, t.sp)


class RefineEngine:
    """``MultiScaleRefine`` (models/layers/enhancement.py:55-112)."""

    def __init__(self, refine):
        self.m = refine
        r = refine

        def conv_in():
            w = r.conv_in.weight  # [64,3,3,3] -> [64,64,1,1], k = (r*3+s)*3 + c
            w2 = torch.zeros(w.shape[0], 64, 1, 1, dtype=torch.float32)
            w2[:, :27, 0, 0] = w.detach().float().cpu().permute(0, 2, 3, 1).reshape(w.shape[0], 27)
            return w2, r.conv_in.bias

        self.conv_in = _Bound(conv_in, kind=HYRES_CONV)
        self.scales = [[_conv(s[0]), _conv(s[2])] for s in (r.scale1, r.scale2, r.scale3)]
        # fusion[0] (1x1, 192 -> 64) over cat([s1, up2(s2), up4(s3)]) * att is linear and per pixel, and bilinear
        # up-sampling commutes with a 1x1 convolution, so the 192-channel concat is never materialised:
        #   att * (W[:, :64] s1 + up2(W[:, 64:128] s2) + up4(W[:, 128:] s3)) + b
        # the two low-resolution products are small 1x1 layers; their up-sampling runs on the tensor cores inside
        # the full-resolution layer (ops.ConvLayer up_t2 / up_t3)
        def f0_slice(lo, with_bias):
            def get():
                w = r.fusion[0].weight[:, lo:lo + 64]
                return w, (r.fusion[0].bias if with_bias else None)
            return get
        self.fusion0 = [_Bound(f0_slice(0, True), kind=HYRES_CONV), _Bound(f0_slice(64, False), kind=HYRES_CONV),
                        _Bound(f0_slice(128, False), kind=HYRES_CONV)]
        self.fusion2 = _conv(r.fusion[2])
        self._versions = None
        self._small = None

    def _all_bound(self):
        return [self.conv_in, self.fusion2] + self.fusion0 + [c for s in self.scales for c in s]

    def sync(self, force=False):
        key = tuple((p.data_ptr(), p._version) for p in self.m.parameters())
        if force or (self._versions is not None and key != self._versions):
            for b in self._all_bound():
                b.refresh()
            self._small = None
        self._versions = key

    def small(self, dev):
        """fp32 side parameters: SE fc weights, 7x7 attention weights, PReLU slopes."""
        if self._small is None or self._small["dev"] != dev:
            r = self.m
            self._small = dict(
                dev=dev,
                fc1=r.se_block.fc[0].weight.detach().float().to(dev).contiguous(),
                fc2=r.se_block.fc[2].weight.detach().float().to(dev).contiguous(),
                w7=r.spatial_att.conv.weight.detach().float().to(dev).reshape(-1).contiguous(),
                slopes=[float(p.weight.detach().float().cpu().item()) for p in
                        (r.act_in, r.scale1[1], r.scale1[3], r.scale2[1], r.scale2[3], r.scale3[1], r.scale3[3],
                         r.fusion[1])],
            )
        return self._small

    def __call__(self, r_hat, jpeg=None):
        """x0 = jpeg + r_hat (fp32 NCHW [B,3,H,W]; or r_hat alone) -> (x0, refined fp32 NCHW [B,3,H,W])."""
        B, _, H, W = r_hat.shape
        dev = r_hat.device
        sm = self.small(dev)
        sl = sm["slopes"]
        if jpeg is None:
            x0, feat0 = ops.conv3ch(self.conv_in.layer, 3, 1, r_hat, act=ACT_PRELU, slope=sl[0])
        else:
            x0, feat0 = ops.conv3ch(self.conv_in.layer, 3, 1, jpeg, r_hat, sign=1, act=ACT_PRELU, slope=sl[0])
        feat, feat_h, feat_q, _ = ops.refine_se_scale_down(feat0, sm["fc1"], sm["fc2"])
        t, _, _ = self.scales[0][0](feat, act=ACT_PRELU, slope=sl[1])
        f1, _, _ = self.scales[0][1](t, act=ACT_PRELU, slope=sl[2])
        # the half / quarter resolution branches are kept with a one-pixel replicated border: the bilinear
        # up-samplings (tensor-core GEMMs in the statistics and fusion kernels) then need no clamping, and the 1x1
        # products of the fusion layer are computed on the padded grid (pointwise, so the border stays a replica)
        t, _, _ = self.scales[1][0](feat_h, act=ACT_PRELU, slope=sl[3])
        f2p, _, _ = self.scales[1][1](t, act=ACT_PRELU, slope=sl[4], out_pad=1)
        t, _, _ = self.scales[2][0](feat_q, act=ACT_PRELU, slope=sl[5])
        f3p, _, _ = self.scales[2][1](t, act=ACT_PRELU, slope=sl[6], out_pad=1)
        ops.replicate_border(f2p)
        ops.replicate_border(f3p)
        stats = ops.refine_stats3_tc(f1, f2p, f3p)
        att = ops.refine_spatial_att(stats, sm["w7"])
        t2, _, _ = self.fusion0[1](f2p)
        t3, _, _ = self.fusion0[2](f3p)
        h, _, _ = self.fusion0[0](f1, epi=EPI_PIXSCALE, pixscale=att, act=ACT_PRELU, slope=sl[7], up_t2=t2, up_t3=t3)
        _, _, refined = self.fusion2(h, out_bf16=False, out_f32="nchw")
        return x0, refined
