"""Tensor-level wrappers over the C-ABI ops.

Activations are NHWC: ``torch.bfloat16`` tensors of shape ``[B, H, W, C]`` on a
CUDA device.  Every wrapper launches on ``torch.cuda.current_stream()`` so the
calls compose with CUDA-graph capture; none of them synchronises.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_CLAMP01, ACT_NONE, ACT_PRELU, ACT_RELU, EPI_ADD, EPI_GATE, EPI_GDN,
                   EPI_IGDN, EPI_LINEAR, EPI_PIXSCALE, HYRES_CONV, HYRES_DECONV_K5S2)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk_nhwc(t, name, dtype=torch.bfloat16):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} "
                         f"cuda={t.is_cuda} contiguous={t.is_contiguous()}")


class ConvLayer:
    """One packed convolution layer (weights live in the library, bf16 K-major)."""

    def __init__(self, weight, bias=None, kind=HYRES_CONV, stride=1, pad=0, dil=1, cin0=None,
                 cin1=0, tap_mask=None):
        w = weight.detach().to("cpu", torch.float32).contiguous()
        b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
        if kind == HYRES_DECONV_K5S2:
            cin_total, cout, R, S = w.shape
        else:
            cout, cin_total, R, S = w.shape
        if cin0 is None:
            cin0 = cin_total
        self.kind, self.cin0, self.cin1, self.cout = kind, cin0, cin1, cout
        self.R, self.S, self.stride, self.pad, self.dil = R, S, stride, pad, dil
        self._w_shape = tuple(w.shape)
        mask = None
        if tap_mask is not None:
            mask = tap_mask.detach().to("cpu", torch.uint8).contiguous()
        h = C.c_void_p()
        lib = L.lib()
        L.check(lib.hyres_conv_create(C.byref(h), kind, cin0, cin1, cin_total, cout, R, S, stride,
                                      pad, dil, _ptr(w), _ptr(b), _ptr(mask)), "hyres_conv_create")
        self._h = h

    def update(self, weight, bias=None):
        w = weight.detach().to("cpu", torch.float32).contiguous()
        if tuple(w.shape) != self._w_shape:
            raise ValueError("ConvLayer.update: weight shape changed")
        b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
        L.check(L.lib().hyres_conv_update(self._h, _ptr(w), _ptr(b)), "hyres_conv_update")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                L.lib().hyres_conv_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def macs_per_pos(self):
        return L.lib().hyres_conv_macs_per_pos(self._h)

    def out_size(self, H, W):
        oh, ow = C.c_int(), C.c_int()
        L.check(L.lib().hyres_conv_out_size(self._h, H, W, C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    def __call__(self, x0, x1=None, epi=EPI_LINEAR, act=ACT_NONE, slope=0.0, aux0=None, aux1=None,
                 pixscale=None, out_bf16=True, out_sq=False, out_f32=None, mt=0):
        """Run the layer.

        out_bf16 / out_sq: True (allocate), False, or a preallocated NHWC tensor (its last
        dim may be wider than cout: the result lands in channels [0, cout) of that view).
        out_f32: None, "nhwc", "nchw", or a preallocated fp32 tensor in NHWC or (if
        ``out_f32_nchw`` attribute semantics are needed) pass a permuted view -- strides are
        taken from the tensor, dims interpreted as [B, OH, OW, C].
        Returns (bf16, sq, f32) with None for absent outputs.
        """
        _chk_nhwc(x0, "x0")
        B, H, W, c0 = x0.shape
        if c0 != self.cin0:
            raise ValueError(f"x0 has {c0} channels, layer expects {self.cin0}")
        if self.cin1:
            _chk_nhwc(x1, "x1")
            if tuple(x1.shape) != (B, H, W, self.cin1):
                raise ValueError("x1 shape mismatch")
        OH, OW = self.out_size(H, W)
        dev = x0.device
        io = L.ConvIO()
        io.x0, io.x1 = x0.data_ptr(), (x1.data_ptr() if self.cin1 else 0)
        io.B, io.H, io.W = B, H, W
        io.epi, io.act, io.slope = epi, act, float(slope)
        keep = [x0, x1, aux0, aux1, pixscale]
        if aux0 is not None:
            self._chk_aux(aux0, B, OH, OW, "aux0")
            io.aux0, io.ld_aux0 = aux0.data_ptr(), aux0.stride(2)
        if aux1 is not None:
            self._chk_aux(aux1, B, OH, OW, "aux1")
            io.aux1, io.ld_aux1 = aux1.data_ptr(), aux1.stride(2)
        if pixscale is not None:
            if pixscale.dtype != torch.float32 or pixscale.numel() != B * OH * OW or not pixscale.is_contiguous():
                raise ValueError("pixscale must be contiguous fp32 [B,OH,OW]")
            io.pixscale = pixscale.data_ptr()
        o16 = osq = o32 = None
        if out_bf16 is True:
            o16 = torch.empty((B, OH, OW, self.cout), dtype=torch.bfloat16, device=dev)
        elif out_bf16 is not False and out_bf16 is not None:
            o16 = out_bf16
        if o16 is not None:
            self._chk_aux(o16, B, OH, OW, "out_bf16")
            io.out_bf16, io.ld_out = o16.data_ptr(), o16.stride(2)
        if out_sq is True:
            osq = torch.empty((B, OH, OW, self.cout), dtype=torch.bfloat16, device=dev)
        elif out_sq is not False and out_sq is not None:
            osq = out_sq
        if osq is not None:
            self._chk_aux(osq, B, OH, OW, "out_sq")
            io.out_sq, io.ld_sq = osq.data_ptr(), osq.stride(2)
        if isinstance(out_f32, str):
            if out_f32 == "nhwc":
                o32 = torch.empty((B, OH, OW, self.cout), dtype=torch.float32, device=dev)
                view = o32
            elif out_f32 == "nchw":
                o32 = torch.empty((B, self.cout, OH, OW), dtype=torch.float32, device=dev)
                view = o32.permute(0, 2, 3, 1)
            else:
                raise ValueError("out_f32 must be 'nhwc' or 'nchw'")
        elif out_f32 is not None:
            o32 = out_f32
            view = o32
        if o32 is not None:
            if view.dtype != torch.float32 or tuple(view.shape[:3]) != (B, OH, OW) or view.shape[3] < self.cout:
                raise ValueError("out_f32 view must be fp32 [B,OH,OW,>=cout]")
            io.out_f32 = view.data_ptr()
            io.f32_sb, io.f32_sh, io.f32_sw, io.f32_sc = view.stride()
        io.mt_hint = mt
        keep += [o16, osq, o32]
        L.check(L.lib().hyres_conv_run(self._h, C.byref(io), _stream()), "hyres_conv_run")
        return o16, osq, o32

    @staticmethod
    def _chk_aux(t, B, OH, OW, name):
        if t.dtype != torch.bfloat16 or not t.is_cuda:
            raise ValueError(f"{name}: expected CUDA bf16 tensor")
        if tuple(t.shape[:3]) != (B, OH, OW) or t.stride(3) != 1:
            raise ValueError(f"{name}: expected [B={B},OH={OH},OW={OW},C] channel-contiguous, got {tuple(t.shape)}")
        if t.stride(1) != OW * t.stride(2) or t.stride(0) != OH * OW * t.stride(2):
            raise ValueError(f"{name}: pixel stride must be uniform (a channel slice of a dense NHWC tensor)")
