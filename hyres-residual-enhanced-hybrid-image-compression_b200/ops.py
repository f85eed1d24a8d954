"""Tensor-level wrappers over the C-ABI ops.

Activations are NHWC: ``torch.bfloat16`` tensors of shape ``[B, H, W, C]`` on a
CUDA device.  Every wrapper launches on ``torch.cuda.current_stream()`` so the
calls compose with CUDA-graph capture; none of them synchronises.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_CLAMP01, ACT_NONE, ACT_PRELU, ACT_RELU, EPI_ADD, EPI_GATE, EPI_GDN,
                   EPI_IGDN, EPI_LINEAR, EPI_PIXSCALE, HYRES_CONV, HYRES_DECONV_K5S2, SPLIT_ADD, SPLIT_COPY,
                   SPLIT_F16, SPLIT_GATE, SPLIT_GDN, SPLIT_IGDN, SPLIT_ROUND_CHAN, SPLIT_SQUARE)


def sm_count():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk_nhwc(t, name, dtype=torch.bfloat16):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} "
                         f"cuda={t.is_cuda} contiguous={t.is_contiguous()}")


# Deployment export (export.py): while PACKED_COLLECT is a dict every new ConvLayer stores its packed device operands
# in it, keyed by a digest of (geometry, weights, bias); while PACKED_CACHE is a dict a new ConvLayer whose digest is
# found there is created empty and filled from the stored operands instead of being packed again.
PACKED_COLLECT = None
PACKED_CACHE = None
PACKED_STATS = {"imported": 0, "packed": 0}  # layers filled from exported operands / packed from fp32 weights


_SYNC_TLS = __import__("threading").local()


def stream_wait_blocking(device=None):
    """Wait for the current stream's work without spinning: a blocking-sync event lets the calling thread sleep
    (``Stream.synchronize`` busy-waits, and a codec pipeline has several worker threads waiting on the GPU while the
    host cores are needed by the range coder)."""
    ev = getattr(_SYNC_TLS, "ev", None)
    if ev is None:
        ev = _SYNC_TLS.ev = {}
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    e = ev.get(dev)
    if e is None:
        e = ev[dev] = torch.cuda.Event(blocking=True)
    e.record(torch.cuda.current_stream(dev))
    e.synchronize()


def read_small(t):
    """A small CUDA tensor -> list of Python scalars, through pinned memory and a blocking event wait (``.item()`` /
    ``.tolist()`` spin on a host core until the stream drains)."""
    nbytes = t.numel() * t.element_size()
    host = _pinned_bytes("small", nbytes)[:nbytes].view(t.dtype).view(t.shape)
    host.copy_(t, non_blocking=True)
    stream_wait_blocking(t.device)
    return host.tolist()


def split_parts(code):
    """Number of 16-bit parts per fp32 value of an nsplit code (1..3 bf16 parts, or 2 | SPLIT_F16: two half parts)."""
    return int(code) & 15


class ConvLayer:
    """One packed convolution layer (weights live in the library, bf16 K-major)."""

    _prof = None  # list of (start, end) CUDA events while profile_begin() is active
    _prof_info = []

    @classmethod
    def profile_begin(cls):
        """Bracket every conv launch with CUDA events on the launch stream (bench.py roofline)."""
        cls._prof = []
        cls._prof_info = []

    @classmethod
    def profile_end(cls):
        """-> (summed device milliseconds of the conv launches, number of launches)."""
        ev, cls._prof = cls._prof or [], None
        torch.cuda.synchronize()
        cls.last_profile = [dict(info, ms=a.elapsed_time(b)) for (a, b), info in zip(ev, cls._prof_info)]
        return sum(a.elapsed_time(b) for a, b in ev), len(ev)

    def __init__(self, weight, bias=None, kind=HYRES_CONV, stride=1, pad=0, dil=1, cin0=None,
                 cin1=0, tap_mask=None, nsplit=1):
        """nsplit > 1: split-precision layer (hyres_conv_create_split): inputs are the bf16 parts produced by
        ``split_f32`` ([B,H,W,nsplit*cin]), the result is fp32 (``out_f32``) with fp32-equivalent accuracy.
        nsplit = 2 | SPLIT_F16: two IEEE half parts (3 tensor-core products per MAC instead of 6)."""
        if isinstance(weight, (tuple, list, torch.Size)):
            # shape only: an empty layer whose operands arrive later (update_device: training; import: deployment)
            w, wshape = None, tuple(int(v) for v in weight)
        else:
            w = weight.detach().to("cpu", torch.float32).contiguous()
            wshape = tuple(w.shape)
        b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
        if kind == HYRES_DECONV_K5S2:
            cin_total, cout, R, S = wshape
        else:
            cout, cin_total, R, S = wshape
        if cin0 is None:
            cin0 = cin_total
        self.kind, self.cin0, self.cin1, self.cout, self.nsplit = kind, cin0, cin1, cout, split_parts(nsplit)
        self.split_code = int(nsplit)
        self.R, self.S, self.stride, self.pad, self.dil = R, S, stride, pad, dil
        self._w_shape = wshape
        mask = None
        if tap_mask is not None:
            mask = tap_mask.detach().to("cpu", torch.uint8).contiguous()
        # algorithmic MACs per OUTPUT position (live taps only; a transposed k5 s2 layer touches 25/4 taps on average)
        taps = int(mask.sum()) if mask is not None else R * S
        self.alg_macs_per_out = (cin0 + cin1) * cout * taps / (4.0 if kind == HYRES_DECONV_K5S2 else 1.0)
        h = C.c_void_p()
        lib = L.lib()
        key = None
        if w is not None and (PACKED_COLLECT is not None or PACKED_CACHE is not None):
            import hashlib
            d = hashlib.sha1(repr((kind, cin0, cin1, cin_total, cout, R, S, stride, pad, dil, nsplit)).encode())
            d.update(w.numpy().tobytes())
            d.update(b.numpy().tobytes() if b is not None else b"-")
            d.update(mask.numpy().tobytes() if mask is not None else b"-")
            key = d.hexdigest()
        cached = PACKED_CACHE.get(key) if (PACKED_CACHE is not None and key is not None) else None
        L.check(lib.hyres_conv_create_split(C.byref(h), kind, cin0, cin1, cin_total, cout, R, S, stride,
                                            pad, dil, _ptr(w) if cached is None else C.c_void_p(0), _ptr(b), _ptr(mask),
                                            nsplit), "hyres_conv_create")
        self._h = h
        if cached is not None:
            pw, pt, pb = cached
            sizes = [lib.hyres_conv_packed_elems(h, i) for i in range(3)]
            if [pw.numel(), pt.numel(), pb.numel()] != sizes:
                raise ValueError("ConvLayer: exported operands do not match this layer's packing")
            L.check(lib.hyres_conv_import_packed(h, _ptr(pw), _ptr(pt) if pt.numel() else C.c_void_p(0), _ptr(pb)),
                    "hyres_conv_import_packed")
            PACKED_STATS["imported"] += 1
        else:
            PACKED_STATS["packed"] += 1
        if cached is not None:
            pass
        elif PACKED_COLLECT is not None and key is not None:
            pw = torch.empty(lib.hyres_conv_packed_elems(h, 0), dtype=torch.bfloat16)
            pt = torch.empty(lib.hyres_conv_packed_elems(h, 1), dtype=torch.bfloat16)
            pb = torch.empty(lib.hyres_conv_packed_elems(h, 2), dtype=torch.float32)
            L.check(lib.hyres_conv_export_packed(h, _ptr(pw), _ptr(pt) if pt.numel() else C.c_void_p(0), _ptr(pb)),
                    "hyres_conv_export_packed")
            PACKED_COLLECT[key] = (pw, pt, pb)

    def update(self, weight, bias=None):
        w = weight.detach().to("cpu", torch.float32).contiguous()
        if tuple(w.shape) != self._w_shape:
            raise ValueError("ConvLayer.update: weight shape changed")
        b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
        L.check(L.lib().hyres_conv_update(self._h, _ptr(w), _ptr(b)), "hyres_conv_update")

    def update_device(self, weight, bias=None):
        """Re-pack from fp32 CUDA tensors with a kernel on the current stream (no host round trip)."""
        w = weight.detach()
        if tuple(w.shape) != self._w_shape or not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous():
            raise ValueError("ConvLayer.update_device: expected a contiguous fp32 CUDA weight of the original shape")
        b = None
        if bias is not None:
            b = bias.detach()
            if not b.is_cuda or b.dtype != torch.float32 or b.numel() != self.cout or not b.is_contiguous():
                raise ValueError("ConvLayer.update_device: bias must be a contiguous fp32 CUDA vector of cout entries")
        L.check(L.lib().hyres_conv_update_device(self._h, _ptr(w), _ptr(b), _stream()), "hyres_conv_update_device")

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
                 pixscale=None, out_bf16=True, out_sq=False, out_f32=None, mt=0, x0_square=False, out_pad=0,
                 up_t2=None, up_t3=None, cta_limit=0, split_mode=SPLIT_COPY, aux0_f32=None, aux1_f32=None,
                 out_split=None, split_square=False, out_code=None):
        """Run the layer.

        out_bf16 / out_sq: True (allocate), False, or a preallocated NHWC tensor (its last
        dim may be wider than cout: the result lands in channels [0, cout) of that view).
        out_f32: None, "nhwc", "nchw", or a preallocated fp32 tensor in NHWC or (if
        ``out_f32_nchw`` attribute semantics are needed) pass a permuted view -- strides are
        taken from the tensor, dims interpreted as [B, OH, OW, C].
        out_pad: > 0 allocates (or expects) the bf16 output as [B, OH+2p, OW+2p, C] and stores into its interior.
        up_t2 / up_t3: padded bf16 [B, OH/2+2, OW/2+2, 64] / [B, OH/4+2, OW/4+2, 64]; their bilinear x2 / x4
        up-samplings are added to the accumulator before the epilogue (MultiScaleRefine fusion).
        Split-precision layers (nsplit > 1): ``split_mode`` / ``aux0_f32`` / ``aux1_f32`` fuse the fp32
        element-wise stage into the epilogue and ``out_split`` (True or a tensor [B,OH,OW,nsplit*cout]) receives the
        bf16 parts of the result (of its square with ``split_square``); the parts tensor is returned in the ``sq``
        slot.  ``out_code``: nsplit code of those parts (default: the layer's own format).
        Returns (bf16, sq, f32) with None for absent outputs.
        """
        _chk_nhwc(x0, "x0")
        B, H, W, c0 = x0.shape
        if c0 != self.nsplit * self.cin0:
            raise ValueError(f"x0 has {c0} channels, layer expects {self.nsplit} x {self.cin0}")
        if self.cin1:
            _chk_nhwc(x1, "x1")
            if tuple(x1.shape) != (B, H, W, self.nsplit * self.cin1):
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
            o16 = torch.empty((B, OH + 2 * out_pad, OW + 2 * out_pad, self.cout), dtype=torch.bfloat16, device=dev)
        elif out_bf16 is not False and out_bf16 is not None:
            o16 = out_bf16
        if o16 is not None:
            self._chk_aux(o16, B, OH + 2 * out_pad, OW + 2 * out_pad, "out_bf16")
            io.out_bf16, io.ld_out = o16.data_ptr(), o16.stride(2)
            io.out_pad = int(out_pad)
        if (up_t2 is None) != (up_t3 is None):
            raise ValueError("up_t2 and up_t3 go together")
        if up_t2 is not None:
            for t, d, nm in ((up_t2, 2, "up_t2"), (up_t3, 4, "up_t3")):
                if (t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous()
                        or tuple(t.shape) != (B, OH // d + 2, OW // d + 2, 64)):
                    raise ValueError(f"{nm}: expected contiguous CUDA bf16 [B,{OH // d + 2},{OW // d + 2},64]")
            io.up_t2, io.up_t3 = up_t2.data_ptr(), up_t3.data_ptr()
            keep += [up_t2, up_t3]
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
        if out_split is not None and out_split is not False:
            if self.nsplit == 1:
                raise ValueError("out_split needs a split-precision layer")
            code = self.split_code if out_code is None else int(out_code)
            nparts = split_parts(code)
            if out_split is True:
                out_split = torch.empty((B, OH, OW, nparts * self.cout), dtype=torch.bfloat16, device=dev)
            if (out_split.dtype != torch.bfloat16 or not out_split.is_contiguous()
                    or tuple(out_split.shape) != (B, OH, OW, nparts * self.cout)):
                raise ValueError("out_split: expected contiguous bf16 [B,OH,OW,parts*cout]")
            io.out_split, io.out_nsplit, io.split_square = out_split.data_ptr(), code, 1 if split_square else 0
            osq = out_split
        io.split_mode = int(split_mode)
        for t, nm in ((aux0_f32, "aux0_f32"), (aux1_f32, "aux1_f32")):
            if t is not None:
                if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != (B, OH, OW, self.cout):
                    raise ValueError(f"{nm}: expected contiguous fp32 [B,OH,OW,cout]")
        io.aux0_f32 = aux0_f32.data_ptr() if aux0_f32 is not None else 0
        io.aux1_f32 = aux1_f32.data_ptr() if aux1_f32 is not None else 0
        keep += [aux0_f32, aux1_f32]
        io.mt_hint = mt
        io.cta_limit = int(cta_limit)
        io.x0_square = 1 if x0_square else 0
        keep += [o16, osq, o32]
        if ConvLayer._prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.lib().hyres_conv_run(self._h, C.byref(io), _stream()), "hyres_conv_run")
            e1.record()
            ConvLayer._prof.append((e0, e1))
            ConvLayer._prof_info.append(dict(kind=self.kind, cin=self.cin0 + self.cin1, cout=self.cout, k=self.R,
                                             stride=self.stride, dil=self.dil, B=B, H=H, W=W, OH=OH, OW=OW,
                                             epi=epi, f32=o32 is not None, sq=osq is not None, nsplit=self.nsplit,
                                             products=(1, 1, 3, 6)[self.nsplit], split_mode=int(split_mode),
                                             split_out=osq is not None and self.nsplit > 1,
                                             alg_macs=self.alg_macs_per_out * B * OH * OW))
        else:
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


def _wgrad_plan(tc):
    """The convolution whose weight gradient equals this training node's, and whether input / output-gradient swap
    roles (transposed convolution: the weight gradient of its stride-2 data-gradient convolution)."""
    if tc.kind == HYRES_DECONV_K5S2:
        return tc.dgrad_layer(), True
    return tc.wgrad_layer(), False


def wgrad_supported(tc):
    """True when the tcgen05 weight-gradient kernel (csrc/wgrad.cu) covers this training convolution."""
    layer, _ = _wgrad_plan(tc)
    return bool(L.lib().hyres_wgrad_supported(layer._h))


def conv_wgrad(tc, x, g16, want_bias=True):
    """x: the node's input, g16: gradient of its output (both bf16 NHWC) -> (dW fp32 in the parameter's layout,
    db fp32 or None)."""
    layer, swap = _wgrad_plan(tc)
    inp, gout = (g16, x) if swap else (x, g16)
    _chk_nhwc(inp, "wgrad input"), _chk_nhwc(gout, "wgrad output gradient")
    B, H, W, _ = inp.shape
    lib = L.lib()
    nbytes = lib.hyres_wgrad_workspace_bytes(layer._h, B, H, W)
    if nbytes <= 0:
        raise ValueError("conv_wgrad: unsupported layer")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=inp.device)
    dw = torch.empty(tc.w_shape, dtype=torch.float32, device=inp.device)
    L.check(lib.hyres_wgrad_run(layer._h, _ptr(inp), _ptr(gout), B, H, W, _ptr(dw), _ptr(ws), _stream()),
            "hyres_wgrad_run")
    db = colsum_bf16(g16) if want_bias else None
    return dw, db


def colsum_bf16(g):
    """bf16 [..., C] -> fp32 [C]: sum over all leading dimensions (a layer's bias gradient)."""
    _chk_nhwc(g, "g")
    Cc = g.shape[-1]
    rows = g.numel() // Cc
    lib = L.lib()
    ws = torch.empty(lib.hyres_colsum_workspace_bytes(rows, Cc), dtype=torch.uint8, device=g.device)
    out = torch.empty(Cc, dtype=torch.float32, device=g.device)
    L.check(lib.hyres_colsum_bf16(_ptr(g), rows, Cc, _ptr(out), _ptr(ws), _stream()), "hyres_colsum_bf16")
    return out


def ru_supported(c1, c2, c3):
    """True when the three layers form the C=128 bottleneck the fused kernel implements."""
    return bool(L.lib().hyres_ru_supported(c1._h, c2._h, c3._h))


def ru_fused(x, c1, c2, c3, final_relu, out=None):
    """out = [ReLU](x + c3(ReLU(c2(ReLU(c1(x)))))) in one persistent kernel (csrc/ru_fused.cu)."""
    _chk_nhwc(x, "x")
    B, H, W, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    io = L.RuIO()
    io.x, io.ld_x, io.out, io.ld_out = x.data_ptr(), x.stride(2), out.data_ptr(), out.stride(2)
    io.B, io.H, io.W, io.final_relu = B, H, W, 1 if final_relu else 0
    if ConvLayer._prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().hyres_ru_run(c1._h, c2._h, c3._h, C.byref(io), _stream()), "hyres_ru_run")
        e1.record()
        ConvLayer._prof.append((e0, e1))
        ConvLayer._prof_info.append(dict(kind="ru", cin=Cc, cout=Cc, k=3, stride=1, dil=1, B=B, H=H, W=W, OH=H, OW=W,
                                         epi=EPI_ADD, f32=False, sq=False))
    else:
        L.check(L.lib().hyres_ru_run(c1._h, c2._h, c3._h, C.byref(io), _stream()), "hyres_ru_run")
    return out


# --------------------------------------------------------------------------------------
# memory-bound kernels
# --------------------------------------------------------------------------------------
def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA fp32 tensor")
    return t


def conv3ch(layer, ksize, stride, a, b=None, sign=1, want_sum=True, act=ACT_NONE, slope=0.0, out=None):
    """First-layer convolution of a 3-channel fp32 NCHW image, fused with ``src = a + sign*b``
    (csrc/conv_c3.cu).  ``layer``: the 1x1 ConvLayer over the im2col ordering.  Returns
    (src fp32 NCHW -- ``a`` itself when b is None --, activation bf16 NHWC [B,OH,OW,cout])."""
    _f32c(a, "a")
    B, Cc, H, W = a.shape
    if Cc != 3:
        raise ValueError("expected 3 channels")
    src = a
    if b is not None:
        _f32c(b, "b")
        if b.shape != a.shape:
            raise ValueError("a / b shape mismatch")
        src = torch.empty_like(a) if want_sum else None
    OH, OW = (H // 2, W // 2) if stride == 2 else (H, W)
    if out is None:
        out = torch.empty((B, OH, OW, layer.cout), dtype=torch.bfloat16, device=a.device)
    ConvLayer._chk_aux(out, B, OH, OW, "out")

    def run():
        L.check(L.lib().hyres_conv3ch_run(layer._h, ksize, stride, _ptr(a), _ptr(b), sign,
                                          _ptr(src if b is not None else None), _ptr(out), out.stride(2), B, H, W,
                                          act, float(slope), _stream()), "hyres_conv3ch_run")
    if ConvLayer._prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        ConvLayer._prof.append((e0, e1))
        ConvLayer._prof_info.append(dict(kind="c3", cin=3, cout=layer.cout, k=ksize, stride=stride, dil=1, B=B, H=H, W=W,
                                         OH=OH, OW=OW, epi=EPI_LINEAR, f32=False, sq=False))
    else:
        run()
    return src, out


def split_f32(x, mode=SPLIT_COPY, aux0=None, aux1=None, chan=None, relu=False, nsplit=3, want_f32=False,
              want_split=True):
    """Element-wise stage of the split-precision trunk (csrc/precise.cu).  x, aux0, aux1: fp32 NHWC [..., C];
    chan: fp32 [C].  -> (v fp32 or None, bf16 parts [..., nsplit*C] or None); COPY without ReLU returns x itself
    as the fp32 result."""
    _f32c(x, "x")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    for t, nm in ((aux0, "aux0"), (aux1, "aux1")):
        if t is not None:
            _f32c(t, nm)
            if t.shape != x.shape:
                raise ValueError(f"split_f32: {nm} shape mismatch")
    if chan is not None:
        _f32c(chan, "chan")
        if chan.numel() != Cc:
            raise ValueError("split_f32: chan must have C entries")
    trivial = mode == SPLIT_COPY and not relu
    o32 = torch.empty_like(x) if (want_f32 and not trivial) else None
    osp = torch.empty(x.shape[:-1] + (split_parts(nsplit) * Cc,), dtype=torch.bfloat16, device=x.device) if want_split else None
    if o32 is not None or osp is not None:
        L.check(L.lib().hyres_split_f32(_ptr(x), rows, Cc, mode, _ptr(aux0), _ptr(aux1), _ptr(chan), 1 if relu else 0,
                                        _ptr(o32), _ptr(osp), nsplit, _stream()), "hyres_split_f32")
    return (x if (want_f32 and trivial) else o32), osp


def residual_im2col5s2_split(x, jpeg=None, nsplit=3):
    """x, jpeg: fp32 NCHW [B,3,H,W] -> (residual fp32 NCHW or x, bf16 parts [B,H/2,W/2,nsplit*128])."""
    _f32c(x, "x")
    B, Cc, H, W = x.shape
    if Cc != 3:
        raise ValueError("expected 3 channels")
    res = None
    if jpeg is not None:
        _f32c(jpeg, "jpeg")
        res = torch.empty_like(x)
    a = torch.empty((B, H // 2, W // 2, split_parts(nsplit) * 128), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().hyres_residual_im2col5s2_split(_ptr(x), _ptr(jpeg), _ptr(res), _ptr(a), nsplit, B, H, W,
                                                   _stream()), "hyres_residual_im2col5s2_split")
    return (res if res is not None else x), a


def symbols_to_nhwc_f32(symbols, chan=None):
    """int32 [B,C,h,w] (+ per-channel fp32 offset) -> fp32 NHWC [B,h,w,C]."""
    B, Cc, h, w = symbols.shape
    out = torch.empty((B, h, w, Cc), dtype=torch.float32, device=symbols.device)
    L.check(L.lib().hyres_symbols_to_nhwc_f32(_ptr(symbols), _ptr(chan), _ptr(out), B, h, w, Cc, _stream()),
            "hyres_symbols_to_nhwc_f32")
    return out


def final_clamp(x0, refined):
    _f32c(x0, "x0"), _f32c(refined, "refined")
    out = torch.empty_like(x0)
    L.check(L.lib().hyres_final_clamp(_ptr(x0), _ptr(refined), _ptr(out), x0.numel(), _stream()), "hyres_final_clamp")
    return out


def gc_quant_pass(y, params, pass_id, noise=False, seed=0, want_f32=True, want_bf16=True):
    """y fp32 NHWC [B,h,w,M]; params fp32 NHWC [B,h,w,2M] -> (yq fp32 NHWC, yq bf16 NHWC)."""
    _f32c(y, "y")
    B, h, w, M = y.shape
    if params is not None:
        _f32c(params, "params")
    o32 = torch.empty_like(y) if want_f32 else None
    o16 = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    L.check(L.lib().hyres_gc_quant_pass(_ptr(y), _ptr(params), pass_id, 1 if noise else 0, seed, _ptr(o32), _ptr(o16),
                                        B, h, w, M, _stream()), "hyres_gc_quant_pass")
    return o32, o16


def gc_merge_likelihood(y, params_a, params_na, yq_a, yq_na, noise=False, seed=0, want_lik=True, sum_log2=None):
    """-> (y_hat bf16 NHWC, lik fp32 NCHW).  sum_log2: optional 1-element CUDA double accumulator."""
    B, h, w, M = y.shape
    y_hat = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device) if yq_a is not None else None
    lik = torch.empty((B, M, h, w), dtype=torch.float32, device=y.device) if want_lik else None
    L.check(L.lib().hyres_gc_merge_likelihood(_ptr(y), _ptr(params_a), _ptr(params_na), _ptr(yq_a), _ptr(yq_na),
                                              1 if noise else 0, seed, _ptr(y_hat), _ptr(lik), _ptr(sum_log2), B, h,
                                              w, M, _stream()), "hyres_gc_merge_likelihood")
    return y_hat, lik


def _chk_rows(rows, scale_table):
    if rows.dtype != torch.int32 or not rows.is_cuda or not rows.is_contiguous() or tuple(rows.shape) != (3, scale_table.numel()):
        raise ValueError("rows: expected the contiguous int32 CUDA tensor [3, n_scales] of EntropyModel.coder_rows()")


def gc_symbols(y, params, pass_id, scale_table, scale_bound=0.11, want_f32=True, want_bf16=True, rows=None):
    """-> (symbols int32 [B,M,h,w], indexes int32 [B,M,h,w], yq fp32 NHWC, yq bf16 NHWC); with ``rows`` (the host
    coder's table layout, ``EntropyModel.coder_rows``) a fifth result: the coder slots int32 [B,M,h,w]."""
    _f32c(y, "y"), _f32c(params, "params"), _f32c(scale_table, "scale_table")
    B, h, w, M = y.shape
    sym = torch.empty((B, M, h, w), dtype=torch.int32, device=y.device)
    idx = torch.empty((B, M, h, w), dtype=torch.int32, device=y.device)
    o32 = torch.empty_like(y) if want_f32 else None
    o16 = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    slots = None
    if rows is not None:
        _chk_rows(rows, scale_table)
        slots = torch.empty((B, M, h, w), dtype=torch.int32, device=y.device)
    L.check(L.lib().hyres_gc_symbols(_ptr(y), _ptr(params), pass_id, _ptr(scale_table), scale_table.numel(),
                                     float(scale_bound), _ptr(sym), _ptr(idx), _ptr(o32), _ptr(o16), B, h, w, M,
                                     _ptr(rows), _ptr(slots), _stream()), "hyres_gc_symbols")
    if rows is not None:
        return sym, idx, o32, o16, slots
    return sym, idx, o32, o16


def gc_codes(params, pass_id, scale_table, M, rows, scale_bound=0.11):
    """Decoder front-end of checkerboard pass ``pass_id``: -> decoder codes int32 [B,M,h,w] (CDF row index, or
    bit 30 | packed entry at the structurally zero positions, whose symbol round(-mean) needs no decoding)."""
    _f32c(params, "params")
    _chk_rows(rows, scale_table)
    B, h, w, _ = params.shape
    codes = torch.empty((B, M, h, w), dtype=torch.int32, device=params.device)
    L.check(L.lib().hyres_gc_codes(_ptr(params), int(pass_id), _ptr(scale_table), scale_table.numel(), float(scale_bound),
                                   _ptr(rows), C.c_void_p(0), _ptr(codes), B, h, w, M, _stream()), "hyres_gc_codes")
    return codes


def gc_indexes(params, scale_table, M, scale_bound=0.11):
    _f32c(params, "params")
    B, h, w, _ = params.shape
    idx = torch.empty((B, M, h, w), dtype=torch.int32, device=params.device)
    L.check(L.lib().hyres_gc_indexes(_ptr(params), _ptr(scale_table), scale_table.numel(), float(scale_bound),
                                     _ptr(idx), B, h, w, M, _stream()), "hyres_gc_indexes")
    return idx


def gc_dequant(symbols, params, want_f32=True, want_bf16=True, pass_id=-1):
    """symbols int32 [B,M,h,w] (+ means from params NHWC) -> (yq fp32 NHWC, yq bf16 NHWC).  ``pass_id`` 0 / 1: the
    symbols at that pass's structurally zero positions are recomputed as round(-mean), not read (``gc_codes``)."""
    B, M, h, w = symbols.shape
    dev = symbols.device
    o32 = torch.empty((B, h, w, M), dtype=torch.float32, device=dev) if want_f32 else None
    o16 = torch.empty((B, h, w, M), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    L.check(L.lib().hyres_gc_dequant(_ptr(symbols), _ptr(params), _ptr(o32), _ptr(o16), B, h, w, M, int(pass_id),
                                     _stream()), "hyres_gc_dequant")
    return o32, o16


def add_to_bf16(a, b):
    out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    L.check(L.lib().hyres_add_to_bf16(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "hyres_add_to_bf16")
    return out


def eb_forward(z, eb_params, medians, lik_noise=False, out_noise=False, seed=0, lik_bound=1e-9, want_zhat_nchw=False,
               want_lik=True, want_symbols=False, sum_log2=None):
    """z fp32 NHWC [B,h,w,C] -> dict(zhat_bf16 NHWC, zhat_nchw, lik NCHW, symbols [B,C,h,w])"""
    _f32c(z, "z")
    B, h, w, Cc = z.shape
    dev = z.device
    zh16 = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
    zh32 = torch.empty((B, Cc, h, w), dtype=torch.float32, device=dev) if want_zhat_nchw else None
    lik = torch.empty((B, Cc, h, w), dtype=torch.float32, device=dev) if want_lik else None
    sym = torch.empty((B, Cc, h, w), dtype=torch.int32, device=dev) if want_symbols else None
    mode = (1 if lik_noise else 0) | (2 if out_noise else 0)
    L.check(L.lib().hyres_eb_forward(_ptr(z), _ptr(eb_params), _ptr(medians), mode, seed, float(lik_bound), _ptr(zh16),
                                     _ptr(zh32), _ptr(lik), _ptr(sym), _ptr(sum_log2), B, h, w, Cc, _stream()),
            "hyres_eb_forward")
    return dict(zhat_bf16=zh16, zhat_nchw=zh32, lik=lik, symbols=sym)


def eb_dequant(symbols, medians):
    B, Cc, h, w = symbols.shape
    out = torch.empty((B, h, w, Cc), dtype=torch.bfloat16, device=symbols.device)
    L.check(L.lib().hyres_eb_dequant(_ptr(symbols), _ptr(medians), _ptr(out), B, h, w, Cc, _stream()),
            "hyres_eb_dequant")
    return out


def refine_se_scale_down(feat, fc1, fc2):
    """feat bf16 NHWC [B,H,W,64] -> (feat*se bf16, half-res, quarter-res, pooled fp32 [B,64])"""
    _chk_nhwc(feat, "feat")
    B, H, W, Cc = feat.shape
    dev = feat.device
    scratch = torch.empty((B * 64 * Cc,), dtype=torch.float32, device=dev)
    pooled = torch.empty((B, Cc), dtype=torch.float32, device=dev)
    lib = L.lib()
    L.check(lib.hyres_refine_se_pool(_ptr(feat), _ptr(scratch), _ptr(pooled), B, H, W, Cc, _stream()),
            "hyres_refine_se_pool")
    fs = torch.empty_like(feat)
    fh = torch.empty((B, H // 2, W // 2, Cc), dtype=torch.bfloat16, device=dev)
    fq = torch.empty((B, H // 4, W // 4, Cc), dtype=torch.bfloat16, device=dev)
    L.check(lib.hyres_refine_se_scale_down(_ptr(feat), _ptr(pooled), _ptr(fc1), _ptr(fc2), Cc, fc1.shape[0], _ptr(fs),
                                           _ptr(fh), _ptr(fq), B, H, W, _stream()), "hyres_refine_se_scale_down")
    return fs, fh, fq, pooled


def refine_stats3_tc(f1, s2p, s3p):
    """Channel mean / max over the virtual concat [f1 | up2(f2) | up4(f3)] (bf16 NHWC, 64 channels each) -> fp32
    [B,H,W,2] on the tensor cores, without materialising the up-sampled channels: s2p / s3p are the half / quarter
    resolution tensors padded by one replicated pixel ([B,H/2+2,W/2+2,64] / [B,H/4+2,W/4+2,64])."""
    _chk_nhwc(f1, "f1"), _chk_nhwc(s2p, "s2p"), _chk_nhwc(s3p, "s3p")
    B, H, W, Cc = f1.shape
    if Cc != 64 or tuple(s2p.shape) != (B, H // 2 + 2, W // 2 + 2, 64) or tuple(s3p.shape) != (B, H // 4 + 2, W // 4 + 2, 64):
        raise ValueError("refine_stats3_tc: expected 64-channel f1 and padded half / quarter resolution tensors")
    stats = torch.empty((B, H, W, 2), dtype=torch.float32, device=f1.device)
    L.check(L.lib().hyres_refine_stats3_tc(_ptr(f1), _ptr(s2p), _ptr(s3p), _ptr(stats), B, H, W, _stream()),
            "hyres_refine_stats3_tc")
    return stats


def replicate_border(t):
    """In place: the one-pixel border of a padded bf16 NHWC tensor [B,Hp,Wp,C] <- nearest interior pixel."""
    _chk_nhwc(t, "t")
    B, Hp, Wp, Cc = t.shape
    L.check(L.lib().hyres_replicate_border(_ptr(t), B, Hp, Wp, Cc, _stream()), "hyres_replicate_border")
    return t


def refine_spatial_att(stats, w7):
    B, H, W, _ = stats.shape
    att = torch.empty((B, H, W), dtype=torch.float32, device=stats.device)
    L.check(L.lib().hyres_refine_spatial_att(_ptr(stats), _ptr(w7), _ptr(att), B, H, W, _stream()),
            "hyres_refine_spatial_att")
    return att


def nchw_f32_to_nhwc_bf16(x):
    _f32c(x, "x")
    B, Cc, H, W = x.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().hyres_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), B, Cc, H, W, _stream()))
    return out


def nhwc_to_nchw_f32(x):
    _f32c(x, "x")
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    L.check(L.lib().hyres_nhwc_to_nchw_f32(_ptr(x), _ptr(out), B, Cc, H, W, _stream()))
    return out


def nhwc_bf16_to_nchw_f32(x):
    _chk_nhwc(x, "x")
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    L.check(L.lib().hyres_nhwc_bf16_to_nchw_f32(_ptr(x), _ptr(out), B, Cc, H, W, _stream()))
    return out


def reduce_sqdiff(a, b, out):
    L.check(L.lib().hyres_reduce_sqdiff(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()), "hyres_reduce_sqdiff")


def rd_loss_finalize(sum_y, sum_z, sum_se, jpeg_bpp, num_pixels, num_elems, lmbda):
    """-> fp32 [6] = y_bpp, z_bpp, residual_bpp, bpp, mse * 255^2, loss (src/losses/rd_loss.py:23-44), one launch."""
    out = torch.empty(6, dtype=torch.float32, device=sum_y.device)
    L.check(L.lib().hyres_rd_loss_finalize(_ptr(sum_y), _ptr(sum_z), _ptr(sum_se), _ptr(jpeg_bpp), float(num_pixels),
                                           float(num_elems), float(lmbda), _ptr(out), _stream()), "hyres_rd_loss_finalize")
    return out


def reduce_log2(x, out):
    L.check(L.lib().hyres_reduce_log2(_ptr(x), x.numel(), _ptr(out), _stream()), "hyres_reduce_log2")


# ---------------------------------------------------------------------------------------------
# JPEG stage on the device (models/utils/turbo_jpeg_compression.py:17-40,62-77)
# ---------------------------------------------------------------------------------------------
def _jpeg_buffers(dev, B, H, W, want_scan):
    """Scratch of one call, from torch's caching allocator (under CUDA-graph capture: from the graph's pool, so a
    replayed graph never sees memory that a later call reuses)."""
    n = L.lib().hyres_jpeg_workspace_bytes(B, H, W)
    if n <= 0:
        raise ValueError(f"jpeg_forward: unsupported size {H}x{W} (H must be a multiple of 8, W of 16)")
    ws = {"ws": torch.empty(n, dtype=torch.uint8, device=dev)}
    if want_scan:
        ws["wpi"] = L.lib().hyres_jpeg_scan_words(H, W)
        ws["words"] = torch.empty(B * ws["wpi"], dtype=torch.int32, device=dev)
    return ws


def jpeg_forward(x, quality, want_decoded=True, want_sizes=True, want_scan=False):
    """x fp32 NCHW [B,3,H,W] in [0,1] on the device -> dict(decoded fp32 NCHW, sizes int64 [B] (file bytes),
    words int32 [B, words_per_image] + nbits int64 [B] (the entropy-coded scans, big-endian bit strings)).
    Bit-exact with libjpeg-turbo's 4:2:2 baseline round trip (tests/test_gpu_jpeg.py)."""
    _f32c(x, "x")
    B, C3, H, W = x.shape
    if C3 != 3:
        raise ValueError("jpeg_forward: expected [B,3,H,W]")
    need_scan = want_sizes or want_scan
    ws = _jpeg_buffers(x.device, B, H, W, need_scan)
    dec = torch.empty_like(x) if want_decoded else None
    sizes = torch.empty(B, dtype=torch.int64, device=x.device) if want_sizes else None
    nbits = torch.empty(B, dtype=torch.int64, device=x.device) if need_scan else None
    L.check(L.lib().hyres_jpeg_forward(_ptr(x), B, H, W, int(quality), _ptr(ws["ws"]), _ptr(dec), _ptr(sizes),
                                       _ptr(ws["words"]) if need_scan else C.c_void_p(0), _ptr(nbits), _stream()),
            "hyres_jpeg_forward")
    out = {"decoded": dec, "sizes": sizes, "nbits": nbits}
    if want_sizes:
        bpp = torch.empty((), dtype=torch.float32, device=x.device)
        L.check(L.lib().hyres_jpeg_bpp(_ptr(sizes), B, B * H * W, _ptr(bpp), _stream()), "hyres_jpeg_bpp")
        out["bpp"] = bpp
    if want_scan:
        out["words"] = ws["words"].view(B, ws["wpi"])
    return out


def jpeg_assemble(words_host, nbits, H, W, quality):
    """Host: one image's scan words (int32 numpy / CPU tensor, as produced by jpeg_forward) -> JPEG file bytes."""
    import numpy as np
    w = np.ascontiguousarray(words_host, dtype=np.int32)
    cap = int(L.lib().hyres_jpeg_header_bytes()) + 2 * ((int(nbits) + 7) // 8) + 16
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_int64(0)
    L.check(L.lib().hyres_jpeg_assemble(C.c_void_p(w.ctypes.data), int(nbits), H, W, int(quality),
                                        C.c_void_p(out.ctypes.data), cap, C.byref(n)), "hyres_jpeg_assemble")
    return out[:n.value].tobytes()


# ---------------------------------------------------------------------------------------------------------------
# Device-resident entropy coder (csrc/rans_dev.cu): the host coder's byte strings, one warp per string
# ---------------------------------------------------------------------------------------------------------------
_CODER_TLS = __import__("threading").local()


def _pinned_bytes(slot, nbytes):
    """This thread's cached pinned uint8 staging buffer ``slot`` with at least ``nbytes`` bytes."""
    bufs = _CODER_TLS.__dict__.setdefault("pinned", {})
    buf = bufs.get(slot)
    if buf is None or buf.numel() < nbytes:
        buf = bufs[slot] = torch.empty(max(int(nbytes * 1.25) + 4096, 1 << 16), dtype=torch.uint8).pin_memory()
    return buf


class _coder_stream:
    """``with _coder_stream(dev):`` -- launches inside go to this thread's HIGH-PRIORITY side stream, ordered after the
    work already in the current stream; the current stream continues after them.  A coder block needs most of an SM to
    itself; while convolution kernels of other streams keep the SMs full, the block scheduler hands an SM that frees
    up to the highest-priority pending block -- without the priority a coder launch can wait behind a long chain of
    convolution CTAs."""

    def __init__(self, device):
        self.dev = torch.device(device)

    def __enter__(self):
        side = _CODER_TLS.__dict__.setdefault("side", {})
        st = side.get(self.dev.index)
        if st is None:
            st = side[self.dev.index] = torch.cuda.Stream(device=self.dev, priority=-1)
        self.cur = torch.cuda.current_stream(self.dev)
        self.side = st
        st.wait_stream(self.cur)
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return st

    def __exit__(self, *exc):
        self.ctx.__exit__(*exc)
        self.cur.wait_stream(self.side)


def rans_encode_device(groups, escape_room=False):
    """groups: list of ``(symbols, index, tables, slots)`` -- int32 CUDA tensors ``[B, ...]`` of equal shape (``index`` =
    CDF row per symbol, or coder slots from ``gc_symbols`` when ``slots``), ``tables`` a ``coder.DeviceTables`` ->
    list (per group) of lists of B byte strings, identical to ``coder.encode_batch``'s.  One kernel launch per group on
    the current stream (every string is coded by one warp), then two small device -> host copies: the string table
    and the bytes.  ``escape_room``: size the working space for streams made of escapes (the retry after an
    overflow)."""
    dev = groups[0][1].device
    n_str = sum(int(g[1].size(0)) for g in groups)
    meta = torch.zeros(2 + 2 * n_str, dtype=torch.int32, device=dev)
    plans, dst_cap = [], 0
    for symbols, index, tables, slots in groups:
        B = int(index.size(0))
        n = index.numel() // max(B, 1)
        if symbols is not None and symbols.shape != index.shape:
            raise ValueError("`symbols` and `indexes` should have the same size.")
        for t in (symbols, index):
            if t is not None and (t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous()):
                raise ValueError("rans_encode_device: expected contiguous int32 CUDA tensors")
        cap = (3 * n + 256) if escape_room else (n // 2 + 4096)
        plans.append((B, n, cap))
        dst_cap += B * cap
    dst = torch.empty(dst_cap, dtype=torch.int32, device=dev)
    keep = []
    for first in range(0, len(groups), 4):  # one launch per four groups: their strings are coded side by side
        part = list(zip(groups, plans))[first:first + 4]
        arr = (L.RansGroup * len(part))()
        for k, ((symbols, index, tables, slots), (B, n, cap)) in enumerate(part):
            scratch = torch.empty(B * cap, dtype=torch.int32, device=dev)
            keep.append(scratch)
            g = arr[k]
            g.symbols, g.index = symbols.data_ptr() if symbols is not None else None, index.data_ptr()
            g.enc, g.rows, g.scratch = tables.enc.data_ptr(), tables.rows.data_ptr(), scratch.data_ptr()
            g.n, g.cap_words, g.n_entries = n, cap, tables.n_entries
            g.n_rows, g.count, g.slots = tables.n_rows, B, int(bool(slots))
        base = sum(pl[0] for pl in plans[:first])
        with _coder_stream(dev):
            L.check(L.lib().hyres_rans_dev_encode(len(part), C.cast(arr, C.c_void_p), _ptr(dst), dst_cap, _ptr(meta),
                                                  base, _stream()), "hyres_rans_dev_encode")
    hmeta = _pinned_bytes("meta", meta.numel() * 4)[: meta.numel() * 4].view(torch.int32)
    hmeta.copy_(meta, non_blocking=True)
    stream_wait_blocking(dev)
    m = hmeta.numpy()
    total, status = int(m[0]), int(m[1])
    if status == 2 and not escape_room:
        return rans_encode_device(groups, escape_room=True)
    if status != 0:
        raise L.HyresError("hyres_rans_dev_encode: " + ("output buffer too small" if status == 2 else
                                                        "symbol / index outside the CDF tables"))
    hbytes = _pinned_bytes("bytes", total * 4)[: total * 4]
    hbytes.view(torch.int32).copy_(dst[:total], non_blocking=True)
    stream_wait_blocking(dev)
    raw = hbytes.numpy()
    out, k = [], 0
    for B, _, _ in plans:
        grp = []
        for _ in range(B):
            off, ln = int(m[2 + 2 * k]), int(m[3 + 2 * k])
            grp.append(raw[off * 4:(off + ln) * 4].tobytes())
            k += 1
        out.append(grp)
    return out


def rans_upload(string_groups, device):
    """list of lists of byte strings -> (words int32 CUDA tensor with every string back to back, string table int64
    CUDA ``[2, S]`` = first word and word count of each): one host -> device copy for all passes of a decompress."""
    import numpy as np
    strings = [s for grp in string_groups for s in grp]
    S = len(strings)
    lens = [len(s) for s in strings]
    for n in lens:
        if n % 4 or n < 8:
            raise ValueError("Invalid strings: a rANS64 stream is a whole number (>= 2) of 32-bit words")
    head = 16 * S
    total = head + sum(lens)
    buf = _pinned_bytes("up", total)[:total]
    arr = buf.numpy()
    table = arr[:head].view(np.int64).reshape(2, S)
    pos = 0
    for k, s in enumerate(strings):
        table[0, k] = pos // 4
        table[1, k] = lens[k] // 4
        arr[head + pos: head + pos + lens[k]] = np.frombuffer(s, dtype=np.uint8)
        pos += lens[k]
    d = torch.empty(total, dtype=torch.uint8, device=device)
    d.copy_(buf, non_blocking=True)
    # the pinned buffer is reused by this thread's next upload: the copy must have left it
    stream_wait_blocking(device)
    return d[head:].view(torch.int32), d[:head].view(torch.int64).view(2, S)


def rans_decode_device(words, table, first, codes, tables, has_codes, status):
    """Decode strings ``first .. first + B`` of an uploaded set (``rans_upload``) -> int32 CUDA tensor shaped like
    ``codes`` ``[B, ...]`` (CDF row per symbol, or decoder codes from ``gc_codes`` when ``has_codes``: the entries of
    known symbols are then left unwritten).  ``status``: int32 CUDA tensor [1], zeroed by the caller; non-zero after
    a malformed stream.  One launch on the current stream, no synchronisation."""
    if codes.dtype != torch.int32 or not codes.is_cuda or not codes.is_contiguous():
        raise ValueError("rans_decode_device: expected a contiguous int32 CUDA tensor of codes")
    B = int(codes.size(0))
    n = codes.numel() // max(B, 1)
    if first + B > table.size(1):
        raise ValueError("Invalid strings or indexes parameters")
    out = torch.empty_like(codes)
    off, ln = table[0, first:first + B], table[1, first:first + B]
    with _coder_stream(codes.device):
        L.check(L.lib().hyres_rans_dev_decode(_ptr(words), _ptr(off), _ptr(ln), _ptr(codes), B, n, int(bool(has_codes)),
                                              _ptr(tables.sf), _ptr(tables.rows), tables.n_rows, tables.n_entries,
                                              _ptr(out), _ptr(status), _stream()), "hyres_rans_dev_decode")
    return out
