import os, sys, ctypes
sys.path.insert(0, '/root/repo')
import torch
import hyres_b200
from hyres_b200 import _lib, synthetic, train as T
rt = ctypes.CDLL("libcudart.so.12")
def cap_status():
    st = ctypes.c_int(0)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value
# wrap check to report the first library call after which the capture is invalidated
orig_check = _lib.check
state = {"bad": None}
def check(rc, what=""):
    if state["bad"] is None and torch.cuda.is_current_stream_capturing():
        r, s = cap_status()
        if r != 0 or s == 2:
            state["bad"] = what
            print("CAPTURE INVALIDATED at/just before", what, "rc", r, "status", s, flush=True)
    return orig_check(rc, what)
_lib.check = check
from hyres_b200 import ops
ops.L.check = check
torch.manual_seed(0)
net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1); net.update(force=True); net = net.cuda()
tr = T.Trainer(net, capturable=True)
x = synthetic.synthetic_image(2, 64, 64, seed=1).cuda()
for _ in range(2): tr._step_impl(x, True, None, None)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): tr._step_impl(x, True, None, None)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
def try_capture(name, fn):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=side):
            fn()
            r, s = cap_status()
            print(name, "status at end of capture body:", r, s, flush=True)
        g.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception as e:
        print(name, "FAILED", type(e).__name__, str(e)[:200], flush=True)
def fwd():
    with torch.no_grad():
        tr.graph.forward(x, noisequant=True, training=True)
def fwd_bwd():
    out = tr.graph.forward(x, noisequant=True, training=True)
    c = T.rd_loss(out, x, 0.008)
    tr.optimizer.zero_grad(set_to_none=True)
    c["loss"].backward()
    r, s = cap_status(); print(" after backward", r, s, flush=True)
try_capture("forward", fwd)
try_capture("forward+backward", fwd_bwd)
try_capture("full step", lambda: tr._step_impl(x, True, None, None))

def staged():
    G = tr.graph
    codec = G.codec
    def st(tag):
        r, s = cap_status(); print("  ", tag, r, s, flush=True)
    with torch.no_grad():
        jd, bpp = net.jpeg.forward_device(x); st("jpeg")
        res = x - jd; st("sub")
        y = G.g_a(res.permute(0, 2, 3, 1).to(T.BF16)); st("g_a")
        z = G._seq(y.to(T.BF16), codec.h_a, out_f32_last=True); st("h_a")
        nf = lambda shape, tag: torch.empty(shape, device="cuda").uniform_(-0.5, 0.5)
        z_hat, z_lik = G._eb(z, nf, True, True); st("eb")
        latent = G._seq(z_hat.to(T.BF16).contiguous(), codec.h_s); st("h_s")
        pa = G._head(latent, torch.zeros_like(latent)); st("head")
        ctx = G._context(y.to(T.BF16)); st("ctx")
        r_hat = G.g_s(y.to(T.BF16).contiguous()); st("g_s")
        lik = G._gc_likelihood(y, pa[..., :192], pa[..., 192:], nf, True); st("lik")
        ref = G.refine(jd); st("refine")
try_capture("staged forward", staged)
