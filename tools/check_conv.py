"""GPU self-check of the tcgen05 conv kernel against torch (fp32 conv on bf16-rounded
operands).  Diagnostic tool for `gpurun`; the pytest parity tests live in tests/.

usage: python tools/check_conv.py            # run every case, one subprocess each
       python tools/check_conv.py --case N   # run one case in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, kind, cin0, cin1, cout, k, stride, dil, B, H, W, epi, act, extras
    dict(name="1x1_64_64_tile", cin0=64, cout=64, k=1, B=1, H=16, W=8),
    dict(name="1x1_128_64", cin0=128, cout=64, k=1, B=2, H=32, W=16),
    dict(name="3x3_64_64", cin0=64, cout=64, k=3, B=1, H=32, W=24),
    dict(name="3x3_64_64_ragged", cin0=64, cout=64, k=3, B=2, H=20, W=12),
    dict(name="3x3_dil2", cin0=64, cout=64, k=3, dil=2, B=1, H=32, W=24),
    dict(name="5x5_s2_128_128", cin0=128, cout=128, k=5, stride=2, B=1, H=32, W=32),
    dict(name="5x5_s2_128_192_ragged", cin0=128, cout=192, k=5, stride=2, B=2, H=24, W=40),
    dict(name="deconv_128_128", kind=1, cin0=128, cout=128, k=5, B=1, H=16, W=16),
    dict(name="deconv_192_128_ragged", kind=1, cin0=192, cout=128, k=5, B=2, H=12, W=20),
    dict(name="deconv_128_3_f32nchw", kind=1, cin0=128, cout=3, k=5, B=1, H=16, W=24, f32="nchw", bf16=False),
    dict(name="1x1_96_192", cin0=96, cout=192, k=1, B=1, H=16, W=16),
    dict(name="3x3_96_96", cin0=96, cout=96, k=3, B=1, H=16, W=16),
    dict(name="1x1_two_input_768_640", cin0=384, cin1=384, cout=640, k=1, B=1, H=16, W=16),
    dict(name="1x1_anchor_384of768_640", cin0=384, wcin=768, cout=640, k=1, B=1, H=16, W=16),
    dict(name="1x1_640_512", cin0=640, cout=512, k=1, B=1, H=16, W=16),
    dict(name="5x5_masked_192_384", cin0=192, cout=384, k=5, mask=True, B=1, H=16, W=24, f32="nhwc"),
    dict(name="3x3_192_384", cin0=192, cout=384, k=3, B=1, H=16, W=16),
    dict(name="epi_add_relu", cin0=64, cout=128, k=1, B=1, H=32, W=16, epi="add", act="relu"),
    dict(name="epi_gate", cin0=128, cout=128, k=1, B=1, H=32, W=16, epi="gate", f32="nhwc"),
    dict(name="epi_gdn", cin0=128, cout=128, k=1, B=1, H=32, W=16, epi="gdn"),
    dict(name="epi_igdn", cin0=128, cout=128, k=1, B=1, H=32, W=16, epi="igdn"),
    dict(name="epi_pixscale_prelu", cin0=192, cout=64, k=1, B=1, H=32, W=16, epi="pix", act="prelu"),
    dict(name="gdn_x0_square", cin0=128, cout=128, k=1, B=2, H=40, W=20, epi="gdn", square=True),
    dict(name="igdn_x0_square", cin0=128, cout=128, k=1, B=1, H=16, W=24, epi="igdn", square=True),
    dict(name="epi_gate_ragged", cin0=128, cout=128, k=1, B=3, H=24, W=20, epi="gate"),
    dict(name="epi_gate_192_f32", cin0=192, cout=192, k=1, B=2, H=16, W=24, epi="gate", f32="nhwc"),
    dict(name="3x3_64_64_many_tiles", cin0=64, cout=64, k=3, B=3, H=80, W=72, act="prelu"),
    dict(name="3x3_dil2_ragged", cin0=64, cout=64, k=3, dil=2, B=2, H=40, W=20, act="prelu"),
    dict(name="1x1_192_96_relu", cin0=192, cout=96, k=1, B=2, H=16, W=24, act="relu"),
    dict(name="1x1_96_192_add_relu", cin0=96, cout=192, k=1, B=2, H=16, W=24, epi="add", act="relu"),
    dict(name="deconv_128_3_many_tiles", kind=1, cin0=128, cout=3, k=5, B=2, H=40, W=36, f32="nchw", bf16=False, act="clamp"),
    dict(name="out_sq", cin0=128, cout=128, k=1, B=1, H=16, W=16, sq=True),
    dict(name="3x3_64_3_clamp", cin0=64, cout=3, k=3, B=1, H=32, W=16, act="clamp", f32="nchw", bf16=False),
    dict(name="mt2_3x3", cin0=64, cout=64, k=3, B=1, H=64, W=32, mt=2),
    dict(name="mt4_3x3", cin0=64, cout=64, k=3, B=1, H=128, W=32, mt=4),
    dict(name="mt2_5x5s2", cin0=128, cout=128, k=5, stride=2, B=1, H=128, W=64, mt=2),
    dict(name="mt4_deconv", kind=1, cin0=128, cout=128, k=5, B=1, H=64, W=32, mt=4),
    dict(name="auto_big_gate", cin0=128, cout=128, k=1, B=4, H=256, W=384, epi="gate", perf=True),
    dict(name="auto_big_1x1_192_64_pix", cin0=192, cout=64, k=1, B=2, H=512, W=768, epi="pix", act="prelu", perf=True),
    dict(name="auto_big_deconv_128_3", kind=1, cin0=128, cout=3, k=5, B=4, H=256, W=384, f32="nchw", bf16=False, perf=True),
    dict(name="auto_big_gdn_sq", cin0=128, cout=128, k=1, B=4, H=256, W=384, epi="gdn", square=True, perf=True),
    dict(name="auto_big_3x3", cin0=64, cout=64, k=3, B=4, H=256, W=384, perf=True),
    dict(name="auto_big_1x1_128_64", cin0=128, cout=64, k=1, B=4, H=256, W=384, perf=True),
    dict(name="auto_big_5x5s2", cin0=128, cout=128, k=5, stride=2, B=4, H=256, W=384, perf=True),
    dict(name="auto_big_deconv", kind=1, cin0=128, cout=128, k=5, B=4, H=128, W=192, perf=True),
]


def run_case(idx):
    import torch
    import torch.nn.functional as F
    from hyres_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CASES[idx]
    g = torch.Generator(device="cpu").manual_seed(1926 + idx)
    kind = c.get("kind", 0)
    cin0, cin1, cout, k = c["cin0"], c.get("cin1", 0), c["cout"], c["k"]
    wcin = c.get("wcin", cin0 + cin1)
    stride, dil = c.get("stride", 1), c.get("dil", 1)
    B, H, W = c["B"], c["H"], c["W"]
    pad = dil * (k - 1) // 2 if kind == 0 else 2
    dev = "cuda"
    if kind == 0:
        w = torch.randn(cout, wcin, k, k, generator=g) / (wcin * k * k) ** 0.5
    else:
        w = torch.randn(wcin, cout, k, k, generator=g) / (wcin * k * k / 4) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mask = None
    if c.get("mask"):
        mask = torch.zeros(k, k, dtype=torch.uint8)
        mask[0::2, 1::2] = 1
        mask[1::2, 0::2] = 1
    epi = c.get("epi", "lin")
    if epi in ("gdn", "igdn"):
        w = w.abs()
        bias = bias.abs() + 0.5
    x0 = torch.randn(B, H, W, cin0, generator=g).to(dev).bfloat16()
    if epi in ("gdn", "igdn"):
        x0 = x0.abs()
    x1 = torch.randn(B, H, W, cin1, generator=g).to(dev).bfloat16() if cin1 else None
    layer = ops.ConvLayer(w, bias, kind=kind, stride=stride, pad=pad, dil=dil, cin0=cin0, cin1=cin1,
                          tap_mask=mask)
    OH, OW = layer.out_size(H, W)
    aux0 = aux1 = pix = None
    if epi in ("add", "gate", "gdn", "igdn"):
        aux0 = torch.randn(B, OH, OW, cout, generator=g).to(dev).bfloat16()
    if c.get("square"):
        aux0 = x0
    if epi == "gate":
        aux1 = torch.randn(B, OH, OW, cout, generator=g).to(dev).bfloat16()
    if epi == "pix":
        pix = torch.rand(B, OH, OW, generator=g).to(dev)
    epi_id = dict(lin=0, add=1, gate=2, gdn=3, igdn=4, pix=5)[epi]
    act_id = dict(none=0, relu=1, prelu=2, clamp=3)[c.get("act", "none")]
    slope = 0.25
    o16, osq, o32 = layer(x0, x1, epi=epi_id, act=act_id, slope=slope, aux0=aux0, aux1=aux1,
                          pixscale=pix, out_bf16=c.get("bf16", True), out_sq=c.get("sq", False),
                          out_f32=c.get("f32"), mt=c.get("mt", 0), x0_square=c.get("square", False))
    torch.cuda.synchronize()

    # reference: fp32 math on bf16-rounded operands
    wq = w.to(dev).bfloat16().float()
    if mask is not None:
        wq = wq * mask.to(dev).float()
    xin = x0.float()
    if c.get("square"):
        xin = (xin * xin).bfloat16().float()
    if cin1:
        xin = torch.cat([xin, x1.float()], dim=-1)
    xin = xin.permute(0, 3, 1, 2)
    if kind == 0:
        wuse = wq[:, : cin0 + cin1]
        ref = F.conv2d(xin, wuse, None, stride=stride, padding=pad, dilation=dil)
    else:
        ref = F.conv_transpose2d(xin, wq, None, stride=2, padding=2, output_padding=1)
    ref = ref.permute(0, 2, 3, 1)
    bq = bias.to(dev)
    if epi == "pix":
        ref = ref * pix[..., None]
    ref = ref + bq
    if epi == "add":
        ref = ref + aux0.float()
    elif epi == "gate":
        ref = aux1.float() * torch.sigmoid(ref) + aux0.float()
    elif epi == "gdn":
        ref = aux0.float() * torch.rsqrt(ref)
    elif epi == "igdn":
        ref = aux0.float() * torch.sqrt(ref)
    if act_id == 1:
        ref = ref.relu()
    elif act_id == 2:
        ref = torch.where(ref >= 0, ref, ref * slope)
    elif act_id == 3:
        ref = ref.clamp(0, 1)

    res = dict(name=c["name"], idx=idx)
    scale = ref.abs().max().item()

    def cmp(got, want, tag, tol):
        d = (got - want).abs()
        res[tag + "_maxabs"] = d.max().item()
        res[tag + "_rel"] = d.max().item() / max(scale, 1e-9)
        bad = d > tol * max(scale, 1e-9)
        res[tag + "_nbad"] = int(bad.sum().item())
        if bad.any():
            nz = bad.nonzero()
            res[tag + "_firstbad"] = nz[:6].tolist()
            res[tag + "_bad_by_dim"] = [int(bad.sum(dim=[j for j in range(4) if j != i]).ne(0).sum()) for i in range(4)]
        return not bad.any().item()

    ok = True
    if o32 is not None:
        got = o32 if c.get("f32") == "nhwc" else o32.permute(0, 2, 3, 1)
        ok &= cmp(got, ref, "f32", 2e-4)
    if o16 is not None:
        ok &= cmp(o16.float(), ref, "bf16", 1e-2)
    if osq is not None:
        want = ref.bfloat16().float() ** 2
        d = (osq.float() - want).abs().max().item() / max(want.abs().max().item(), 1e-9)
        res["sq_rel"] = d
        ok &= d < 1e-2
    res["ok"] = bool(ok)
    if c.get("perf"):
        kw = dict(epi=epi_id, act=act_id, slope=slope, aux0=aux0, aux1=aux1, pixscale=pix,
                  x0_square=c.get("square", False))
        if o16 is None:
            kw.update(out_bf16=False, out_f32=o32 if c.get("f32") == "nhwc" else o32.permute(0, 2, 3, 1))
        else:
            kw.update(out_bf16=o16)
        for _ in range(3):
            layer(x0, x1, **kw)
        torch.cuda.synchronize()
        for mt in (0,):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n):
                layer(x0, x1, mt=mt, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            flops = 2.0 * layer.macs_per_pos * B * (OH * OW if kind == 0 else H * W * 4)
            res[f"ms_mt{mt}"] = ms
            res[f"tflops_mt{mt}"] = flops / ms / 1e9
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-case", type=int, default=-1)
    ap.add_argument("--only", type=str, default="")
    ap.add_argument("--out", type=str, default="gpurun_out/check_conv.jsonl")
    a = ap.parse_args()
    sel = [i for i, c in enumerate(CASES) if not a.only or a.only in c["name"]]
    if a.from_case >= 0:
        # in-process: run the selected cases from index `from_case` on; a CUDA fault ends the
        # process (the context is gone) and the parent restarts after the faulting case.
        for i in sel:
            if i < a.from_case:
                continue
            print(f"BEGIN {i}", flush=True)
            try:
                r = run_case(i)
            except Exception as e:  # noqa: BLE001
                r = dict(name=CASES[i]["name"], idx=i, ok=False, err=repr(e)[-500:])
                print("RESULT " + json.dumps(r), flush=True)
                if "CUDA" in repr(e) or "cuda" in repr(e):
                    return 2
                continue
            print("RESULT " + json.dumps(r), flush=True)
        return 0
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results, start = {}, 0
    while True:
        todo = [i for i in sel if i >= start]
        if not todo:
            break
        cmd = [sys.executable, __file__, "--from-case", str(todo[0])]
        if a.only:
            cmd += ["--only", a.only]
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            out, err = p.stdout, p.stderr
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            err = "timeout"
        last_begin = None
        for l in out.splitlines():
            if l.startswith("BEGIN "):
                last_begin = int(l[6:])
            elif l.startswith("RESULT "):
                r = json.loads(l[7:])
                results[r["idx"]] = r
        if last_begin is None:
            results[todo[0]] = dict(name=CASES[todo[0]]["name"], idx=todo[0], ok=False, err=(err or out)[-800:])
            start = todo[0] + 1
        elif last_begin not in results:
            results[last_begin] = dict(name=CASES[last_begin]["name"], idx=last_begin, ok=False,
                                       err="process died: " + (err or "")[-600:])
            start = last_begin + 1
        elif last_begin == todo[-1]:
            break
        else:
            start = last_begin + 1
    nfail = 0
    with open(a.out, "w") as fh:
        for i in sorted(results):
            r = results[i]
            nfail += 0 if r.get("ok") else 1
            fh.write(json.dumps(r) + "\n")
            print(json.dumps(r))
    print(f"check_conv: {nfail} failing case(s) of {len(results)}")
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
