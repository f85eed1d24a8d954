"""GPU check of the split-precision (fp32-equivalent) trunk.

Part A (``conv_cases``): every layer geometry of the entropy-critical trunk as a split convolution
(``ops.ConvLayer(nsplit=2|3|18)``) against a float64 convolution of the same fp32 operands; the error of a plain
fp32 cuDNN convolution against the same float64 result is printed beside it (that is the noise floor two fp32
implementations of the reference differ by).

Part B (``symbol_report``): ``LightWeightCheckerboard.encode_symbols`` on the GPU against the CPU oracle in its
fp32 mode (the reference's semantics, oracle/hyres_oracle.py) on the same weights and input: exact-match
fractions of the three symbol streams and two index streams, and for every mismatch the distance of the oracle's
own value from the rounding tie / table edge that decides it ("near-tie": the two fp32 evaluations straddle it).

usage: python tools/check_precise.py [--nsplit 3] [--out gpurun_out/precise.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name, kind (0 conv / 1 deconv k5s2), cin0, cin1, wcin, cout, k, stride, dil, mask, B, H, W
CONV_CASES = [
    dict(name="ga0_im2col_1x1_128_128", cin0=128, cout=128, k=1, B=2, H=40, W=24),
    dict(name="ru_1x1_128_64", cin0=128, cout=64, k=1, B=2, H=40, W=24, relu=True),
    dict(name="ru_3x3_64_64", cin0=64, cout=64, k=3, B=2, H=40, W=24, relu=True),
    dict(name="ru_1x1_64_128", cin0=64, cout=128, k=1, B=2, H=40, W=24),
    dict(name="ga4_5x5s2_128_128", cin0=128, cout=128, k=5, stride=2, B=2, H=48, W=32),
    dict(name="ga7_5x5s2_128_192", cin0=128, cout=192, k=5, stride=2, B=1, H=24, W=40),
    dict(name="ru192_1x1_192_96", cin0=192, cout=96, k=1, B=2, H=16, W=24, relu=True),
    dict(name="ru192_3x3_96_96", cin0=96, cout=96, k=3, B=2, H=16, W=24, relu=True),
    dict(name="ru192_1x1_96_192", cin0=96, cout=192, k=1, B=2, H=16, W=24),
    dict(name="ha0_3x3_192_128", cin0=192, cout=128, k=3, B=1, H=16, W=24, relu=True),
    dict(name="hs0_deconv_128_128", kind=1, cin0=128, cout=128, k=5, B=2, H=6, W=10, relu=True),
    dict(name="hs1_deconv_128_192", kind=1, cin0=128, cout=192, k=5, B=1, H=12, W=20, relu=True),
    dict(name="hs2_3x3_192_384", cin0=192, cout=384, k=3, B=1, H=16, W=24),
    dict(name="ctx_5x5_masked_192_384", cin0=192, cout=384, k=5, mask=True, B=1, H=16, W=24),
    dict(name="head0_two_input_768_640", cin0=384, cin1=384, cout=640, k=1, B=1, H=16, W=24, relu=True),
    dict(name="head0_anchor_384of768_640", cin0=384, wcin=768, cout=640, k=1, B=1, H=16, W=24, relu=True),
    dict(name="head1_1x1_640_512", cin0=640, cout=512, k=1, B=1, H=16, W=24, relu=True),
    dict(name="head2_1x1_512_384", cin0=512, cout=384, k=1, B=1, H=16, W=24),
]


def _torch_ref(x0, x1, w, bias, c, dtype):
    import torch
    import torch.nn.functional as F
    kind, k = c.get("kind", 0), c["k"]
    stride, dil = c.get("stride", 1), c.get("dil", 1)
    x = x0 if x1 is None else torch.cat([x0, x1], -1)
    x = x.permute(0, 3, 1, 2).to(dtype)
    w = w.to(dtype)
    if c.get("mask"):
        m = torch.zeros(k, k, dtype=dtype, device=w.device)
        m[0::2, 1::2] = 1
        m[1::2, 0::2] = 1
        w = w * m
    if kind == 0:
        w = w[:, :x.shape[1]]
        y = F.conv2d(x, w, bias.to(dtype), stride=stride, padding=dil * (k - 1) // 2, dilation=dil)
    else:
        y = F.conv_transpose2d(x, w, bias.to(dtype), stride=2, padding=2, output_padding=1)
    if c.get("relu"):
        y = y.clamp_min(0)
    return y.permute(0, 2, 3, 1)


def conv_case(idx, nsplit):
    import torch
    from hyres_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CONV_CASES[idx]
    g = torch.Generator(device="cpu").manual_seed(4242 + idx)
    kind = c.get("kind", 0)
    cin0, cin1, cout, k = c["cin0"], c.get("cin1", 0), c["cout"], c["k"]
    wcin = c.get("wcin", cin0 + cin1)
    stride, dil = c.get("stride", 1), c.get("dil", 1)
    B, H, W = c["B"], c["H"], c["W"]
    pad = dil * (k - 1) // 2 if kind == 0 else 2
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
    # activations with a wide dynamic range, like the codec's (|y| up to ~12 next to values ~1e-3)
    x0 = (torch.randn(B, H, W, cin0, generator=g) * torch.exp(torch.randn(B, H, W, cin0, generator=g))).cuda()
    x1 = (torch.randn(B, H, W, cin1, generator=g) * 3).cuda() if cin1 else None
    layer = ops.ConvLayer(w, bias, kind=kind, stride=stride, pad=pad, dil=dil, cin0=cin0, cin1=cin1, tap_mask=mask,
                          nsplit=nsplit)
    _, s0 = ops.split_f32(x0, nsplit=nsplit)
    s1 = ops.split_f32(x1, nsplit=nsplit)[1] if cin1 else None
    _, _, got = layer(s0, s1, act=ops.ACT_RELU if c.get("relu") else ops.ACT_NONE, out_bf16=False, out_f32="nhwc")
    torch.cuda.synchronize()
    wd, bd = w.cuda(), bias.cuda()
    ref64 = _torch_ref(x0, x1, wd, bd, c, torch.float64)
    ref32 = _torch_ref(x0, x1, wd, bd, c, torch.float32)
    scale = float(ref64.abs().max())
    # the split of the input itself: parts must add up to the fp32 value
    if nsplit & ops.SPLIT_F16:
        hp = s0.view(torch.float16).float().reshape(B, H, W, 2, cin0)
        parts = hp[:, :, :, 0] + hp[:, :, :, 1] / 2048.0
    else:
        parts = s0.float().reshape(B, H, W, nsplit, cin0).sum(3)
    split_err = float((parts - x0).abs().max() / x0.abs().max())
    err = float((got.double() - ref64).abs().max()) / scale
    err32 = float((ref32.double() - ref64).abs().max()) / scale
    rms = float((got.double() - ref64).pow(2).mean().sqrt()) / scale
    rms32 = float((ref32.double() - ref64).pow(2).mean().sqrt()) / scale
    exact = nsplit != 2  # three bf16 parts or two half parts: fp32-equivalent
    tol = 3e-6 if exact else 2e-4
    return dict(name=c["name"], nsplit=nsplit, max_err=err, max_err_fp32_cudnn=err32, rms_err=rms, rms_err_fp32_cudnn=rms32,
                split_err=split_err, ok=bool(err < tol and split_err < (1e-6 if exact else 1e-4)))


def near_tie_distance_symbols(t):
    """|frac(t) - 0.5| for t = (value - mean): 0 at an exact rounding tie."""
    import torch
    return ((t - torch.floor(t)) - 0.5).abs()


def _stream_stats(k, got, want, dist, eps):
    """Mismatch statistics of one integer stream; ``dist``: distance of the oracle's own pre-rounding value from the
    tie / table edge that decides the integer."""
    bad = got != want
    n_bad = int(bad.sum())
    ent = dict(n=int(want.numel()), mismatches=n_bad, match=1.0 - n_bad / want.numel(),
               near_ties_in_stream=int((dist < eps).sum()))
    if n_bad:
        ent["max_tie_distance_of_mismatches"] = float(dist[bad].max())
        ent["max_abs_diff"] = int((got - want).abs().max())
        ent["unexplained"] = int((dist[bad] >= eps).sum())
    else:
        ent["unexplained"] = 0
    return ent


def symbol_report(pnet, onet, oracle, x, sym_eps=2e-4, idx_eps=2e-4):
    """pnet: product LightWeightCheckerboard (cuda); onet: oracle LightWeightCheckerboard (cpu); x: fp32 NCHW cpu
    residual.  The oracle runs in fp32 mode (the reference's semantics).  Two comparisons:

    * ``end_to_end``: the oracle's own compress() against the product's streams.  One flipped anchor symbol changes
      the context of ~12 x 192 non-anchor elements, so second-pass mismatches include the legitimate consequences
      of first-pass near-ties.
    * ``stagewise``: every stage of the oracle is fed the *product's* integers of the stage before (symbols are
      exact in fp32), so each stream is compared on identical stage inputs; every remaining mismatch must sit on
      a rounding tie (|frac(v) - 0.5| < sym_eps) or a scale-table edge (relative distance < idx_eps) of the
      oracle's own value -- ``unexplained`` counts those that do not."""
    import torch
    with torch.no_grad():
        s = pnet.encode_symbols(x.cuda())
        with oracle.precision("fp32"):
            oc = onet.compress(x, return_intermediates=True)
    nchw = lambda t: t.permute(0, 3, 1, 2).float().cpu()  # noqa: E731
    rep = dict(shape=list(x.shape), precision=pnet.codec_precision)

    def rel(got, want):
        return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))

    y_o, z_o = oc["_y"], oc["_z"]
    rep["y_rel_err"] = rel(nchw(s["y"]), y_o)
    rep["z_rel_err"] = rel(nchw(s["z"]), z_o)
    rep["params_a_rel_err"] = rel(nchw(s["params_a"]), oc["_anchor_params"])
    rep["params_na_rel_err"] = rel(nchw(s["params_na"]), oc["_non_anchor_params"])
    M = y_o.shape[1]
    B, _, h, w = y_o.shape
    ii = torch.arange(h).view(1, 1, h, 1)
    jj = torch.arange(w).view(1, 1, 1, w)
    anchor = ((ii + jj) % 2 == 0).expand(B, M, h, w)
    y_a = torch.where(anchor, y_o, torch.zeros_like(y_o))
    y_na = torch.where(anchor, torch.zeros_like(y_o), y_o)
    gc, eb = onet.gaussian_conditional, onet.entropy_bottleneck
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    table = gc.scale_table.detach()
    bound = float(gc.scale_bound)

    def edge_distance(scales):
        sc = scales.clamp_min(bound)
        return ((sc.unsqueeze(-1) - table.view(1, 1, 1, 1, -1)).abs() / table.view(1, 1, 1, 1, -1)).min(-1).values

    got = {k: s[k].cpu() for k in ("sym_z", "sym_a", "sym_na", "idx_a", "idx_na")}
    # -- end to end --
    pa_o, pna_o = oc["_anchor_params"], oc["_non_anchor_params"]
    e2e = {
        "sym_z": _stream_stats("sym_z", got["sym_z"], oc["_sym_z"].int(), near_tie_distance_symbols(z_o - med), sym_eps),
        "sym_a": _stream_stats("sym_a", got["sym_a"], oc["_sym_a"].int(), near_tie_distance_symbols(y_a - pa_o[:, M:]), sym_eps),
        "idx_a": _stream_stats("idx_a", got["idx_a"], oc["_idx_a"].int(), edge_distance(pa_o[:, :M]), idx_eps),
        "sym_na": _stream_stats("sym_na", got["sym_na"], oc["_sym_na"].int(), near_tie_distance_symbols(y_na - pna_o[:, M:]), sym_eps),
        "idx_na": _stream_stats("idx_na", got["idx_na"], oc["_idx_na"].int(), edge_distance(pna_o[:, :M]), idx_eps),
    }
    rep["end_to_end"] = e2e
    # -- stage by stage, teacher-forced with the product's integers --
    with torch.no_grad(), oracle.precision("fp32"):
        z_hat = got["sym_z"].float() + med
        latent = oracle._run(onet.h_s, z_hat)
        pa_t = oracle._run(onet.param_aggregation, torch.cat([latent, torch.zeros_like(latent)], 1))
        sc_a, mu_a = pa_t.chunk(2, 1)
        sym_a_t = gc.quantize(y_a, "symbols", mu_a)
        idx_a_t = gc.build_indexes(sc_a)
        yq_a = got["sym_a"].float() + mu_a
        ctx = onet.context_prediction(yq_a)
        pna_t = oracle._run(onet.param_aggregation, torch.cat([latent, ctx], 1))
        sc_na, mu_na = pna_t.chunk(2, 1)
        sym_na_t = gc.quantize(y_na, "symbols", mu_na)
        idx_na_t = gc.build_indexes(sc_na)
    rep["params_a_rel_err_stagewise"] = rel(nchw(s["params_a"]), pa_t)
    rep["params_na_rel_err_stagewise"] = rel(nchw(s["params_na"]), pna_t)
    st = {
        "sym_z": e2e["sym_z"],
        "sym_a": _stream_stats("sym_a", got["sym_a"], sym_a_t.int(), near_tie_distance_symbols(y_a - mu_a), sym_eps),
        "idx_a": _stream_stats("idx_a", got["idx_a"], idx_a_t.int(), edge_distance(sc_a), idx_eps),
        "sym_na": _stream_stats("sym_na", got["sym_na"], sym_na_t.int(), near_tie_distance_symbols(y_na - mu_na), sym_eps),
        "idx_na": _stream_stats("idx_na", got["idx_na"], idx_na_t.int(), edge_distance(sc_na), idx_eps),
    }
    rep["stagewise"] = st
    for name, grp in (("end_to_end", e2e), ("stagewise", st)):
        rep[name + "_mismatches"] = sum(v["mismatches"] for v in grp.values())
        rep[name + "_unexplained"] = sum(v["unexplained"] for v in grp.values())
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsplit", type=int, default=0, help="0 = 2, 3 and 18 (two half parts)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "precise.json"))
    ap.add_argument("--skip-model", action="store_true")
    ap.add_argument("--shapes", default="1x64x64,2x96x160,1x256x256")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    out = dict(conv=[], model=[])
    for ns in ([2, 3, 18] if args.nsplit == 0 else [args.nsplit]):
        for i in range(len(CONV_CASES)):
            r = conv_case(i, ns)
            out["conv"].append(r)
            print(json.dumps(r), flush=True)
    if not args.skip_model:
        import hyres_b200
        from oracle import hyres_oracle as O
        torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
        onet = O.make_model(seed=1926, wrapper=True, lively=True)
        pnet = hyres_b200.ResidualJPEGCompression()
        pnet.load_state_dict(onet.state_dict())
        pnet = pnet.cuda().eval()
        for shp in args.shapes.split(","):
            B, H, W = map(int, shp.split("x"))
            x = O.synthetic_image(B, H, W, seed=9)
            jd, _ = onet.jpeg(x)
            res = x - jd
            for mode in ("bf16", "fp32x2", "fp32x3", "fp32h2"):
                pnet.residual_model.codec_precision = mode
                r = symbol_report(pnet.residual_model, onet.residual_model, O, res)
                out["model"].append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
