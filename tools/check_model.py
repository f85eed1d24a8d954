"""GPU end-to-end check of the product model against the CPU oracle (diagnostic tool for gpurun).

Prints per-stage error statistics in both oracle precisions, symbol / index match rates,
string equality and a few timings.  The pytest parity tests in tests/ assert on the same
quantities; this script is for reading numbers.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import hyres_b200  # noqa: E402
from oracle import hyres_oracle as O  # noqa: E402


def nchw(t):
    return t.permute(0, 3, 1, 2).float().cpu()


def err(name, got, want, out):
    d = (got - want).abs()
    scale = want.abs().max().item()
    out[name] = dict(maxabs=d.max().item(), rel=d.max().item() / max(scale, 1e-12), mean=d.mean().item(), scale=scale)
    print(f"  {name:28s} maxabs {d.max().item():.3e}  rel {d.max().item() / max(scale, 1e-12):.3e}  mean {d.mean().item():.3e}  scale {scale:.3g}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=256)
    ap.add_argument("--W", type=int, default=256)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/check_model.json")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {}
    onet = O.make_model(wrapper=True, lively=True)
    pnet = hyres_b200.ResidualJPEGCompression()
    pnet.load_state_dict(onet.state_dict())
    pnet = pnet.cuda().eval()
    ocodec, pcodec = onet.residual_model, pnet.residual_model

    x = O.synthetic_image(a.B, a.H, a.W)
    jpeg_dec, jpeg_bpp = onet.jpeg(x)
    residual = x - jpeg_dec
    print("residual std", residual.std().item(), "jpeg bpp", jpeg_bpp)

    # ---------------- codec stages via the compress front-end ----------------
    with torch.no_grad():
        s = pcodec.encode_symbols(residual.cuda())
        torch.cuda.synchronize()
        for mode in ("bf16", "fp32"):
            print(f"[codec stages vs oracle {mode}]")
            with O.precision(mode):
                oc = ocodec.compress(residual, return_intermediates=True)
            r = res.setdefault("stages_" + mode, {})
            err("y", nchw(s["y"]), oc["_y"], r)
            err("z", nchw(s["z"]), oc["_z"], r)
            err("params_a", nchw(s["params_a"]), oc["_anchor_params"], r)
            err("params_na", nchw(s["params_na"]), oc["_non_anchor_params"], r)
            for k in ("sym_z", "sym_a", "idx_a", "sym_na", "idx_na"):
                got, want = s[k].cpu(), oc["_" + k]
                match = (got == want).float().mean().item()
                maxd = (got - want).abs().max().item()
                r[k] = dict(match=match, maxdiff=maxd)
                print(f"  {k:28s} match {match:.6f}  maxdiff {maxd}")
            if mode == "bf16":
                oc_bf16 = oc

        # ---------------- integer kernels on identical float inputs ----------------
        print("[integer kernels fed the oracle's own y / params (bit-exact contract)]")
        from hyres_b200 import ops
        oy = oc_bf16["_y"].permute(0, 2, 3, 1).contiguous().cuda()
        table = pcodec._scale_table("cuda")
        for name, prm, ps in (("a", "_anchor_params", 0), ("na", "_non_anchor_params", 1)):
            p = oc_bf16[prm].permute(0, 2, 3, 1).contiguous().cuda()
            sym, idx, yq32, _ = ops.gc_symbols(oy, p, ps, table, 0.11)
            ms = (sym.cpu() == oc_bf16["_sym_" + name]).all().item()
            mi = (idx.cpu() == oc_bf16["_idx_" + name]).all().item()
            res["exact_sym_" + name], res["exact_idx_" + name] = ms, mi
            print(f"  pass {name}: symbols exact {ms}  indexes exact {mi}")
        oz = oc_bf16["_z"].permute(0, 2, 3, 1).contiguous().cuda()
        ebp, med = pcodec.engine().eb_params()
        eb = ops.eb_forward(oz, ebp, med, want_symbols=True)
        res["exact_sym_z"] = (eb["symbols"].cpu() == oc_bf16["_sym_z"]).all().item()
        print("  z symbols exact", res["exact_sym_z"])

        # ---------------- strings ----------------
        pc = pcodec.compress(residual.cuda())
        eq = dict(z=pc["strings"][1] == oc_bf16["strings"][1], a=pc["strings"][0][0] == oc_bf16["strings"][0][0],
                  na=pc["strings"][0][1] == oc_bf16["strings"][0][1])
        res["strings_equal_vs_bf16_oracle"] = eq
        print("[strings vs oracle bf16]", eq, "bytes", [sum(len(t) for t in pc["strings"][0][0]), sum(len(t) for t in pc["strings"][0][1]), sum(len(t) for t in pc["strings"][1])])
        # coder alone on the oracle's own symbols -> must be byte-identical
        gc = pcodec.gaussian_conditional
        mine = gc.encode_symbols(oc_bf16["_sym_a"], oc_bf16["_idx_a"])
        res["coder_bytes_equal_on_oracle_symbols"] = mine == oc_bf16["strings"][0][0]
        print("  coder on oracle symbols byte-identical:", res["coder_bytes_equal_on_oracle_symbols"])

        # ---------------- decompress ----------------
        pd = pcodec.decompress(pc["strings"], pc["shape"])
        with O.precision("bf16"):
            od = ocodec.decompress(pc["strings"], pc["shape"])
        print("[decompress of the product's own strings]")
        err("dec x_hat vs oracle bf16", pd["x_hat"].cpu(), od["x_hat"], res.setdefault("decompress", {}))

        # ---------------- forward ----------------
        pf = pcodec(residual.cuda())
        for mode in ("bf16", "fp32"):
            with O.precision(mode):
                of = ocodec(residual)
            print(f"[codec forward vs oracle {mode}]")
            r = res.setdefault("forward_" + mode, {})
            err("x_hat", pf["x_hat"].cpu(), of["x_hat"], r)
            err("lik_y", pf["likelihoods"]["y"].cpu(), of["likelihoods"]["y"], r)
            err("lik_z", pf["likelihoods"]["z"].cpu(), of["likelihoods"]["z"], r)
            npx = a.B * a.H * a.W
            bp = [(-pf["likelihoods"][k].log2().sum() / npx).item() for k in ("y", "z")]
            bo = [(-of["likelihoods"][k].log2().sum() / npx).item() for k in ("y", "z")]
            r["bpp"] = dict(product=bp, oracle=bo)
            print("  bpp product", bp, "oracle", bo)
        # forward/decompress self-consistency: same y_hat => clamp(forward x_hat) == decompress x_hat
        d = (pf["x_hat"].clamp(0, 1) - pd["x_hat"]).abs().max().item()
        res["forward_vs_decompress"] = d
        print("[forward.clamp vs decompress] maxabs", d)

        # ---------------- wrapper ----------------
        pw = pnet(x, jpeg=(jpeg_dec, jpeg_bpp))
        for mode in ("bf16", "fp32"):
            with O.precision(mode):
                ow = onet(x, jpeg=(jpeg_dec, jpeg_bpp))
            print(f"[wrapper forward vs oracle {mode}]")
            r = res.setdefault("wrapper_" + mode, {})
            err("residual", pw["residual"].cpu(), ow["residual"], r)
            err("residual_hat", pw["residual_hat"].cpu(), ow["residual_hat"], r)
            err("x_hat", pw["x_hat"].cpu(), ow["x_hat"], r)
            psnr = lambda t: (-10 * torch.log10((t - x).pow(2).mean())).item()  # noqa: E731
            r["psnr"] = dict(product=psnr(pw["x_hat"].cpu()), oracle=psnr(ow["x_hat"]))
            print("  psnr product", r["psnr"]["product"], "oracle", r["psnr"]["oracle"])
        crit = hyres_b200.RateDistortionLoss(lmbda=0.008)
        lo = crit(pw, x.cuda())
        ocrit = O.RateDistortionLoss(lmbda=0.008)
        with O.precision("bf16"):
            ol = ocrit(onet(x, jpeg=(jpeg_dec, jpeg_bpp)), x)
        res["loss"] = {k: (float(lo[k]), float(ol[k])) for k in ("loss", "bpp_loss", "mse_loss")}
        print("[rd loss product vs oracle bf16]", res["loss"])

        # ---------------- timings ----------------
        xr = residual.cuda()
        for _ in range(3):
            pcodec(xr)
        torch.cuda.synchronize()
        t0 = time.time()
        n = 10
        for _ in range(n):
            pcodec(xr)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / n
        res["codec_forward_ms"] = dt * 1e3
        print(f"[timing] codec forward {dt * 1e3:.2f} ms  ({a.B * a.H * a.W / dt / 1e6:.1f} Mpx/s, eager, wall clock)")
        t0 = time.time()
        with O.precision("fp32"):
            ocodec(residual)
        res["oracle_forward_ms"] = (time.time() - t0) * 1e3
        print(f"[timing] oracle codec forward {res['oracle_forward_ms']:.1f} ms on {torch.get_num_threads()} threads")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1, default=str)


if __name__ == "__main__":
    main()
