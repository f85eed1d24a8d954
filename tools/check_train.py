"""GPU check of the training step (hyres_b200.train) against the CPU oracle's autograd.

``grad_report``: one forward + backward of the RD loss on the same weights, image, injected JPEG stage and injected
noise tensors; per parameter the cosine similarity and norm ratio of the product's gradient (bf16 tensor-core
convolutions, fp32 master weights) against the oracle's fp32 gradient.
``loss_curves``: N optimisation steps (Adam, gradient clipping, auxiliary optimiser) on both sides.

usage: python tools/check_train.py [--steps 20] [--out gpurun_out/train_check.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_noise(seed, B, M, N, h, w, noisequant):
    """The U(-1/2, 1/2) tensors in the order the reference's forward draws them (EntropyBottleneck, [anchor,
    non-anchor quantisers,] GaussianConditional), from the CPU generator the oracle itself will use."""
    import torch
    torch.manual_seed(seed)
    shapes = [("z", (N, 1, B * (h // 4) * (w // 4)))]
    if noisequant:
        shapes += [("y_a", (B, M, h, w)), ("y_na", (B, M, h, w))]
    shapes += [("y_lik", (B, M, h, w))]
    return {tag: torch.empty(shape).uniform_(-0.5, 0.5) for tag, shape in shapes}


def product_noise_fn(noise, dev):
    def fn(shape, tag):
        t = noise[tag]
        if tag != "z":
            t = t.permute(0, 2, 3, 1)
        t = t.to(dev).contiguous()
        assert tuple(t.shape) == tuple(shape), (tag, t.shape, shape)
        return t
    return fn


def oracle_step_loss(onet, O, x, jpeg, lmbda, noisequant, seed):
    import torch
    torch.manual_seed(seed)  # the oracle draws its noise from the global CPU generator, in forward order
    with O.precision("fp32"):
        out = onet(x, noisequant=noisequant, jpeg=jpeg)
        crit = O.RateDistortionLoss(lmbda=lmbda)(out, x)
    return out, crit


def grad_report(pnet, onet, O, x, lmbda=0.008, noisequant=True, seed=77):
    import torch
    from hyres_b200 import train as T
    codec = onet.residual_model
    B, _, H, W = x.shape
    jpeg = onet.jpeg(x)
    noise = make_noise(seed, B, codec.M, codec.N, H // 8, W // 8, noisequant)
    onet.train()
    onet.zero_grad()
    _, oc = oracle_step_loss(onet, O, x, jpeg, lmbda, noisequant, seed)
    oc["loss"].backward()
    pnet.train()
    pnet.zero_grad()
    g = T.TrainGraph(pnet)
    out = g.forward(x.cuda(), noisequant=noisequant, jpeg=jpeg, noise_fn=product_noise_fn(noise, "cuda"), training=True)
    pc = T.rd_loss(out, x.cuda(), lmbda)
    pc["loss"].backward()
    rep = {"losses": {k: (float(pc[k]), float(oc[k])) for k in ("loss", "bpp_loss", "y_bpp_loss", "z_bpp_loss", "mse_loss")},
           "params": {}}
    onamed = dict(onet.named_parameters())
    cos_w, n_w = 0.0, 0.0
    for name, p in pnet.named_parameters():
        og = onamed[name].grad
        if og is None and p.grad is None:
            continue
        if og is None or p.grad is None:
            rep["params"][name] = {"missing": "oracle" if og is None else "product"}
            continue
        a, b = p.grad.detach().float().cpu().reshape(-1), og.reshape(-1)
        na, nb = float(a.norm()), float(b.norm())
        cos = float((a @ b) / (na * nb + 1e-30))
        rep["params"][name] = {"cos": cos, "norm_ratio": na / (nb + 1e-30), "oracle_norm": nb, "numel": a.numel()}
        cos_w += cos * nb
        n_w += nb
    rep["weighted_cos"] = cos_w / max(n_w, 1e-30)
    cs = [v["cos"] for v in rep["params"].values() if "cos" in v and v["oracle_norm"] > 1e-6]
    rep["min_cos"] = min(cs)
    rep["median_cos"] = sorted(cs)[len(cs) // 2]
    rep["missing"] = [k for k, v in rep["params"].items() if "missing" in v]
    return rep


def loss_curves(pnet, onet, O, x, steps=20, lmbda=0.008, lr=1e-4, aux_lr=1e-3, noisequant=True):
    import torch
    from hyres_b200 import train as T
    codec = onet.residual_model
    B, _, H, W = x.shape
    jpeg = onet.jpeg(x)
    trainer = T.Trainer(pnet, lmbda=lmbda, lr=lr, aux_lr=aux_lr, clip_max_norm=1.0)
    named = dict(onet.named_parameters())
    main = [named[n] for n in sorted(named) if not n.endswith(".quantiles")]
    aux = [named[n] for n in sorted(named) if n.endswith(".quantiles")]
    opt = torch.optim.Adam(main, lr=lr, betas=(0.9, 0.999))
    aopt = torch.optim.Adam(aux, lr=aux_lr, betas=(0.9, 0.999))
    onet.train()
    got, want = [], []
    for k in range(steps):
        seed = 1000 + k
        noise = make_noise(seed, B, codec.M, codec.N, H // 8, W // 8, noisequant)
        r = trainer.step(x.cuda(), noisequant=noisequant, jpeg=jpeg, noise_fn=product_noise_fn(noise, "cuda"))
        got.append({k2: float(v) for k2, v in r.items()})
        opt.zero_grad()
        aopt.zero_grad()
        _, oc = oracle_step_loss(onet, O, x, jpeg, lmbda, noisequant, seed)
        oc["loss"].backward()
        torch.nn.utils.clip_grad_norm_(onet.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        al = onet.aux_loss()
        al.backward()
        aopt.step()
        aopt.zero_grad()
        want.append({k2: float(v) for k2, v in oc.items()} | {"aux_loss": float(al)})
    return got, want


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_check.json"))
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    import hyres_b200
    from oracle import hyres_oracle as O
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))

    def fresh():
        onet = O.make_model(seed=1926, wrapper=True, lively=True)
        pnet = hyres_b200.ResidualJPEGCompression()
        pnet.load_state_dict(onet.state_dict())
        return onet, pnet.cuda()

    x = O.synthetic_image(2, 64, 64, seed=3)
    out = {}
    for nq in (True, False):
        onet, pnet = fresh()
        r = grad_report(pnet, onet, O, x, noisequant=nq)
        worst = sorted(((v["cos"], k) for k, v in r["params"].items() if "cos" in v and v["oracle_norm"] > 1e-6))[:8]
        print(json.dumps({"noisequant": nq, "losses": r["losses"], "weighted_cos": r["weighted_cos"],
                          "median_cos": r["median_cos"], "min_cos": r["min_cos"], "missing": r["missing"],
                          "worst": worst}), flush=True)
        out["grad_noisequant_%s" % nq] = r
    onet, pnet = fresh()
    got, want = loss_curves(pnet, onet, O, x, steps=args.steps)
    for k, (g, w) in enumerate(zip(got, want)):
        print(k, "loss %.4f / %.4f  bpp %.4f / %.4f  mse %.2f / %.2f  aux %.2f / %.2f" % (
            g["loss"], w["loss"], g["bpp_loss"], w["bpp_loss"], g["mse_loss"], w["mse_loss"], g["aux_loss"], w["aux_loss"]),
            flush=True)
    out["curve"] = {"product": got, "oracle": want}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
