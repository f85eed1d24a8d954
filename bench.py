#!/usr/bin/env python
"""Headline benchmark of the HyRES residual-codec hot path (BASELINE.json metric: enc+dec Mpixel/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload codec|forward]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload `codec` (default; BASELINE.json configs[2]): ResidualJPEGCompression.compress + .decompress of one
2048x1408 CLIC-shape image per GPU per step, as eight 704x512 tiles (tier T-A of SURVEY section 8e) -- JPEG q=1
stage, residual codec with the fp32-equivalent (split-bf16) entropy trunk, rANS strings out (host threads) and
back in, synthesis + MultiScaleRefine.  Steps run through hyres_b200.CodecPipeline (several images in flight on
worker threads / CUDA streams; every image still goes through the public compress / decompress).

  value : Mpixel/s, inputs resident in HBM, x_hat left on the device (the strings do cross the host: that is the path)
  e2e   : Mpixel/s with pinned HOST images in and x_hat back in pinned HOST memory, all copies inside the timed region
  roofline     : the dominant kernel of this workload, conv_tc_kernel on the split-precision layers (tensor bound):
                 algorithmic FLOPs of those layers / their summed CUDA-event launch time (6 tensor-core products per
                 algorithmic MAC are executed: `executed` says what the tensor pipe actually sustained)
  forward      : the round-1 headline kept beside it -- BASELINE.json configs[1], full forward + RD loss, batch 16 of
                 768x512, bf16 trunk -- with the roofline of its dominant kernel (ru_fused_kernel)
  cpu_baseline : the CPU oracle (the reference's PyTorch semantics, fp32) compress + decompress on a bounded sample

Under torchrun every rank runs its own images (weak scaling) and the step's statistics (stream bytes, squared
error, pixels) are all-reduced over NCCL inside the timed step.  `--impl reference` times the CPU path alone, on
every host core, for the driver's ratio.  `--workload forward` prints the forward line instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
OUT = sys.stdout  # where the JSON line goes (main() re-points it at the real stdout before fd 1 is redirected)

H, W, BATCH = 512, 768, 16
LMBDA = 0.008
# canonical algorithmic work, BASELINE.md section 3 (1 MAC = 2 FLOP; masked conv = 12 taps;
# anchor pass of the parameter head K = 384)
MAC_PER_PX_CONV = 483_234  # codec forward 370 624 + MultiScaleRefine 112 610
RU_MAC_PER_POS = 128 * 64 + 9 * 64 * 64 + 64 * 128  # one fused ResidualUnit, per position
RU_DRAM_BYTES_PER_LAUNCH = 761_778_944  # dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_ncu_full.md


CODEC_TILES, CODEC_H, CODEC_W = 8, 704, 512  # one 2048x1408 image as a 2x4 grid of tiles
CODEC_IMAGES = 8  # images per GPU per step (each goes through the public compress / decompress on its own)
# canonical algorithmic work of compress + decompress, SURVEY.md section 8d: encode 215 488 + decode 322 642 MAC/px
MAC_PER_PX_ENCDEC = 538_130


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), burst=float(d["bf16_tflops"]),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, burst=1660.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_ready(self, timeout=2.0):
        """Block (GPU idle) until the first sample has been written, so the timed region is covered from its start."""
        t = time.perf_counter()
        while self.p is not None and time.perf_counter() - t < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.01)

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples whose time stamp lies in [t0, t1] (datetime; the timed
        region); if fewer than three fall inside, of all samples
        -- `window` says which."""
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                util = float(c[9]) if len(c) > 9 and c[9].replace(".", "").isdigit() else None
                rows.append((ts, float(c[1]), float(c[2]), [n for n, v in zip(names, c[5:9]) if v.lower().startswith("active")], util))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return None
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use, window = (inside, "timed region") if len(inside) >= 3 else (rows, "timed region + the idle wait before it")
        utils = [r[4] for r in use if r[4] is not None]
        return {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": max(r[2] for r in use),
                "reasons": sorted({n for r in use for n in r[3]}), "samples": len(use), "window": window,
                "gpu_util_pct": statistics.median(utils) if utils else None}


def oracle_step_factory(sample_images, threads):
    """The reference's CPU path (oracle restatement, fp32, reference semantics) on a bounded sample."""
    import torch
    from oracle import hyres_oracle as O
    torch.set_num_threads(threads)
    net = O.make_model(seed=1926, wrapper=True)
    crit = O.RateDistortionLoss(lmbda=LMBDA)
    x = O.synthetic_image(sample_images, H, W)

    def step():
        with torch.no_grad(), O.precision("fp32"):
            out = net(x)  # JPEG round trip (libjpeg-turbo, CPU) + residual codec + refine, as the reference runs it
            return float(crit(out, x)["loss"])
    return step, sample_images * H * W


def parity_vs_oracle(dev):
    """The checker half of the cpu_baseline leg: the oracle (fp32, the reference's semantics) and the product compress
    the SAME 704x512 tile with the SAME weights; count the integers (symbols / CDF indexes of z and of both
    checkerboard passes) that differ and say whether the three byte strings are identical."""
    import torch
    import hyres_b200
    from oracle import hyres_oracle as O
    onet = O.make_model(seed=1926, wrapper=True, lively=True)
    pnet = hyres_b200.ResidualJPEGCompression()
    pnet.load_state_dict(onet.state_dict())
    pnet = pnet.to(dev).eval()
    x = O.synthetic_image(1, CODEC_H, CODEC_W, seed=7)
    with torch.no_grad():
        jd, _ = onet.jpeg(x)
        res = x - jd
        with O.precision("fp32"):
            oc = onet.residual_model.compress(res, return_intermediates=True)
        s = pnet.residual_model.encode_symbols(res.to(dev))
        c = pnet.residual_model.compress(res.to(dev))
    out = {"tile": f"{CODEC_W}x{CODEC_H}", "trunk": pnet.residual_model.codec_precision, "elements": {}, "end_to_end": {}}
    for k in ("sym_z", "sym_a", "sym_na", "idx_a", "idx_na"):
        want = oc["_" + k].int()
        out["elements"][k] = int(want.numel())
        out["end_to_end"][k] = int((s[k].cpu() != want).sum())
    out["strings_identical"] = {"z": c["strings"][1] == oc["strings"][1], "anchor": c["strings"][0][0] == oc["strings"][0][0],
                                "non_anchor": c["strings"][0][1] == oc["strings"][0][1]}
    # stage by stage: every oracle stage is fed the product's integers of the stage before, so each stream is compared
    # on identical stage inputs (tools/check_precise.py)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_precise
    rep = check_precise.symbol_report(pnet.residual_model, onet.residual_model, O, res)
    out["stagewise"] = {k: v["mismatches"] for k, v in rep["stagewise"].items()}
    out["stagewise_not_on_a_tie"] = rep["stagewise_unexplained"]
    out["y_rel_err"], out["params_rel_err_stagewise"] = rep["y_rel_err"], rep["params_na_rel_err_stagewise"]
    out["note"] = ("mismatching integers between the product and the fp32 CPU oracle, same weights and residual.  "
                   "end_to_end: the oracle's own compress(); one hyper-latent on a rounding tie changes the parameters "
                   "of its whole receptive field and one anchor tie the context of its neighbours, so these counts "
                   "include such consequences.  stagewise: identical stage inputs; stagewise_not_on_a_tie counts the "
                   "mismatches that do NOT sit within 2e-4 of a rounding tie / scale-table edge of the oracle's own value")
    return out


def oracle_codec_step_factory(tiles, threads):
    """The reference's CPU compress + decompress (models/hyres.py:79-134 through the oracle restatement, fp32,
    its C rANS coder, libjpeg-turbo for the JPEG stage) on `tiles` 704x512 tiles."""
    import torch
    from oracle import hyres_oracle as O
    torch.set_num_threads(threads)
    net = O.make_model(seed=1926, wrapper=True)
    x = O.synthetic_image(tiles, CODEC_H, CODEC_W, seed=7)

    def step():
        with torch.no_grad(), O.precision("fp32"):
            c = net.compress(x)
            d = net.decompress(c)
            return float(d["x_hat"].mean())
    return step, tiles * CODEC_H * CODEC_W


def run_reference_codec(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step, px = oracle_codec_step_factory(1, cores)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = px / dt / 1e6
    sample = (f"1 synthetic {CODEC_W}x{CODEC_H} tile per step (of the {CODEC_TILES}-tile image), compress + decompress "
              "(CPU JPEG stage, fp32 convolutions on all cores, single-threaded C rANS per string as in the reference)")
    print(file=OUT, flush=True, *[json.dumps({
        "impl": "reference", "metric": "hyres_encdec_mpixel_per_s", "value": v, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": codec_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })])


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (here: the oracle port,
    because compressai is not installable -- DESIGN.md section 3), all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_images = 1
    step, px = oracle_step_factory(sample_images, cores)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = px / dt / 1e6
    sample = f"{sample_images} synthetic {W}x{H} image per step (of the {BATCH}-image batch), full forward (CPU JPEG stage included) + RD loss, fp32"
    print(file=OUT, flush=True, *[json.dumps({
        "impl": "reference", "metric": "hyres_forward_mpixel_per_s", "value": v, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })])


def codec_config(n, workers=None):
    c = {"workload": "BASELINE.json configs[2]: ResidualJPEGCompression.compress + .decompress (JPEG q=1 stage, residual "
                     "codec N=128 M=192 with checkerboard two-pass symbols + CDF indexes, rANS strings, MultiScaleRefine) of "
                     "%d 2048x1408 synthetic images per GPU per step, each as eight 704x512 tiles (tier T-A) through its own "
                     "compress / decompress call" % CODEC_IMAGES,
         "images_per_gpu_per_step": CODEC_IMAGES, "tiles_per_image": CODEC_TILES, "height": CODEC_H, "width": CODEC_W,
         "global_tiles": CODEC_IMAGES * CODEC_TILES * n,
         "trunk_precision": "fp32h2 (fp32 activations carried as two IEEE half parts, 3 tensor-core products per MAC: "
                            "fp32-equivalent) for g_a/h_a/h_s/context/parameter head; bf16 for g_s/refine",
         "sharding": "by image (tile set) per rank, no data-path collective; per-step NCCL all-reduce of 4 doubles "
                     "(stream bytes, squared error, pixels, images) inside the timed step",
         "l2": "inputs + activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
         "weights": "random init, seed 1926",
         "timing": "host wall clock between device-wide synchronizes (max over ranks): a step alternates kernels on "
                   "several streams with host rANS threads, so stream-local CUDA events cannot bracket it"}
    if workers is not None:
        c["images_in_flight"] = workers
    return c


def workload_config(n):
    return {"workload": "BASELINE.json configs[1]: ResidualJPEGCompression (JPEG q=1 stage + residual codec "
                        "N=128 M=192 + MultiScaleRefine) forward + RD loss, batch 16 of 768x512 synthetic images per GPU",
            "batch_per_gpu": BATCH, "height": H, "width": W, "global_batch": BATCH * n, "lambda": LMBDA,
            "sharding": "by image, no data-path collective; the 4-double statistics all-reduce (sum log2 lik_y, sum log2 "
                        "lik_z, squared error, pixels) runs over NCCL after every step, inside the timed region",
            "l2": "inputs + activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
            "weights": "random init, seed 1926",
            "launch": "timed steps replayed from a CUDA graph of one step (bench.py --no-graph: eager launches)"}


def bench_codec(args, net, dev, rank, world, lib, peaks):
    """compress + decompress of one 8-tile image per step through CodecPipeline.  Returns the JSON line (rank 0)."""
    import datetime
    import torch
    import hyres_b200
    from hyres_b200 import dist as D, ops, synthetic

    net.residual_model.coder = args.coder
    device_coder = net.residual_model.uses_device_coder()
    # images in flight: the host coder is fed by ~8; a coder warp runs a string's chain ~10x slower than a host core, so
    # the device coder needs more images resident to keep the convolution kernels busy (they only cost memory)
    workers = args.in_flight if args.in_flight > 0 else (16 if device_coder else 8)
    tiles, h, w = CODEC_TILES, CODEC_H, CODEC_W
    px_img = tiles * h * w
    px_step = CODEC_IMAGES * px_img
    # a few distinct images per rank, cycled (the JPEG stage and the coder see different data every step)
    hosts = [synthetic.synthetic_image(tiles, h, w, seed=7 + 17 * rank + k).pin_memory() for k in range(4)]
    devs = [t.to(dev) for t in hosts]
    pipe = hyres_b200.CodecPipeline(net, workers=workers, reuse_host_buffers=True)
    stats = torch.zeros(4, dtype=torch.float64, device=dev)  # stream bytes, squared error, pixels, images

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(n, resident):
        """n steps of CODEC_IMAGES images; every result is consumed (its stream size and reconstruction error enter the
        step statistics, which are all-reduced across ranks once per step: the bpp / MSE a multi-GPU job reports)."""
        src = devs if resident else hosts
        last = None
        acc = torch.zeros(4, dtype=torch.float64, device=dev)
        images = n * CODEC_IMAGES
        for k, (c, x_hat) in enumerate(pipe.roundtrip((src[i % len(src)] for i in range(images)), to_host=not resident)):
            nbytes = pipe._stream_bytes(c)
            ref = devs[k % len(devs)]
            xh = x_hat if resident else x_hat.to(dev, non_blocking=True)
            se = (xh - ref).double().pow(2).sum()
            acc += torch.stack([torch.tensor(float(nbytes), dtype=torch.float64, device=dev), se,
                                torch.tensor(float(px_img), dtype=torch.float64, device=dev),
                                torch.tensor(1.0, dtype=torch.float64, device=dev)])
            if (k + 1) % CODEC_IMAGES == 0:
                stats.copy_(acc)
                acc.zero_()
                D.reduce_stats(stats)
                last = stats.clone()
        torch.cuda.synchronize()
        return last

    with torch.no_grad():
        pipe.warm(devs[0])  # every context: eager pass, CUDA-graph capture of its GPU phases, first replay
        run(max(args.warmup, (workers + CODEC_IMAGES) // CODEC_IMAGES), True)
        l0 = lib.hyres_launch_count()
        sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else dev.index)
        sampler.start()
        sampler.wait_ready()
        barrier()
        t_start = datetime.datetime.now()
        cpu0 = os.times()
        t0 = time.perf_counter()
        g = run(args.steps, True)
        barrier()
        ms_local = (time.perf_counter() - t0) * 1e3 / args.steps
        cpu1 = os.times()
        host_cpu_ms = ((cpu1.user - cpu0.user) + (cpu1.system - cpu0.system)) * 1e3 / args.steps
        t_end = datetime.datetime.now()
        clocks = sampler.stop(t_start, t_end)
        launches = lib.hyres_launch_count() - l0
        ms = D.max_over_ranks(ms_local, dev)

        # ---- end to end: pinned host images in, x_hat back in pinned host memory ----
        run((workers + CODEC_IMAGES) // CODEC_IMAGES, False)
        barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        ge = run(args.steps, False)
        barrier()
        e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, dev)
        h2d, d2h = pipe.h2d_bytes // args.steps, pipe.d2h_bytes // args.steps

        # ---- one un-pipelined compress / decompress: latency of the single public calls + per-launch conv times ----
        prof = None
        if rank == 0:
            x = devs[0]
            net.decompress(net.compress(x))  # this thread's stream: warm the allocator before timing single calls
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = net.compress(x)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            net.decompress(c)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            le0 = lib.hyres_launch_count()
            ops.ConvLayer.profile_begin()
            c = net.compress(x)
            net.decompress(c)
            conv_ms, conv_n = ops.ConvLayer.profile_end()
            eager_launches = lib.hyres_launch_count() - le0  # this thread launches eagerly: every kernel is counted
            rows = ops.ConvLayer.last_profile
            split = [r for r in rows if r.get("nsplit", 1) > 1]
            # the single most expensive layer geometry of the step and its algorithmic HBM traffic (fp32 tensors of the
            # reference's own formulation: input, output, and the skip / gate operands the layer adds)
            geo = {}
            for r in split:
                key = (r["cin"], r["cout"], r["k"], r["stride"], r["kind"], r["OH"], r["OW"], r["split_mode"])
                gg = geo.setdefault(key, dict(ms=0.0, n=0, bytes=0.0, macs=0.0, r=r))
                extra = {0: 0, 1: 1, 2: 2, 3: 1, 4: 1}.get(r["split_mode"], 0)
                gg["ms"] += r["ms"]
                gg["n"] += 1
                gg["bytes"] += 4.0 * r["B"] * (r["H"] * r["W"] * r["cin"] + r["OH"] * r["OW"] * r["cout"] * (1 + extra))
                gg["macs"] += r["alg_macs"]
            top = max(geo.values(), key=lambda e: e["ms"])
            split_bytes = sum(e["bytes"] for e in geo.values())
            prof = dict(enc_ms=(t1 - t0) * 1e3, dec_ms=(t2 - t1) * 1e3, conv_ms=conv_ms, conv_n=conv_n,
                        split_ms=sum(r["ms"] for r in split), split_n=len(split),
                        split_flops=2.0 * sum(r["alg_macs"] for r in split),
                        split_executed=2.0 * sum(r["alg_macs"] * r["products"] for r in split),
                        split_bytes=split_bytes,
                        eager_launches=int(eager_launches),
                        top=dict(ms=top["ms"], n=top["n"], bytes=top["bytes"], flops=2.0 * top["macs"],
                                 what="%dx%d %d->%d%s at %dx%dx%d" % (top["r"]["k"], top["r"]["k"], top["r"]["cin"], top["r"]["cout"],
                                                                    {1: " + skip", 2: " + gate", 3: " + GDN"}.get(top["r"]["split_mode"], ""),
                                                                    top["r"]["B"], top["r"]["OH"], top["r"]["OW"])))
    pipe.close()
    if rank != 0:
        return None
    value = world * px_step / (ms * 1e-3) / 1e6
    gbytes, gse, gpx, gimg = [float(v) for v in g.tolist()]
    achieved = prof["split_flops"] / (prof["split_ms"] * 1e-3) / 1e12
    executed = prof["split_executed"] / (prof["split_ms"] * 1e-3) / 1e12
    hbm_gbs = prof["split_bytes"] / (prof["split_ms"] * 1e-3) / 1e9
    line = {
        "metric": "hyres_encdec_mpixel_per_s", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 / bf16 tensor-core products, fp32 accumulation; the entropy-critical trunk carries "
                                      "fp32 activations as 2 half parts (fp32-equivalent); symbols int32",
        "data": "synthetic", "config": dict(codec_config(world, workers), coder=(
            "device: every rANS string coded by one warp beside the convolution kernels (csrc/rans_dev.cu), no host work "
            "per symbol" if device_coder else "host: rANS strings coded on the box's cores (csrc/rans.cpp)") +
            " [--coder %s, %d host cores for this process]" % (args.coder, hyres_b200.coder.host_cores_per_process())),
        "e2e": {"value": world * px_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "h2d / d2h count the image in and the reconstruction out; " + (
                    "the byte strings cross PCIe on top of that in both arms (resident and e2e): they are the product"
                    if device_coder else
                    "the symbol / index tensors cross PCIe on top of that in both arms (resident and e2e), because the "
                    "entropy coder runs on the host")},
        "gpu_launches": prof["eager_launches"] * CODEC_IMAGES * args.steps,
        "gpu_launches_per_step": prof["eager_launches"] * CODEC_IMAGES,
        "launch": "the library's kernels of one compress + decompress (counted on an eager pass); in the timed steps the "
                  "launches of a GPU phase are replayed from CUDA graphs, one graph per phase and image in "
                  "flight (%d counted eager launches per step remain: JPEG stage, coder, copies)" % (int(launches) // max(1, args.steps)),
        "host": {"cores": os.cpu_count(), "cpu_ms_per_step": host_cpu_ms,
                 "note": "process CPU time (user + system, all threads: range coder, JPEG file decode, launches) per step "
                         "on rank 0; divided by the core count it is the floor the host puts under ms_per_step"},
        "clocks": clocks,
        "global_stats": {"bpp": 8.0 * gbytes / gpx, "mse_255": gse / (gpx * 3) * 255 ** 2, "images": gimg,
                         "note": "last step, all-reduced over ranks (NCCL) inside the timed loop"},
        "single_call_latency": {"compress_ms": prof["enc_ms"], "decompress_ms": prof["dec_ms"],
                                "note": "one un-pipelined public compress() / decompress() of one 8-tile image"},
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel on the %d split-precision (%s) layers of one compress + "
                                                  "decompress" % (prof["split_n"], net.residual_model.codec_precision),
                     "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                     "traffic": None, "executed": executed,
                     "executed_frac": executed / peaks["tflops"],
                     "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm"],
                             "note": "the same launches against the HBM roof: algorithmic fp32 bytes of every layer (input + "
                                     "output + added operands, as the reference formulates it) / their summed CUDA-event "
                                     "time; the kernel is about as far from this roof as from the tensor one: the fused fp32 epilogue "
                                     "(instruction issue) bounds most launches, profiles/r02_ncu_conv_tc.md"},
                     "ms_per_step": prof["split_ms"] * CODEC_IMAGES, "peak_source": peaks["source"],
                     "note": "achieved counts ALGORITHMIC FLOPs (one MAC per weight tap); each is executed as 3 half-precision "
                             "tensor-core products (6 bf16 ones in the three GDN gamma layers) so that symbols equal the "
                             "fp32 reference's; `executed` is what the tensor pipe sustains.  Most of these layers are "
                             "HBM-bound at this tile size (1x1 / 3x3 layers of 64-128 channels carrying fp32 activations)",
                     "dominant_layer": {"layer": prof["top"]["what"], "launches_per_step": prof["top"]["n"] * CODEC_IMAGES,
                                        "ms_per_step": prof["top"]["ms"] * CODEC_IMAGES, "bound": "hbm",
                                        "achieved": prof["top"]["bytes"] / (prof["top"]["ms"] * 1e-3) / 1e9, "peak": peaks["hbm"],
                                        "unit": "GB/s", "frac": prof["top"]["bytes"] / (prof["top"]["ms"] * 1e-3) / 1e9 / peaks["hbm"],
                                        "tflops_algorithmic": prof["top"]["flops"] / (prof["top"]["ms"] * 1e-3) / 1e12,
                                        "traffic": 1_237_650_000 if prof["top"]["what"].startswith("1x1 64->128 + skip") else None,
                                        "note": "the layer geometry with the largest share of the step's GPU time; achieved = "
                                                "algorithmic fp32 bytes (input + output + added operands) / CUDA-event time; "
                                                "traffic = dram bytes of one launch, profiles/r02_ncu_conv_tc.md"},
                     "all_tensor_kernels": {"launches": prof["conv_n"] * CODEC_IMAGES,
                                            "ms_per_step": prof["conv_ms"] * CODEC_IMAGES,
                                            "achieved": 2.0 * MAC_PER_PX_ENCDEC * px_img / (prof["conv_ms"] * 1e-3) / 1e12},
                     "step_tflops": 2.0 * MAC_PER_PX_ENCDEC * px_step / (ms * 1e-3) / 1e12},
    }
    return line


def bench_forward(args, net, dev, rank, world, lib, peaks, steps):
    """BASELINE.json configs[1]: forward + RD loss, batch 16 of 768x512 (the round-1 headline).  Returns a dict."""
    import datetime
    import torch
    import hyres_b200
    from hyres_b200 import dist as D, ops, synthetic

    crit = hyres_b200.RateDistortionLoss(lmbda=LMBDA)
    x_host = synthetic.synthetic_image(BATCH, H, W, seed=1926 + rank).pin_memory()
    x_dev = x_host.to(dev)
    px_step = BATCH * H * W
    stats = torch.zeros(2, dtype=torch.float64, device=dev)
    gstats = torch.zeros(4, dtype=torch.float64, device=dev)
    se = torch.zeros(1, dtype=torch.float64, device=dev)

    def step_resident():
        stats.zero_()
        out = net(x_dev, stats=stats)  # JPEG stage (device) + residual codec + refine
        lo = crit(out, x_dev, stats=stats)
        # the four sums a multi-GPU job reduces: sum log2 lik_y, sum log2 lik_z, squared error, pixels
        se.zero_()
        ops.reduce_sqdiff(out["x_hat"], x_dev, se)
        gstats[0:2].copy_(stats)
        gstats[2:3].copy_(se)
        gstats[3].fill_(float(px_step))
        return lo

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        graph, lo = None, None
        l0 = lib.hyres_launch_count()
        if not args.no_graph:
            try:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    lo = step_resident()
                for _ in range(2):
                    graph.replay()
            except Exception as exc:  # noqa: BLE001
                print(f"bench: CUDA graph capture failed ({type(exc).__name__}: {exc}); timing eager launches", file=sys.stderr)
                graph = None
        if graph is None:
            l0 = lib.hyres_launch_count()
            lo = step_resident()
        launches = lib.hyres_launch_count() - l0
        red = gstats.clone()
        sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else dev.index)
        sampler.start()
        sampler.wait_ready()
        barrier()
        t_start = datetime.datetime.now()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            if graph is not None:
                graph.replay()
            else:
                lo = step_resident()
            red.copy_(gstats)
            D.reduce_stats(red)  # NCCL all-reduce of the step's four sums, on the same stream, every step
        e1.record()
        barrier()
        t_end = datetime.datetime.now()
        ms_local = e0.elapsed_time(e1) / steps
        clocks = sampler.stop(t_start, t_end)
        ms = D.max_over_ranks(ms_local, dev)
        loss_val = float(lo["loss"])
        glob = D.rd_from_stats(red, LMBDA, jpeg_bpp=0.0)

        pipe = hyres_b200.HostPipeline(net, crit)

        def host_batches(n):
            for _ in range(n):
                yield x_host

        for _ in pipe.run(host_batches(3)):
            pass
        barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        e2e_results = list(pipe.run(host_batches(steps)))
        torch.cuda.synchronize()
        e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps, dev)
        h2d = pipe.h2d_bytes // steps
        d2h = pipe.d2h_bytes // steps
        e2e_loss = e2e_results[-1]["loss"]

        conv_ms, conv_n = None, 0
        if rank == 0:
            ops.ConvLayer.profile_begin()
            step_resident()
            torch.cuda.synchronize()
            conv_ms, conv_n = ops.ConvLayer.profile_end()
    if rank != 0:
        return None
    value = world * px_step / (ms * 1e-3) / 1e6
    flops_step = 2.0 * MAC_PER_PX_CONV * px_step
    all_tc = flops_step / (conv_ms * 1e-3) / 1e12 if conv_ms else None
    ru = [r for r in ops.ConvLayer.last_profile if r["kind"] == "ru" and r["H"] == H // 2]
    ru_flops = 2.0 * RU_MAC_PER_POS * BATCH * (H // 2) * (W // 2)
    ru_ms = sum(r["ms"] for r in ru) / len(ru) if ru else None
    achieved = ru_flops / (ru_ms * 1e-3) / 1e12 if ru_ms else None
    return {
        "metric": "hyres_forward_mpixel_per_s", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": world * px_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "forward + RD loss: the step's result is the loss (24 bytes back); x_hat stays on the device"},
        "gpu_launches": int(launches) * steps,
        "gpu_launches_per_step": int(launches),
        "clocks": clocks,
        "global_stats": {"bpp_residual": float(glob["residual_bpp_loss"]), "mse_255": float(glob["mse_loss"]),
                         "note": "all-reduced over ranks (NCCL) after every timed step"},
        "roofline": {"bound": "tensor", "kernel": "ru_fused_kernel (fused 1x1 -> 3x3 -> 1x1 + skip at 16x256x384x128, "
                                                  "%d launches per step)" % len(ru),
                     "achieved": achieved, "peak": peaks["burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["burst"] if achieved else None,
                     "frac_of_sustained": achieved / peaks["tflops"] if achieved else None,
                     "traffic": RU_DRAM_BYTES_PER_LAUNCH, "algorithmic_bytes": 2 * BATCH * (H // 2) * (W // 2) * 128 * 2,
                     "ms_per_launch": ru_ms, "peak_source": peaks["source"] + "; per-launch timing -> burst peak",
                     "share_of_step": (ru_ms * len(ru)) / ms_local if ru_ms else None,
                     "all_tensor_kernels": {"launches": conv_n, "ms_per_step": conv_ms, "achieved": all_tc,
                                            "frac": all_tc / peaks["tflops"] if all_tc else None,
                                            "share_of_step": conv_ms / ms_local if conv_ms else None},
                     "step_tflops": flops_step / (ms * 1e-3) / 1e12,
                     "step_frac": flops_step / (ms * 1e-3) / 1e12 / peaks["tflops"]},
        "loss": loss_val, "e2e_loss": e2e_loss,
    }


TRAIN_BATCH, TRAIN_H, TRAIN_W = 16, 256, 256


def train_config(n):
    return {"workload": "BASELINE.json configs[4]: training step (forward + backward, lambda = 0.008 RD loss, gradient "
                        "clipping, Adam + auxiliary Adam) on a batch of 16 synthetic 256x256 crops per GPU, noise quantiser",
            "batch_per_gpu": TRAIN_BATCH, "height": TRAIN_H, "width": TRAIN_W, "global_batch": TRAIN_BATCH * n,
            "lambda": LMBDA, "precision": "bf16 tensor-core convolutions (forward, data gradient, weight gradient), "
                                          "bf16 activations / gradients between layers, fp32 master weights and Adam",
            "parallelism": f"dp{n}: one process per GPU, gradients all-reduced over NCCL in 8 MB buckets launched from "
                           "autograd hooks during backward (replaces nn.DataParallel, src/training.py:211-212)",
            "weights": "random init, seed 1926"}


def bench_train(args, dev, rank, world, lib, peaks, steps):
    """One optimisation step per timed step (src/utils/engine.py:29-90).  Returns a dict (rank 0) or None."""
    import torch
    import hyres_b200
    from hyres_b200 import dist as D, synthetic, train as T

    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.to(dev)
    trainer = T.Trainer(net, lmbda=LMBDA, lr=1e-4, aux_lr=1e-3, clip_max_norm=1.0, capturable=not args.no_graph)
    x_host = synthetic.synthetic_image(TRAIN_BATCH, TRAIN_H, TRAIN_W, seed=11 + rank).pin_memory()
    x_dev = x_host.to(dev)
    px_step = TRAIN_BATCH * TRAIN_H * TRAIN_W

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r = trainer.step(x_dev)
    graphed = False
    l0 = lib.hyres_launch_count()
    if not args.no_graph:
        try:
            trainer.capture(x_dev, warmup=1)
            graphed = True
        except Exception as exc:  # noqa: BLE001
            print(f"bench: training-step capture failed ({type(exc).__name__}: {exc}); timing eager launches", file=sys.stderr)
    launches_captured = lib.hyres_launch_count() - l0
    l0 = lib.hyres_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = trainer.step(x_dev)
    e1.record()
    barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    launches = (lib.hyres_launch_count() - l0) // steps
    if graphed:  # replayed launches are not counted by the library: one captured step (after its eager warm-up step)
        launches = launches_captured // 2
    loss = float(r["loss"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = trainer.step(x_host.to(dev, non_blocking=True))
        loss_e2e = float(r["loss"])  # the step's result is read back every step
    torch.cuda.synchronize()
    e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps, dev)
    if rank != 0:
        return None
    flops = 3 * 2.0 * MAC_PER_PX_CONV * px_step  # forward + data gradient + weight gradient of every convolution
    return {"metric": "hyres_train_mpixel_per_s", "value": world * px_step / (ms * 1e-3) / 1e6, "unit": "Mpixel/s",
            "n_gpus": world, "steps": steps, "ms_per_step": ms, "scaling": "weak", "config": train_config(world),
            "e2e": {"value": world * px_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches_per_step": int(launches),
            "launch": "whole step replayed from a CUDA graph" if graphed else "eager launches",
            "weight_gradient": "tcgen05 (csrc/wgrad.cu)" if T.wgrad_native_active() else "ATen convolution_backward (cuDNN)",
            "step_tflops": flops / (ms * 1e-3) / 1e12, "step_frac": flops / (ms * 1e-3) / 1e12 / peaks["tflops"],
            "loss": loss, "loss_e2e": loss_e2e}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="codec", choices=["codec", "forward", "train"])
    ap.add_argument("--coder", default=os.environ.get("HYRES_CODER", "auto"), choices=["auto", "host", "device"],
                    help="codec: where the rANS strings are coded (auto: on the device when this process has fewer than "
                         "16 host cores to itself)")
    ap.add_argument("--in-flight", type=int, default=int(os.environ.get("HYRES_CODEC_IN_FLIGHT", "0")),
                    help="images in flight in the codec pipeline (worker threads / CUDA streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forward", action="store_true", help="codec workload: skip the forward sub-benchmark")
    ap.add_argument("--no-train", action="store_true", help="codec workload: skip the training-step sub-benchmark")
    ap.add_argument("--no-graph", action="store_true", help="forward: time eager launches instead of CUDA-graph replays")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its version banner there
    # when the first communicator is created) is sent to stderr
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        (run_reference_codec if args.workload == "codec" else run_reference)(args, rank)
        return

    import torch
    import hyres_b200
    from hyres_b200 import _lib, dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA sm_100 device: the hot path has no CPU fallback")
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.lib()
    _lib.check(lib.hyres_device_check(local), "hyres_device_check")

    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.to(dev).eval()
    peaks = load_peaks()

    if args.workload == "forward":
        line = bench_forward(args, net, dev, rank, world, lib, peaks, args.steps)
    elif args.workload == "train":
        line = bench_train(args, dev, rank, world, lib, peaks, args.steps)
        if line is not None:
            line.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "bf16",
                         "data": "synthetic"})
    else:
        line = bench_codec(args, net, dev, rank, world, lib, peaks)
        if not args.no_forward:
            fwd = bench_forward(args, net, dev, rank, world, lib, peaks, max(10, args.steps))
            if line is not None:
                line["forward"] = {k: fwd[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "e2e",
                                                       "gpu_launches_per_step", "roofline", "global_stats", "loss",
                                                       "config")}
        if not args.no_train:
            del net
            torch.cuda.empty_cache()
            tr = bench_train(args, dev, rank, world, lib, peaks, max(5, min(args.steps, 10)))
            if line is not None:
                line["train"] = tr
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if args.workload == "train":
            step, px, what = None, 0, ""
        elif args.workload == "forward":
            step, px = oracle_step_factory(1, cores)
            what = f"1 synthetic {W}x{H} image (1/16 of the batch), full forward (CPU JPEG stage included) + RD loss"
        else:
            step, px = oracle_codec_step_factory(1, cores)
            what = (f"1 synthetic {CODEC_W}x{CODEC_H} tile (1/8 of the image), compress + decompress (CPU JPEG stage, "
                    "single-threaded C rANS per string as in the reference)")
        if step is not None:
            step()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                step()
            dt = (time.perf_counter() - t0) / n
            line["cpu_baseline"] = {"value": px / dt / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                    "sample": what + f", fp32 oracle, mean of {n} after 1 warm-up"}
            if args.workload == "codec":
                line["cpu_baseline"]["parity"] = parity_vs_oracle(dev)
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
