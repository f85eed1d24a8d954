#!/usr/bin/env python
"""Headline benchmark of the HyRES residual-codec hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): full ResidualJPEGCompression forward + rate-distortion loss on
a batch of 16 synthetic 768x512 images per GPU (N=128, M=192, random-init weights, seed 1926).
One step = one pass of the hot path over one batch, JPEG stage included on both arms: on the B200 arm
it runs on the device (csrc/jpeg.cu, bit-exact with libjpeg-turbo), on the reference arm it is the
per-image libjpeg-turbo loop the reference runs on the CPU (models/utils/turbo_jpeg_compression.py).

  value : Mpixel/s with inputs resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e   : Mpixel/s through the public API with pinned HOST inputs, H2D copies and the D2H read of
          the loss inside the timed region
  roofline     : the dominant kernel, ru_fused_kernel (tensor bound): its algorithmic FLOPs per launch / its
                 average CUDA-event launch time; `all_tensor_kernels` = canonical FLOPs of the step / summed
                 time of every tcgen05 launch
  cpu_baseline : the CPU oracle (the reference's PyTorch semantics, fp32) on a bounded sample

`--impl reference` times that CPU path alone, on every host core, for the driver's ratio.
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), hbm=float(d["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md, sustained)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

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
                rows.append((ts, float(c[1]), float(c[2]), [n for n, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return None
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use, window = (inside, "timed region") if len(inside) >= 3 else (rows, "timed region + the idle wait before it")
        return {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": max(r[2] for r in use),
                "reasons": sorted({n for r in use for n in r[3]}), "samples": len(use), "window": window}


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


def workload_config(n):
    return {"workload": "BASELINE.json configs[1]: ResidualJPEGCompression (JPEG q=1 stage + residual codec "
                        "N=128 M=192 + MultiScaleRefine) forward + RD loss, batch 16 of 768x512 synthetic images per GPU",
            "batch_per_gpu": BATCH, "height": H, "width": W, "global_batch": BATCH * n, "lambda": LMBDA,
            "sharding": "by image, no data-path collective; 4-double statistics all-reduce",
            "l2": "inputs + activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
            "weights": "random init, seed 1926",
            "launch": "timed steps replayed from a CUDA graph of one step (bench.py --no-graph: eager launches)"}


def codec_cfg3(net, dev, reps=3):
    """BASELINE.json configs[2] on one GPU: compress + decompress (strings out and back, host rANS included) of one
    2048x1408 image as eight 704x512 tiles (tier T-A of SURVEY section 8e).  Extra evidence next to the headline."""
    import torch
    from hyres_b200 import synthetic
    tiles, h, w = 8, 704, 512
    x = synthetic.synthetic_image(tiles, h, w, seed=7)
    xd = x.to(dev)
    with torch.no_grad():
        for _ in range(2):
            d = net.decompress(net.compress(xd))
        t_enc = t_dec = 0.0
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = net.compress(xd)  # device JPEG encoder + residual codec; decompress decodes the JPEG files on the CPU
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            d = net.decompress(c)
            torch.cuda.synchronize()
            t_enc += t1 - t0
            t_dec += time.perf_counter() - t1
    px = tiles * h * w
    nbytes = sum(len(s) for grp in (c["strings"][0][0], c["strings"][0][1], c["strings"][1]) for s in grp)
    return {"workload": "compress + decompress, 8 tiles of 704x512 (one 2048x1408 image), JPEG stage and host rANS included",
            "enc_ms": t_enc / reps * 1e3, "dec_ms": t_dec / reps * 1e3, "encdec_mpixel_per_s": px * reps / (t_enc + t_dec) / 1e6,
            "residual_bpp": 8.0 * nbytes / px, "host_cores": os.cpu_count(), "x_hat_shape": list(d["x_hat"].shape)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
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
        run_reference(args, rank)
        return

    import torch
    import hyres_b200
    from hyres_b200 import _lib, dist as D, ops, synthetic

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
    crit = hyres_b200.RateDistortionLoss(lmbda=LMBDA)

    x_host = synthetic.synthetic_image(BATCH, H, W, seed=1926 + rank).pin_memory()
    x_dev = x_host.to(dev)
    px_step = BATCH * H * W
    stats = torch.zeros(2, dtype=torch.float64, device=dev)

    def step_resident():
        stats.zero_()
        out = net(x_dev, stats=stats)  # JPEG stage (device) + residual codec + refine
        return crit(out, x_dev, stats=stats)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        # The timed steps are replayed from a CUDA graph of one step (the same launches on the same buffers; under
        # capture the two AttentionBlock branches at 1/8 resolution also overlap on two streams).  --no-graph, or a
        # failed capture, times eager launches instead.
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
        launches = lib.hyres_launch_count() - l0  # kernels of ONE step (counted at capture / at the eager call)
        import datetime
        sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
        sampler.start()
        sampler.wait_ready()  # nvidia-smi needs ~0.1 s to deliver its first sample
        barrier()
        t_start = datetime.datetime.now()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                lo = step_resident()
        e1.record()
        barrier()
        t_end = datetime.datetime.now()
        ms_local = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop(t_start, t_end)
        ms = D.max_over_ranks(ms_local, dev)
        loss_val = float(lo["loss"])

        # ---- end to end through the public API with host buffers ----
        # hyres_b200.HostPipeline: every step's x leaves pinned host memory inside the timed region (on a copy
        # stream, under the previous step's kernels), the JPEG stage runs on the device, and every step's loss
        # is read back.
        pipe = hyres_b200.HostPipeline(net, crit)

        def host_batches(n):
            for _ in range(n):
                yield x_host

        for _ in pipe.run(host_batches(3)):
            pass
        barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        e2e_results = list(pipe.run(host_batches(args.steps)))
        torch.cuda.synchronize()
        e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, dev)
        h2d = pipe.h2d_bytes // args.steps
        d2h = pipe.d2h_bytes // args.steps
        e2e_loss = e2e_results[-1]["loss"]

        # ---- roofline of the dominant kernel: per-launch CUDA events around every conv launch ----
        conv_ms, conv_n = None, 0
        if rank == 0:
            ops.ConvLayer.profile_begin()
            step_resident()
            torch.cuda.synchronize()
            conv_ms, conv_n = ops.ConvLayer.profile_end()

    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    peaks = load_peaks()
    value = world * px_step / (ms * 1e-3) / 1e6
    flops_step = 2.0 * MAC_PER_PX_CONV * px_step
    all_tc = flops_step / (conv_ms * 1e-3) / 1e12 if conv_ms else None
    # dominant kernel: ru_fused_kernel (the 14 C=128 ResidualUnit / ResidualBottleneckBlock instances at H/2).
    # algorithmic MACs per position = 128*64 + 576*64 + 64*128; traffic per launch from the
    # ncu --set full capture of the same shape (profiles/r01_ncu_full.md): dram read + write.
    ru = [r for r in ops.ConvLayer.last_profile if r["kind"] == "ru" and r["H"] == H // 2]
    ru_flops = 2.0 * RU_MAC_PER_POS * BATCH * (H // 2) * (W // 2)
    ru_ms = sum(r["ms"] for r in ru) / len(ru) if ru else None
    achieved = ru_flops / (ru_ms * 1e-3) / 1e12 if ru_ms else None
    line = {
        "metric": "hyres_forward_mpixel_per_s", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": world * px_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "ru_fused_kernel (fused 1x1 -> 3x3 -> 1x1 + skip at 16x256x384x128, "
                                                  "%d launches per step)" % len(ru),
                     "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops"] if achieved else None,
                     "traffic": RU_DRAM_BYTES_PER_LAUNCH, "algorithmic_bytes": 2 * BATCH * (H // 2) * (W // 2) * 128 * 2,
                     "ms_per_launch": ru_ms, "peak_source": peaks["source"],
                     "note": "N = 64 tcgen05.mma is bound by shared-memory operand fetch at 2/3 of the dense peak "
                             "(profiles/r01_microbench.md); the kernel is shared-memory-bandwidth bound",
                     "share_of_step": (ru_ms * len(ru)) / ms_local if ru_ms else None,
                     "all_tensor_kernels": {"launches": conv_n, "ms_per_step": conv_ms, "achieved": all_tc,
                                            "frac": all_tc / peaks["tflops"] if all_tc else None,
                                            "share_of_step": conv_ms / ms_local if conv_ms else None},
                     "step_tflops": flops_step / (ms * 1e-3) / 1e12,
                     "step_frac": flops_step / (ms * 1e-3) / 1e12 / peaks["tflops"]},
        "loss": loss_val, "e2e_loss": e2e_loss,
    }
    if world == 1:
        line["codec_cfg3"] = codec_cfg3(net, dev)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        step, px = oracle_step_factory(1, cores)
        step()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": px / dt / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                "sample": f"1 synthetic {W}x{H} image (1/16 of the batch), full forward (CPU JPEG stage "
                                          f"included) + RD loss, fp32 oracle, mean of {n} after 1 warm-up"}
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
