"""Throughput pipeline for compress / decompress (models/hyres.py:79-134, models/checkerboard.py:167-240).

One ``compress`` or ``decompress`` call is a chain GPU -> host -> GPU -> host -> GPU: the two rANS passes run on
host threads and the second pass needs the first pass's symbols (context model).  Called back to back, the GPU
idles while the host codes and the host idles while the GPU convolves.  ``CodecPipeline`` keeps several batches
in flight, each on its own worker thread and CUDA stream: the coder calls (ctypes, GIL released) of one batch
overlap the kernels and PCIe copies of the others.  Every batch still goes through the model's public
``compress`` / ``decompress``, so the strings are exactly what single calls return.
"""
import os
import queue
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib, ops


class CodecPipeline:
    def __init__(self, model, workers=3, reuse_host_buffers=False, use_graphs=True):
        """model: ``ResidualJPEGCompression`` (or ``LightWeightCheckerboard``) on a CUDA sm_100 device.
        ``reuse_host_buffers``: decoded images come back in a per-worker ring of pinned buffers (no 35 MB
        ``pin_memory`` per batch); a yielded ``x_hat`` is then valid until the next result is taken from the iterator.
        ``use_graphs``: every worker replays the GPU phases of compress / decompress (the launches between two host
        steps) from CUDA graphs captured on its second batch of a shape -- identical kernels and results, but one
        launch per phase instead of 20 - 80 issued under the interpreter lock that all workers share."""
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("CodecPipeline needs the model on a CUDA sm_100 device (no CPU fallback)")
        self.workers = max(1, int(workers))
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="hyres-codec")
        self._tls = threading.local()
        # one context per image in flight: its CUDA stream and the CUDA graphs captured on it; a job checks one out
        self._contexts = queue.LifoQueue()
        for _ in range(self.workers):
            self._contexts.put({"stream": torch.cuda.Stream(device=self.dev), "graphs": {}})
        self._lock = threading.Lock()
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._wrapper = hasattr(model, "residual_model")
        self.reuse_host_buffers = bool(reuse_host_buffers)
        # build the packed layers (and their weight uploads) once, outside the worker threads
        codec = model.residual_model if self._wrapper else model
        self._codec = codec
        self._graphs_before = codec.graph_phases
        codec.graph_phases = bool(use_graphs)
        # Device coder: every image in flight keeps one coder launch resident for tens of milliseconds (a warp per
        # string, eight warps per block, each block holding a whole SM).  The convolution kernels take a whole SM per
        # CTA with a fixed share of the tiles, so a CTA that had to wait for -- or share -- a coder block's SM would hold
        # up its launch: leave SMs to the coder.  How many does not grow with the images in flight: at the throughput the
        # convolutions allow (~22 ms per 8-tile image) the coder blocks of one image (3 x ~70 ms encoding, 1 x ~110 ms
        # decoding) keep ~13 blocks resident on average, whatever the depth of the pipeline.
        self._reserved_before = None
        if codec.uses_device_coder():
            reserve = int(os.environ.get("HYRES_CODER_SMS", "0")) or min((3 * self.workers + 1) // 2, 16)
            self._reserved_before = _lib.lib().hyres_set_reserved_sms(reserve)
        codec.engine()
        if codec.codec_precision != "bf16":
            codec.precise(codec.codec_precision)
        if self._wrapper:
            model.refine_engine()

    def close(self):
        self._pool.shutdown(wait=True)
        self._codec.graph_phases = self._graphs_before
        if self._reserved_before is not None:
            _lib.lib().hyres_set_reserved_sms(self._reserved_before)
            self._reserved_before = None

    class _Checkout:
        """``with pipe._checkout() as ctx``: a free context, installed as this thread's graph cache for the model."""

        def __init__(self, pipe):
            self.pipe, self.ctx = pipe, None

        def __enter__(self):
            self.ctx = self.pipe._contexts.get()
            self.pipe._codec._phase_tls.graphs = self.ctx["graphs"]
            return self.ctx

        def __exit__(self, *exc):
            self.pipe._codec._phase_tls.graphs = {}
            self.pipe._contexts.put(self.ctx)

    def _checkout(self):
        return CodecPipeline._Checkout(self)

    def warm(self, x, rounds=3):
        """Run ``rounds`` compress + decompress round trips of ``x`` on every context (eager pass, graph capture,
        first replay), so that no capture happens inside a measured or latency-critical region."""
        n = self._contexts.qsize()
        held = [self._contexts.get() for _ in range(n)]
        try:
            for ctx in held:
                self._contexts.put(ctx)
                for _ in range(rounds):
                    self._roundtrip_job(x, False)  # checks out the only free context: this one
                assert self._contexts.get() is ctx
        finally:
            for ctx in held:
                self._contexts.put(ctx)
        torch.cuda.synchronize(self.dev)

    def _host_buffer(self, like):
        if not self.reuse_host_buffers:
            return torch.empty(like.shape, dtype=like.dtype).pin_memory()
        ring = self._tls.__dict__.setdefault("ring", {})
        key = (tuple(like.shape), like.dtype)
        ent = ring.get(key)
        if ent is None:
            # a worker can finish at most `workers` more jobs before the consumer asks for the next result
            ent = ring[key] = [[torch.empty(like.shape, dtype=like.dtype).pin_memory() for _ in range(self.workers + 2)], 0]
        ent[1] = (ent[1] + 1) % (self.workers + 2)
        return ent[0][ent[1]]

    def _count(self, h2d=0, d2h=0):
        with self._lock:
            self.h2d_bytes += h2d
            self.d2h_bytes += d2h

    @staticmethod
    def _stream_bytes(c):
        n = sum(len(s) for grp in (c["strings"][0][0], c["strings"][0][1], c["strings"][1]) for s in grp)
        for b in c.get("jpeg_buffers", ()):
            n += b.getbuffer().nbytes
        return n

    # -- jobs (run on a worker thread, on that worker's stream) --
    @torch.no_grad()
    def _compress_job(self, x):
        torch.cuda.set_device(self.dev)
        with self._checkout() as ctx, torch.cuda.stream(ctx["stream"]):
            if not x.is_cuda:
                self._count(h2d=x.numel() * x.element_size())
                x = x.to(self.dev, non_blocking=True)
            c = self.model.compress(x)
            ops.stream_wait_blocking(self.dev)
        return c

    @torch.no_grad()
    def _decompress_job(self, c, to_host):
        torch.cuda.set_device(self.dev)
        with self._checkout() as ctx, torch.cuda.stream(ctx["stream"]):
            d = self.model.decompress(c) if self._wrapper else self.model.decompress(c["strings"], c["shape"])
            x_hat = d["x_hat"]
            if not to_host and self._codec.graph_phases:
                x_hat = x_hat.clone()  # a graph's output buffer is overwritten by this worker's next batch
            if to_host:
                host = self._host_buffer(x_hat)
                host.copy_(x_hat, non_blocking=True)
                self._count(d2h=x_hat.numel() * x_hat.element_size())
                x_hat = host
            ops.stream_wait_blocking(self.dev)
        return x_hat

    def _roundtrip_job(self, x, to_host):
        c = self._compress_job(x)
        return c, self._decompress_job(c, to_host)

    def _ordered(self, jobs):
        """Submit jobs keeping ``workers`` in flight; yield results in submission order."""
        pending = deque()
        for fn, args in jobs:
            pending.append(self._pool.submit(fn, *args))
            while len(pending) > self.workers:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()

    # -- public --
    def compress(self, batches):
        """batches: iterable of fp32 ``[B,3,H,W]`` tensors (pinned host or device) -> iterator of the dicts
        ``model.compress`` returns, in order."""
        return self._ordered((self._compress_job, (x,)) for x in batches)

    def decompress(self, compressed, to_host=True):
        """compressed: iterable of dicts from ``compress`` -> iterator of ``x_hat`` (pinned host tensors, or device
        tensors with ``to_host=False``), in order."""
        return self._ordered((self._decompress_job, (c, to_host)) for c in compressed)

    def roundtrip(self, batches, to_host=True):
        """compress followed by decompress of every batch -> iterator of (compressed dict, x_hat)."""
        return self._ordered((self._roundtrip_job, (x, to_host)) for x in batches)
