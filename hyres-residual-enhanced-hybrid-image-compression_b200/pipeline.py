"""Host-side batch pipeline: pinned host batches -> device -> forward (+ loss) -> host scalars.

The reference moves every batch with a blocking ``d.to(device)`` and reads the loss with ``.item()``
(src/utils/engine.py:36,92-104), so PCIe time and kernel time add up.  Here the copy of batch k+1 runs on
a copy stream under the kernels of batch k, and the scalars of batch k-1 are read while batch k runs, so a
step costs max(copy, compute) instead of their sum.  Nothing else changes: the model is called through its
public ``forward``; the JPEG stage runs on the device inside it (``csrc/jpeg.cu``) unless a batch carries a
precomputed JPEG result, which is then injected exactly as ``ResidualJPEGCompression.forward`` accepts it.
"""
import torch


class _Slot:
    def __init__(self, dev):
        self.x = self.jd = None
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.result_dev = torch.zeros(3, dtype=torch.float64, device=dev)
        self.result_host = torch.zeros(3, dtype=torch.float64).pin_memory()
        self.stats = torch.zeros(2, dtype=torch.float64, device=dev)
        self.used = False
        self.graph = None       # CUDA graph of forward + loss on this slot's static buffers (device JPEG batches)
        self.graph_shape = None
        self.graph_key = None
        self.eager_runs = 0


class HostPipeline:
    """``run(batches)``: batches is an iterable of ``x_host`` (the JPEG stage then runs on the device) or of
    ``(x_host, jpeg_decoded_host, jpeg_bpp)`` (injected JPEG result), tensors fp32 ``[B,3,H,W]`` in pinned host
    memory.  Yields, in order, one dict per batch with the host floats ``loss``, ``bpp_loss``, ``mse_loss``
    (``src/losses/rd_loss.py:18-44``)."""

    def __init__(self, model, criterion, depth=2, use_graph=True):
        """``use_graph``: after one eager pass per slot, forward + loss of a slot are replayed from a CUDA graph
        captured on the slot's static device buffers (same kernels, same results; the AttentionBlock branches at
        1/8 resolution additionally overlap on two captured streams).  Falls back to eager launches if the capture
        fails or when a batch carries an injected JPEG result."""
        self.model, self.criterion = model, criterion
        self.use_graph = use_graph
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA sm_100 device (no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = [_Slot(self.dev) for _ in range(max(2, depth))]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _unpack(batch):
        if torch.is_tensor(batch):
            return batch, None, None
        x, jd, bpp = batch
        return x, jd, bpp

    def _stage(self, slot, x_host, jd_host):
        if slot.x is None or slot.x.shape != x_host.shape:
            slot.x = torch.empty(x_host.shape, dtype=torch.float32, device=self.dev)
            slot.graph = None  # captured on the old buffer
        if jd_host is not None and (slot.jd is None or slot.jd.shape != jd_host.shape):
            slot.jd = torch.empty(jd_host.shape, dtype=torch.float32, device=self.dev)
        if slot.used:
            self.copy_stream.wait_event(slot.free)  # the kernels of the batch that used this slot are done
        with torch.cuda.stream(self.copy_stream):
            slot.x.copy_(x_host, non_blocking=True)
            if jd_host is not None:
                slot.jd.copy_(jd_host, non_blocking=True)
            slot.ready.record(self.copy_stream)
        self.h2d_bytes += x_host.numel() * 4 + (jd_host.numel() * 4 if jd_host is not None else 0)

    def _body(self, slot, jpeg):
        slot.stats.zero_()
        out = self.model(slot.x, jpeg=jpeg, stats=slot.stats)
        lo = self.criterion(out, slot.x, stats=slot.stats)
        slot.result_dev.copy_(torch.stack([lo["loss"].double(), lo["bpp_loss"].double(), lo["mse_loss"].double()]))

    def _launch(self, slot, jd_host, jpeg_bpp):
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(slot.ready)
        graphable = self.use_graph and jd_host is None and not self.model.training
        if graphable and slot.graph is not None and slot.graph_key != self._weights_key():
            slot.graph = None  # parameters changed since the capture: the packed weights must be refreshed eagerly
        if graphable and slot.graph is not None and slot.graph_shape == tuple(slot.x.shape):
            slot.graph.replay()
        else:
            self._body(slot, None if jd_host is None else (slot.jd, jpeg_bpp))
            slot.eager_runs += 1
            if graphable and slot.graph is None and slot.eager_runs >= 1:
                self._capture(slot)
        slot.free.record(cur)
        slot.result_host.copy_(slot.result_dev, non_blocking=True)
        slot.done.record(cur)
        slot.used = True
        self.d2h_bytes += 24

    def _weights_key(self):
        return sum(p._version for p in self.model.parameters())

    def _capture(self, slot):
        """Capture forward + loss on the slot's buffers; the eager pass just launched has produced this batch's
        result already, the graph serves the following batches."""
        try:
            g = torch.cuda.CUDAGraph()
            keep = slot.result_dev.clone()
            with torch.cuda.graph(g):
                self._body(slot, None)
            slot.result_dev.copy_(keep)  # capture does not execute; keep the eager result for this batch
            slot.graph, slot.graph_shape, slot.graph_key = g, tuple(slot.x.shape), self._weights_key()
        except Exception:  # noqa: BLE001 -- capture is an optimisation; eager launches stay correct
            slot.graph, self.use_graph = None, False

    @staticmethod
    def _collect(slot):
        slot.done.synchronize()
        r = slot.result_host
        return {"loss": float(r[0]), "bpp_loss": float(r[1]), "mse_loss": float(r[2])}

    @torch.no_grad()
    def run(self, batches):
        it = iter(batches)
        n = len(self.slots)
        pending = []  # slots launched, results not yet read
        k = 0
        def pull():
            b = next(it, None)
            return None if b is None else self._unpack(b)

        nxt = pull()
        if nxt is not None:
            self._stage(self.slots[0], nxt[0], nxt[1])
        while nxt is not None:
            cur_batch, slot = nxt, self.slots[k % n]
            nxt = pull()
            self._launch(slot, cur_batch[1], cur_batch[2])
            if nxt is not None:
                # the slot about to be restaged must have had its results read (its pinned buffer is reused)
                tgt = self.slots[(k + 1) % n]
                while pending and pending[0] is tgt:
                    yield self._collect(pending.pop(0))
                self._stage(tgt, nxt[0], nxt[1])
            pending.append(slot)
            while len(pending) > 1:
                yield self._collect(pending.pop(0))
            k += 1
        while pending:
            yield self._collect(pending.pop(0))
