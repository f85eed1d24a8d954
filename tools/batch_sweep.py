"""Step time against batch size, eager and replayed from a CUDA graph (is a step launch-bound? would sub-batches whose
activations stay in the 126 MB L2 pay?).  usage: python tools/batch_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyres_b200  # noqa: E402
from hyres_b200 import synthetic  # noqa: E402

torch.manual_seed(1926)
net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
net.update(force=True)
net = net.cuda().eval()
crit = hyres_b200.RateDistortionLoss(lmbda=0.008)
for B in (16, 8, 4):
    x = synthetic.synthetic_image(B, 512, 768).cuda()
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")

    def step():
        stats.zero_()
        out = net(x, stats=stats)
        return crit(out, x, stats=stats)["loss"]

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    with torch.no_grad():
        for _ in range(3):
            step()
        n = 10 * 16 // B
        eager = timed(step, n)
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = step()
            g.replay()
            torch.cuda.synchronize()
            ref = float(step())
            graph = timed(g.replay, n)
            note = f"graph {graph:.3f} ms/step ({graph * 16 / B:.3f} per 16 images), loss {float(loss):.6f} vs eager {ref:.6f}"
        except Exception as e:  # noqa: BLE001
            note = f"graph capture failed: {type(e).__name__}: {str(e)[:120]}"
    print(f"B={B}: eager {eager:.3f} ms/step ({eager * 16 / B:.3f} per 16 images); {note}")
