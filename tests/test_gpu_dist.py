"""Two ranks, one job: each rank runs its shard of a batch through the product forward on the GPU, the four
rate / distortion sums are all-reduced (hyres_b200.dist.reduce_stats -- NCCL when two GPUs are visible, gloo with
both ranks on cuda:0 otherwise) and the global loss must equal a single process's loss over the whole batch
(replaces nn.DataParallel's gather, src/training.py:211-212 + src/losses/rd_loss.py:23-44)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch
import hyres_b200
from hyres_b200 import dist as D, ops, synthetic
use_nccl = torch.cuda.device_count() >= 2
rank, world, local = D.init_from_env(backend="nccl" if use_nccl else "gloo")
dev = torch.device("cuda", local if use_nccl else 0)
torch.cuda.set_device(dev)
torch.manual_seed(1926)
net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
net.update(force=True)
net = net.to(dev).eval()
x = synthetic.synthetic_image(4, 64, 96, seed=5).to(dev)
lo, hi = D.shard_range(4, rank, world)
xs = x[lo:hi].contiguous()
stats = torch.zeros(2, dtype=torch.float64, device=dev)
with torch.no_grad():
    out = net(xs, stats=stats)
    se = torch.zeros(1, dtype=torch.float64, device=dev)
    ops.reduce_sqdiff(out["x_hat"], xs, se)
g = torch.cat([stats, se, torch.tensor([float(xs.shape[0] * 64 * 96)], dtype=torch.float64, device=dev)])
jb = out["jpeg_bpp_loss"].double() * xs.shape[0]  # per-shard JPEG bits / (64*96)
g = torch.cat([g, jb.reshape(1)])
D.reduce_stats(g)
if rank == 0:
    r = D.rd_from_stats(g[:4], 0.008, jpeg_bpp=float(g[4]) / 4)
    print("RESULT " + json.dumps({k: float(v) for k, v in r.items()} | {"backend": "nccl" if use_nccl else "gloo"}))
torch.distributed.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_global_stats_equal_single_process(build_lib, tmp_path):
    import hyres_b200
    from hyres_b200 import ops, synthetic
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    line = [ln for ln in outs[0][0].splitlines() if ln.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    # single process, whole batch
    torch.manual_seed(1926)
    net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
    net.update(force=True)
    net = net.cuda().eval()
    x = synthetic.synthetic_image(4, 64, 96, seed=5).cuda()
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        out = net(x, stats=stats)
        lo = hyres_b200.RateDistortionLoss(lmbda=0.008)(out, x, stats=stats)
    for k in ("y_bpp_loss", "z_bpp_loss", "bpp_loss", "mse_loss", "loss"):
        assert got[k] == pytest.approx(float(lo[k]), rel=2e-6), (k, got, float(lo[k]))
    print(json.dumps(got))


SPATIAL_WORKER = r'''
import json, os, sys, hashlib
sys.path.insert(0, sys.argv[1])
import torch
import hyres_b200
from hyres_b200 import dist as D, spatial, synthetic
use_nccl = torch.cuda.device_count() >= 2
rank, world, local = D.init_from_env(backend="nccl" if use_nccl else "gloo")
dev = torch.device("cuda", local if use_nccl else 0)
torch.cuda.set_device(dev)
from oracle import hyres_oracle as O  # test infrastructure: weights with lively symbol statistics
net = hyres_b200.ResidualJPEGCompression(jpeg_quality=1)
net.load_state_dict(O.make_model(seed=1926, wrapper=True, lively=True).state_dict())
net = net.to(dev).eval()
x = synthetic.synthetic_image(1, 576, 768, seed=8).to(dev)
with torch.no_grad():
    c = spatial.compress_sharded(net, x, rows=1, cols=2, halo=256)
    if rank == 0:
        whole = net.compress(x)
        same = c["strings"] == whole["strings"]
        print("RESULT " + json.dumps({"same": bool(same), "bytes": sum(len(s) for s in c["strings"][0][0])}))
    else:
        assert c is None
torch.distributed.destroy_process_group()
'''


def test_two_rank_spatial_sharding_gives_the_whole_image_strings(build_lib, tmp_path):
    """hyres_b200.spatial.compress_sharded with two ranks (one tile + halo each, integers all-reduced) against the
    single-rank whole-image compress()."""
    script = tmp_path / "spatial_worker.py"
    script.write_text(SPATIAL_WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    line = [ln for ln in outs[0][0].splitlines() if ln.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    assert got["same"] and got["bytes"] > 1000, got
