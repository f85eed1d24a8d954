"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on a CPU-only container (oracle vs golden vectors, host logic, C-ABI load);
`-m gpu` runs the parity tests proper on a B200 through the C-ABI library.
"""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def state_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        v = sd[k]
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def build_lib():
    """The C-ABI library, built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    import hyres_b200
    return hyres_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import hyres_oracle as O
    return O


@pytest.fixture(scope="session")
def oracle_net(oracle):
    """Seeded random-init wrapper model with 'lively' statistics and CDF tables (CPU oracle)."""
    import torch
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    return oracle.make_model(seed=1926, wrapper=True, lively=True)


@pytest.fixture(scope="session")
def golden_weights_ok(oracle_net):
    """True when the regenerated weights are the ones the golden fixtures were made with."""
    g = load_golden("codec64")
    return state_digest(oracle_net.state_dict()) == bytes(g["state_digest"]).decode()
