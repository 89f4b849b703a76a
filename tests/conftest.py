import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("MDC_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    g = {}
    for name in ("qweights", "vectors", "h5_weights", "sv_roms"):
        g[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    for name in ("kat", "int_goldens"):
        with open(os.path.join(GOLDEN, name + ".json")) as fh:
            g[name] = json.load(fh)
    return g


@pytest.fixture(scope="session")
def qsets(golden):
    """weight sets A-D as (conv_tab, dense_bias, dense_tabs)"""
    q = golden["qweights"]
    return {k: (q[f"{k}_conv_tab"], q[f"{k}_dense_bias"], q[f"{k}_dense_tabs"]) for k in "ABCD"}


@pytest.fixture(scope="session")
def h5w(golden):
    """float checkpoints: tag -> [conv_k, conv_b, dense_k, dense_b]"""
    h = golden["h5_weights"]
    tags = sorted({k.rsplit("_", 2)[0] for k in h})
    return {t: [h[f"{t}_conv_k"], h[f"{t}_conv_b"], h[f"{t}_dense_k"], h[f"{t}_dense_b"]] for t in tags}


@pytest.fixture(scope="session")
def reference_dir():
    if not os.path.isdir(REFERENCE):
        pytest.skip("reference checkout not present (only in the build container)")
    return REFERENCE


def philox(seed):
    return np.random.Generator(np.random.Philox(seed))
