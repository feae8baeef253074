"""Shared test plumbing: the `gpu` marker, repo paths, golden-fixture loader.

`-m "not gpu"` covers the oracle against the committed golden vectors, the
host-side logic and the C-ABI symbol table; `-m gpu` holds the parity tests
proper, which call the CUDA path through the C-ABI on a B200.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SCALAR_KEYS = ("k", "emb_dim")


def load_golden(name):
    """tests/golden/<name>.npz -> dict of torch tensors (k / emb_dim -> python ints)."""
    out = {}
    with np.load(os.path.join(GOLDEN, name)) as z:
        for key in z.files:
            a = z[key]
            out[key] = a.item() if key in SCALAR_KEYS else torch.from_numpy(a.copy())
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden
