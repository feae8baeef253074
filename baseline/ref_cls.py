"""Loader for the unmodified reference copy under ``baseline/_ref/`` (see install_ref.py) and the
"DGCNN-cls" network of the headline metric built on it.  MEASUREMENT / TEST INFRASTRUCTURE ONLY:
imported by bench.py's reference legs and by tests, never by the product package.

The reference fork ships only the backbone (models/dgcnn.py:47-103); main_cls.py:25,56 imports a
``DGCNN_cls`` that does not exist there (SURVEY.md §0 trap 2).  Following BASELINE.md §3.4,
DGCNN-cls = the reference's own ``DGCNN`` + upstream DGCNN's classification head (global max + avg
pool -> 2*emb -> 512 -> 256 -> classes); the head below uses the same submodule names as the
product's ``ClsHead`` so state_dicts interchange.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "models", "dgcnn.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_dgcnn_module():
    """The reference's models/dgcnn.py, loaded by path under a private module name."""
    if "_ecb200_ref_dgcnn" in sys.modules:
        return sys.modules["_ecb200_ref_dgcnn"]
    if not available():
        raise FileNotFoundError(f"{REF}/models/dgcnn.py missing: run python baseline/install_ref.py where "
                                "/root/reference is mounted")
    return _load("_ecb200_ref_dgcnn", os.path.join(REF, "models", "dgcnn.py"))


def reference_stack(dgcnn_module, tag: str):
    """The reference's models/layers.py (PositionEmbedding) and models/model_partseg.py (Net) executed
    on top of ``dgcnn_module`` -- either the reference's own models/dgcnn.py or the drop-in -- exactly as
    their ``from models.dgcnn import ...`` lines would bind it.  Returns (layers, model_partseg)."""
    saved = {k: sys.modules.get(k) for k in ("models", "models.dgcnn", "models.layers", "models.model_partseg")}
    pkg = types.ModuleType("models")
    pkg.__path__ = []
    sys.modules["models"] = pkg
    sys.modules["models.dgcnn"] = dgcnn_module
    try:
        layers = _load("models.layers", os.path.join(REF, "models", "layers.py"))
        partseg = _load("models.model_partseg", os.path.join(REF, "models", "model_partseg.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    layers.__name__ = f"_ecb200_ref_layers_{tag}"
    partseg.__name__ = f"_ecb200_ref_partseg_{tag}"
    return layers, partseg


class RefDGCNNCls(nn.Module):
    """reference DGCNN backbone (unmodified, models/dgcnn.py:47-103) + upstream cls head."""

    def __init__(self, args, output_channels: int = 40):
        super().__init__()
        self.backbone = reference_dgcnn_module().DGCNN(args)
        emb = args.emb_dim
        head = nn.Module()
        head.linear1 = nn.Linear(emb * 2, 512, bias=False)
        head.bn6 = nn.BatchNorm1d(512)
        head.dp1 = nn.Dropout(p=float(getattr(args, "dropout", 0.5)))
        head.linear2 = nn.Linear(512, 256)
        head.bn7 = nn.BatchNorm1d(256)
        head.dp2 = nn.Dropout(p=float(getattr(args, "dropout", 0.5)))
        head.linear3 = nn.Linear(256, output_channels)
        self.head = head

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.head
        f = self.backbone(x)
        b = f.size(0)
        f = torch.cat((F.adaptive_max_pool1d(f, 1).view(b, -1), F.adaptive_avg_pool1d(f, 1).view(b, -1)), 1)
        f = h.dp1(F.leaky_relu(h.bn6(h.linear1(f)), negative_slope=0.2))
        f = h.dp2(F.leaky_relu(h.bn7(h.linear2(f)), negative_slope=0.2))
        return h.linear3(f)


def reference_loss(pred: torch.Tensor, gold: torch.Tensor) -> torch.Tensor:
    """The reference's loss.py:4-21 (label-smoothed CE), loaded from the copy."""
    if "_ecb200_ref_loss" not in sys.modules:
        _load("_ecb200_ref_loss", os.path.join(REF, "loss.py"))
    return sys.modules["_ecb200_ref_loss"].cross_entropy(pred, gold)
