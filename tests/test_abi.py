"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol
include/edgeconv_b200.h declares with the arity the ctypes table binds, argument
validation works without touching the device, and the Python mirror keeps the
reference's interface (names, state_dict keys, error behaviour)."""
import ctypes
import os
import re
from types import SimpleNamespace

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "edgeconv_b200.h")


@pytest.fixture(scope="module")
def ec():
    import __graft_entry__ as ge
    ge.build()
    import dgcnn_pytorch_b200 as ec
    return ec


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(ecb200_\w+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_header_symbols_exported_and_bound(ec):
    decls = _declared()
    assert len(decls) >= 20
    lib = ctypes.CDLL(ec._lib.LIB_PATH)
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        if name in ec._lib.PLAIN:
            assert ec._lib.PLAIN[name] == nargs, name
            continue
        assert name in ec._lib.SIGNATURES, f"{name} has no ctypes binding"
        assert len(ec._lib.SIGNATURES[name]) == nargs, name
    assert set(ec._lib.SIGNATURES) | set(ec._lib.PLAIN) == set(decls)


def test_version_and_error_string(ec):
    assert ec._lib.version() == 100
    # argument validation happens before any CUDA call, so it is testable without a GPU
    with pytest.raises(RuntimeError, match="null pointer"):
        ec._lib.call("ecb200_knn", None, None, 1, 3, 8, 2, 1, None, None)
    one = ctypes.c_void_p(16)
    with pytest.raises(RuntimeError, match="out of range"):
        ec._lib.call("ecb200_knn", one, one, 1, 3, 8, 9, 1, one, None)    # k > N, like topk
    with pytest.raises(RuntimeError, match="ECB200_MAX_K"):
        ec._lib.call("ecb200_knn", one, one, 1, 3, 4096, 65, 1, one, None)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        ec._lib.call("ecb200_edge_gather", one, one, one, 1, 8, 2, 6, one, one, None, None, None)


def test_argument_validation_of_the_newer_entry_points(ec):
    """Every entry point checks its arguments before touching the device (no GPU needed)."""
    one = ctypes.c_void_p(16)
    call = ec._lib.call
    with pytest.raises(RuntimeError, match="multiple of 32"):     # tensor-core kNN takes C = 32..128
        call("ecb200_knn_tc", one, one, one, 1, 48, 64, 4, 1, one, one, 1 << 30, None)
    with pytest.raises(RuntimeError, match="exceeds 40"):
        call("ecb200_knn_tc", one, one, one, 1, 64, 256, 41, 1, one, one, 1 << 30, None)
    with pytest.raises(RuntimeError, match="out of range"):      # k > N, the trigger of Tensor.topk
        call("ecb200_knn_tc", one, one, one, 2, 64, 16, 20, 1, one, one, 16, None)
    # survivor lists live in shared memory since round 2: the workspace is a token allocation
    assert ec._lib.load().ecb200_knn_tc_workspace_bytes(2, 256, 20) == 256
    with pytest.raises(RuntimeError, match="come in pairs"):
        call("ecb200_prepare_weights", one, 8, 4, 0, one, one, None, None, None, None)
    with pytest.raises(RuntimeError, match="C in \\{32,64,128\\}"):
        call("ecb200_gemm_dx_tc", one, one, one, one, 1, 48, 64, 64, one, None)
    with pytest.raises(RuntimeError, match="fused mode"):          # dU_in without the hi/lo outputs
        call("ecb200_bwd_dense", one, one, one, one, one, one, 1, 1, 8, 8, None, one, None, None, None)
    with pytest.raises(RuntimeError, match="fused mode"):
        call("ecb200_bwd_scatter", one, one, one, one, one, one, one, one, 1, 8, 2, 8, None, one, None, None, None)
    with pytest.raises(RuntimeError, match="ld_pm"):
        call("ecb200_edge_apply", one, one, one, 0.2, 1, 8, 8, None, one, 4, None)
    with pytest.raises(RuntimeError, match="bad shape"):
        call("ecb200_embed_pool", one, one, one, 0.2, 1, 8, 6, one, one, None)   # E % 4 != 0
    with pytest.raises(RuntimeError, match="outside"):
        call("ecb200_peer_allreduce", one, 5000, one, 0, 2, one, None)
    with pytest.raises(RuntimeError, match="bad rank/world"):
        call("ecb200_peer_allreduce", one, 8, one, 2, 2, one, None)
    assert ec._lib.load().ecb200_peer_buffer_bytes(8) == 2 * 8 * 4160 * 16


def test_argument_validation_of_the_fp16_knn_entry_points(ec):
    one = ctypes.c_void_p(16)
    call = ec._lib.call
    with pytest.raises(RuntimeError, match="must be 64 or 128"):
        call("ecb200_knn_tc_f16", one, one, one, one, one, 1, 96, 256, 8, one, None, None)
    with pytest.raises(RuntimeError, match="out of range"):
        call("ecb200_knn_tc_f16", one, one, one, one, one, 1, 64, 16, 20, one, None, None)
    with pytest.raises(RuntimeError, match="exceeds 40"):
        call("ecb200_knn_tc_xyz", one, one, one, one, 1, 256, 41, one, None, None)
    with pytest.raises(RuntimeError, match="C <= 4"):
        call("ecb200_pack_xyz_f16", one, 1, 5, 64, one, one, one, one, None)
    with pytest.raises(RuntimeError, match="C must be even"):
        call("ecb200_split_f16", one, 1, 7, 64, one, one, one, one, one, one, None, None, None, None)
    with pytest.raises(RuntimeError, match="at most 128"):
        call("ecb200_split_f16", one, 1, 256, 64, one, one, one, one, one, one, None, None, None, None)
    with pytest.raises(RuntimeError, match="come as a pair"):
        call("ecb200_split_f16", one, 1, 8, 64, one, one, one, one, one, one, one, None, None, None)
    with pytest.raises(RuntimeError, match="bad arguments"):
        call("ecb200_absmax", one, 0, one, None)
    with pytest.raises(RuntimeError, match="null pointer"):
        call("ecb200_edge_apply_amax", None, one, one, 0.2, 1, 8, 8, one, None, 8, one, None)
    with pytest.raises(RuntimeError, match="bad shape"):
        call("ecb200_hog_1x1", one, one, 1, 8, 9, one, one, None)          # k > N
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        call("ecb200_hog_1x1", one, one, 1, 8, 4, ctypes.c_void_p(20), one, None)
    assert ec.ops.AMAX_SLOTS == 32
    # kernel choice of knn(): packed fp16 for C % 64 == 0, tf32 halves for the other multiples of 32, the
    # one-K-step kernel for xyz-like inputs, FP32 FMA for everything else
    kind = ec.ops.knn_tc_kind
    assert [kind(64, 1024, 20), kind(128, 1024, 40), kind(96, 512, 20), kind(3, 1024, 20), kind(9, 512, 20),
            kind(64, 512, 41), kind(3, 32, 8), kind(5, 512, 8)] == ["f16", "f16", "tf32", "xyz", "", "", "", ""]


def test_cpu_tensors_are_rejected_not_served(ec):
    x = torch.randn(2, 3, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ec.knn(x, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ec.get_graph_feature(x, k=4)
    net = ec.DGCNN(SimpleNamespace(emb_dim=32, k=4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x)
    with pytest.raises(ValueError):
        ec.get_graph_feature(torch.randn(3, 16), k=4)


def test_module_mirrors_reference_interface(ec):
    import edgeconv_oracle as orc
    args = SimpleNamespace(emb_dim=48, k=6)
    ours, ref = ec.DGCNN(args), orc.DGCNNOracle(args)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert all(sd_o[k].shape == sd_r[k].shape for k in sd_o)
    ours.load_state_dict(sd_r)                       # reference checkpoints load unchanged
    assert ours.k == 6 and ours.emb_dims == 48
    assert ec.DGCNN(SimpleNamespace(emb_dims=16, k=4)).conv5[0].weight.shape == (16, 512, 1, 1)
    # SyncBatchNorm conversion swaps conv{n}[1] in place; the fused forward reads it from there
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(ours)
    assert isinstance(conv.conv1[1], torch.nn.SyncBatchNorm)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dgcnn.pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "edgeconv_oracle" not in text and "/root/reference" not in text.replace(
                    "/root/reference/models/dgcnn.py", ""), f
