"""ctypes binding of libedgeconv_b200.so (the C ABI in include/edgeconv_b200.h).

This is the stub a maintainer of the reference would add next to
models/dgcnn.py.  There is no fallback of any kind: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libedgeconv_b200.so")

P, I, F, D, Z, LL = c_void_p, c_int, c_float, c_double, c_size_t, c_longlong

# name -> argument types, in the order of include/edgeconv_b200.h (all return int)
SIGNATURES = {
    "ecb200_sqnorms": (P, I, I, I, P, P),
    "ecb200_knn": (P, P, I, I, I, I, I, P, P),
    "ecb200_split_tf32": (P, I, I, I, P, P, P, P),
    "ecb200_knn_tc": (P, P, P, I, I, I, I, I, P, P, Z, P),
    "ecb200_debug_tc_scores": (P, P, P, I, I, I, P, P),
    "ecb200_absmax": (P, LL, P, P),
    "ecb200_split_f16": (P, I, I, I, P, P, P, P, P, P, P, P, P, P),
    "ecb200_knn_tc_f16": (P, P, P, P, P, I, I, I, I, P, P, P),
    "ecb200_pack_xyz_f16": (P, I, I, I, P, P, P, P, P),
    "ecb200_knn_tc_xyz": (P, P, P, P, I, I, I, P, P, P),
    "ecb200_debug_tc_scores_f16": (P, P, I, I, I, P, P),
    "ecb200_debug_tc_timeline": (P, P, P, I, I, I, I, P, P, P, P),
    "ecb200_split_rows_tf32": (P, LL, P, P, P),
    "ecb200_point_gemm_tc": (P, P, P, P, LL, I, I, P, P),
    "ecb200_graph_feature": (P, P, I, I, I, I, I, P, P),
    "ecb200_graph_feature_bwd": (P, P, I, I, I, I, I, P, P),
    "ecb200_pack_weight": (P, I, I, I, P, P),
    "ecb200_prepare_weights": (P, I, I, I, P, P, P, P, P, P),
    "ecb200_point_gemm": (P, P, I, I, I, I, P, P),
    "ecb200_edge_gather": (P, P, P, I, I, I, I, P, P, P, P, P),
    "ecb200_bn_finalize": (P, P, P, P, P, I, F, I, P, P, P, P, P),
    "ecb200_bn_update_running": (P, I, F, P, P, P, P),
    "ecb200_edge_apply": (P, P, P, F, I, I, I, P, P, LL, P),
    "ecb200_edge_apply_amax": (P, P, P, F, I, I, I, P, P, LL, P, P),
    "ecb200_bwd_prep": (P, P, LL, P, P, P, P, P, F, I, I, I, P, P, P),
    "ecb200_bwd_finalize": (P, P, P, P, P, I, I, P, P, P, P, P),
    "ecb200_reverse_graph": (P, I, I, I, P, P, P, P),
    "ecb200_bwd_dense": (P, P, P, P, P, P, I, I, I, I, P, P, P, P, P),
    "ecb200_bwd_scatter": (P, P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P),
    "ecb200_gemm_dx": (P, P, I, I, I, I, P, P),
    "ecb200_gemm_dw": (P, P, I, I, I, I, P, P),
    "ecb200_unpack_weight_grad": (P, I, I, I, P, P),
    "ecb200_transpose_split_tf32": (P, I, I, P, P, P),
    "ecb200_gemm_dx_tc": (P, P, P, P, I, I, I, I, P, P),
    "ecb200_gemm_dw_tc": (P, P, P, P, LL, I, I, P, P),
    "ecb200_peer_allreduce": (P, I, P, I, I, P, P),
    "ecb200_peer_allreduce_bn_finalize": (P, I, P, I, I, P, P, P, F, P, P, P, P, P),
    "ecb200_peer_allreduce_bwd_finalize": (P, P, I, P, I, I, P, P, P, P, P, P, P, P, P),
    "ecb200_colstats": (P, LL, I, P, P),
    "ecb200_embed_pool": (P, P, P, F, I, I, I, P, P, P),
    "ecb200_embed_pool_bwd_stats": (P, P, P, P, P, P, P, F, I, I, I, P, P),
    "ecb200_embed_gemm": (P, P, P, P, LL, I, I, P, P, P),
    "ecb200_two_conv_fwd": (P, P, P, P, F, P, P, P, I, I, I, I, I, P, P, P),
    "ecb200_embed_pool_bwd_dz": (P, P, P, P, P, P, P, P, F, I, I, I, P, P),
    "ecb200_hog_1x1": (P, P, I, I, I, P, P, P),
    "ecb200_rows_bn_bwd_stats": (P, P, P, P, I, I, P, P, P, P),
    "ecb200_rows_bn_bwd_dx": (P, P, P, P, P, P, P, I, I, P, P),
}

# kernels each entry point enqueues (memsets are not counted)
KERNELS_PER_CALL = {name: 1 for name in SIGNATURES}
KERNELS_PER_CALL["ecb200_hog_1x1"] = 2
KERNELS_PER_CALL["ecb200_reverse_graph"] = 1      # one CTA per cloud (3 kernels only when N counters exceed shared memory)

# entry points that do not return an error code
PLAIN = {"ecb200_version": 0, "ecb200_last_error": 0, "ecb200_knn_tc_workspace_bytes": 3,
         "ecb200_peer_buffer_bytes": 1}

_lib = None
launch_count = 0          # kernels of this library enqueued so far (bench.py's gpu_launches)
_event_hook = None        # optional callable(name) -> context manager, set by bench.py


def set_event_hook(hook) -> None:
    """bench.py installs a hook that brackets chosen entry points with CUDA events on the
    launching stream (per-kernel durations for the roofline figure)."""
    global _event_hook
    _event_hook = hook


def load() -> ctypes.CDLL:
    """Load the shared library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python dgcnn.pytorch_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU or PyTorch fallback for the "
            "EdgeConv path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = list(argtypes)
        fn.restype = c_int
    lib.ecb200_knn_tc_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.ecb200_knn_tc_workspace_bytes.restype = c_size_t
    lib.ecb200_peer_buffer_bytes.argtypes = [c_int]
    lib.ecb200_peer_buffer_bytes.restype = c_size_t
    lib.ecb200_version.argtypes = []
    lib.ecb200_version.restype = c_int
    lib.ecb200_last_error.argtypes = []
    lib.ecb200_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an entry point; non-zero return -> RuntimeError(ecb200_last_error())."""
    global launch_count
    lib = load()
    if _event_hook is not None:
        with _event_hook(name):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    launch_count += KERNELS_PER_CALL[name]
    if rc != 0:
        msg = lib.ecb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")


def version() -> int:
    return load().ecb200_version()
