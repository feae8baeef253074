"""compute_hog_1x1 of the reference's part-segmentation model (models/model_partseg.py:15-92) on the
device (second half).

The reference copies the gathered neighbourhoods to the host, runs ``np.linalg.svd`` on B*N matrices of
k x 3 and copies the result back -- every forward, stalling the GPU.  ``compute_hog_1x1`` here keeps
everything on the device: the kNN graph comes from the same one-entry cache ``DGCNN.forward`` and
``PositionEmbedding`` fill (ops.knn_cached), the principal directions from a 3x3 eigen-solver kernel,
the orientation histograms from a second kernel (csrc/hog.cu).

Drop-in use (the reference's ``Net`` looks the function up in its own module)::

    import models.model_partseg as ps
    from dgcnn_pytorch_b200.hog import compute_hog_1x1
    ps.compute_hog_1x1 = compute_hog_1x1

One deliberate difference: the sign of every principal direction, which the reference inherits from
LAPACK's SVD and which moves the zenith bin, is fixed to v_z >= 0.  No independent implementation can
reproduce LAPACK's per-matrix sign; weights trained on the reference's features see different (equally
valid) features here.  The test suite's CPU restatement of the reference with ``canonical_sign=True``
is the reference with that one change and is what the tests compare with.
"""
from __future__ import annotations

from ctypes import c_void_p

import torch

from . import _lib, ops


@torch.no_grad()
def compute_hog_1x1(x: torch.Tensor, k: int, use_cpu: bool = False, idx: torch.Tensor = None) -> torch.Tensor:
    """x [B,3,N] -> [B,N,18] float32 on x's device (no gradient, as in the reference, whose SVD leaves
    autograd through numpy).  ``use_cpu`` is accepted for signature compatibility and ignored: there is
    no CPU path.  ``idx``: optional precomputed knn(x, k) [B,N,k]."""
    if x.dim() != 3 or x.shape[1] != 3:
        raise ValueError(f"compute_hog_1x1 expects x of shape [B, 3, N], got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("edgeconv_b200 runs on CUDA tensors only: there is no CPU fallback "
                           f"(got a tensor on {x.device}); the CPU path is the reference itself")
    x = x.detach().float().contiguous()
    B, _, N = x.shape
    if idx is None:
        idx32 = ops.knn_cached(x, int(k))
    else:
        idx32 = ops.check_neighbour_indices(idx, B, N)
    k = idx32.shape[-1]
    with torch.cuda.device(x.device):
        ws = torch.empty(N, 4, device=x.device, dtype=torch.float32)
        hist = torch.empty(B, N, 18, device=x.device, dtype=torch.float32)
        _lib.call("ecb200_hog_1x1", c_void_p(x.data_ptr()), c_void_p(idx32.data_ptr()), B, N, k,
                  c_void_p(ws.data_ptr()), c_void_p(hist.data_ptr()),
                  c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    return hist
