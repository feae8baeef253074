"""edgeconv-b200: the DGCNN EdgeConv hot path as hand-written sm_100a kernels.

The package directory is literally ``dgcnn.pytorch_b200`` (not an importable
dotted name); import it through the ``dgcnn_pytorch_b200`` shim at the repo
root:  ``import dgcnn_pytorch_b200 as ec``.
"""
from . import _lib, ops
from .dgcnn import DGCNN, edgeconv_block, get_graph_feature, knn, two_conv_edge_block
from .hog import compute_hog_1x1
from .model import ClsHead, DGCNN_cls, DGCNN_semseg, IOStream, PointNet, cal_loss
from .ops import edgeconv
from .runtime import GraphedTrainStep

__all__ = ["compute_hog_1x1", "DGCNN", "GraphedTrainStep", "DGCNN_cls", "DGCNN_semseg", "PointNet", "ClsHead", "IOStream", "cal_loss",
           "edgeconv", "edgeconv_block", "two_conv_edge_block", "get_graph_feature", "knn", "ops", "_lib"]
