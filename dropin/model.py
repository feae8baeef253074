"""``model.py`` as main_cls.py:25 and main_semseg.py:20 import it (the reference fork deleted the
file, SURVEY.md §0 trap 2): upstream DGCNN's model classes on top of the fused EdgeConv path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgcnn_pytorch_b200.model import DGCNN_cls, DGCNN_semseg, PointNet  # noqa: E402,F401
from dgcnn_pytorch_b200.dgcnn import get_graph_feature, knn  # noqa: E402,F401
