"""Drop-in for the reference's ``models/dgcnn.py``: put ``<this repo>/dropin`` in front of the
reference checkout on ``sys.path`` (or copy this three-line file over models/dgcnn.py) and
``models/model_partseg.py``, ``models/layers.py``, ``main_partseg_dist.py`` run unchanged on the
fused sm_100a kernels.  Same names and signatures as /root/reference/models/dgcnn.py:6,15,47."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dgcnn_pytorch_b200 import DGCNN, get_graph_feature, knn  # noqa: E402,F401
