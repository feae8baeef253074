"""``util.py`` as main_cls.py:28 / main_semseg.py:23 import it: ``cal_loss`` (= the reference's
loss.py:4-21 under upstream's name) and ``IOStream`` (the reference's util.py:10-20 logger)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgcnn_pytorch_b200.model import IOStream, cal_loss  # noqa: E402,F401
