"""One small DGCNN_cls forward+backward (the per-rank shapes of tools/check_ddp_equivalence.py) for
`compute-sanitizer --tool memcheck python tools/memcheck_small.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from types import SimpleNamespace
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
dev = torch.device("cuda:0")
torch.manual_seed(3)
args = SimpleNamespace(emb_dims=128, k=12, dropout=0.0)
m = ec.DGCNN_cls(args).to(dev).train()
x = orc.synthetic_xyz(32, 256, seed=5).to(dev)
y = torch.randint(0, 40, (32,), generator=torch.Generator().manual_seed(5)).to(dev)
for b0 in (0, 28):
    loss = ec.cal_loss(m(x[b0:b0 + 4]), y[b0:b0 + 4])
    loss.backward()
torch.cuda.synchronize()
print("memcheck run done, loss", float(loss))
