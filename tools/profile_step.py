"""Three eager DGCNN-cls train steps at config 1 (B=32, N=1024, k=20) for ncu launch lists / full captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from types import SimpleNamespace
import torch
import dgcnn_pytorch_b200 as ec
from dgcnn_pytorch_b200.synthetic import synthetic_xyz
fwd_only = len(sys.argv) > 1 and sys.argv[1] == "fwd"
dev = torch.device("cuda:0")
torch.manual_seed(1)
model = ec.DGCNN_cls(SimpleNamespace(emb_dim=1024, k=20, dropout=0.5)).to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
x = synthetic_xyz(32, 1024, seed=100).to(dev)
y = torch.randint(0, 40, (32,), generator=torch.Generator().manual_seed(7)).to(dev)
for i in range(3):
    if fwd_only:
        with torch.no_grad():
            model(x)
    else:
        opt.zero_grad(set_to_none=True)
        loss = ec.cal_loss(model(x), y)
        loss.backward()
        opt.step()
torch.cuda.synchronize()
print("ok")
