"""One GPU: in eval mode (running statistics) the clouds are independent, so the gradients of the
batch loss must equal the sum of the gradients of its shards.  Runs the 8 four-cloud shards of the
batch of tools/check_ddp_equivalence.py against the full batch (also 2 shards of 16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from types import SimpleNamespace
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(3)
args = SimpleNamespace(emb_dims=128, k=12, dropout=0.0)
m = ec.DGCNN_cls(args).to(dev).eval()
with torch.no_grad():   # non-trivial running statistics
    for mod in m.modules():
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5)
B = 32
x = orc.synthetic_xyz(B, 256, seed=5).to(dev)
y = torch.randint(0, 40, (B,), generator=torch.Generator().manual_seed(5)).to(dev)


def grads(shard):
    m.zero_grad(set_to_none=True)
    for b0 in range(0, B, shard):
        (ec.cal_loss(m(x[b0:b0 + shard]), y[b0:b0 + shard]) * (shard / B)).backward()
    return {n: p.grad.clone() for n, p in m.named_parameters()}


full = grads(B)
for shard in (16, 4):
    g = grads(shard)
    devs = sorted((((g[n] - full[n]).abs().max() / full[n].abs().max().clamp_min(1e-12)).item(), n)
                  for n in full if full[n].abs().max().item() > 1e-6)
    print(f"shards of {shard}: worst relative deviations " + ", ".join(f"{n} {d:.1e}" for d, n in devs[-4:][::-1]),
          flush=True)
