"""torchrun --nproc-per-node N tools/check_torch_syncbn_head.py

Torch-only control experiment for tools/check_ddp_equivalence.py: the classification head alone
(Linear -> SyncBatchNorm -> LeakyReLU -> Linear -> SyncBatchNorm -> LeakyReLU -> Linear, plain torch
modules, none of this repository's kernels) on sharded random features against the same weights
with plain BatchNorm1d on the whole batch."""
import os
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.backends.cuda.matmul.allow_tf32 = False
per = int(os.environ.get("HEAD_CHECK_LOCAL", "4"))
B, Fin = per * world, 256


def head():
    return nn.Sequential(nn.Linear(Fin, 512, bias=False), nn.BatchNorm1d(512), nn.LeakyReLU(0.2),
                         nn.Linear(512, 256), nn.BatchNorm1d(256), nn.LeakyReLU(0.2), nn.Linear(256, 40))


torch.manual_seed(3)
ref = head().to(dev).train()
sd = {k: v.clone() for k, v in ref.state_dict().items()}
g = torch.Generator().manual_seed(5)
x = torch.randn(B, Fin, generator=g).to(dev)
y = torch.randint(0, 40, (B,), generator=g).to(dev)
model = head().to(dev)
model.load_state_dict(sd)
model = nn.SyncBatchNorm.convert_sync_batchnorm(model).train()
b0, b1 = rank * per, (rank + 1) * per
loss = F.cross_entropy(model(x[b0:b1]), y[b0:b1])
loss.backward()
for p in model.parameters():
    dist.all_reduce(p.grad)
    p.grad /= world
if rank == 0:
    F.cross_entropy(ref(x), y).backward()
    devs = sorted(((((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp_min(1e-12)).item(), n)
                   for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters())), reverse=True)
    print(f"torch-only SyncBN head, {world} ranks x {per} samples: worst relative gradient deviations "
          + ", ".join(f"{n} {d:.1e}" for d, n in devs[:3]), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
