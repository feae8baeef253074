"""torchrun --nproc-per-node N tools/check_ddp_equivalence.py

Sharded training step == single-GPU step on the whole batch: every rank runs DGCNN_cls
(SyncBatchNorm semantics, BatchNorm statistics over NVLink peer memory, FlatGradSync) on its
shard of a global batch; rank 0 also runs the same weights in one process on the full batch with
plain BatchNorm.  Loss, averaged gradients and BatchNorm running statistics must agree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from types import SimpleNamespace
import torch
import torch.distributed as dist
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
from dgcnn_pytorch_b200.dist import FlatGradSync, PeerStatsExchange, shard_range

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, N, k = int(os.environ.get("DDP_CHECK_B", 4 * world)), 256, 12
args = SimpleNamespace(emb_dims=128, k=k, dropout=0.0)
torch.manual_seed(3)
ref = ec.DGCNN_cls(args).to(dev).train()                      # identical on every rank (same seed)
sd = {n: v.clone() for n, v in ref.state_dict().items()}
off = int(os.environ.get("DDP_CHECK_OFFSET", "0"))     # clouds [off, off+B) of a larger batch
x = orc.synthetic_xyz(B + off, N, seed=5)[off:].contiguous().to(dev)
y = torch.randint(0, 40, (B + off,), generator=torch.Generator().manual_seed(5))[off:].to(dev)

model = ec.DGCNN_cls(args).to(dev)
model.load_state_dict(sd)
model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).train()
sync = FlatGradSync(model.parameters(), overlap=os.environ.get("ECB200_GRAD_OVERLAP", "1") == "1")
mode = os.environ.get("ECB200_STATS_EXCHANGE", "peer")
if mode == "peer":
    PeerStatsExchange.enable()
b0, b1 = shard_range(B, world, rank)
sync.zero()
# mean over the local shard; shards are equal-sized, so averaging the rank gradients gives the
# gradient of the global mean loss
logits = model(x[b0:b1])
loss = ec.cal_loss(logits, y[b0:b1])
loss.backward()
if os.environ.get("DDP_CHECK_VERBOSE"):
    # per-rank forward check against the full-batch reference, and the all-reduce against a
    # manual mean of the gathered per-rank gradients
    with torch.no_grad():
        sdr = {n: v.clone() for n, v in ref.state_dict().items()}
        full = ref(x)
        ref.load_state_dict(sdr)
    fdev = ((logits.detach() - full[b0:b1]).abs().max() / full.abs().max()).reshape(1)
    fall = [torch.empty_like(fdev) for _ in range(world)]
    dist.all_gather(fall, fdev)
    gath = [torch.empty_like(sync.flat) for _ in range(world)]
    dist.all_gather(gath, sync.flat)
    manual = torch.stack(gath).mean(0)
sync.average()
if os.environ.get("DDP_CHECK_VERBOSE") and rank == 0:
    print("    per-rank logits deviation vs full-batch reference:", [f"{float(v):.1e}" for v in fall], flush=True)
    print(f"    all-reduce vs manual mean of gathered grads: {((sync.flat - manual).abs().max() / manual.abs().max()).item():.2e}",
          flush=True)
gl = loss.detach().clone()
dist.all_reduce(gl)
gl /= world
ok = True
if rank == 0:
    ref.zero_grad(set_to_none=True)
    lr = ec.cal_loss(ref(x), y)
    lr.backward()
    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    devs = [("loss", abs(gl.item() - lr.item()) / abs(lr.item()), abs(lr.item()))]
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad.abs().max().item() < 1e-6:
            continue
        devs.append((f"grad {n}", rel(p.grad, q.grad), q.grad.abs().max().item()))
    for (n, bp), (_, bq) in zip(model.named_buffers(), ref.named_buffers()):
        if bp.dtype.is_floating_point:
            devs.append((f"buffer {n}", rel(bp, bq), bq.abs().max().item()))
    devs.sort(key=lambda t: -t[1])
    worst = devs[0]
    if os.environ.get("DDP_CHECK_VERBOSE"):
        for n, r, sc in devs[:8]:
            print(f"    {n:40s} rel {r:.2e}  (scale {sc:.2e})", flush=True)
    ok = worst[1] < 2e-4
    print(f"ddp equivalence ({world} ranks, stats exchange = {mode}, overlap = "
          f"{os.environ.get('ECB200_GRAD_OVERLAP', '1')}): worst relative deviation "
          f"{worst[1]:.2e} at {worst[0]} -> {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0 if ok else 1)
