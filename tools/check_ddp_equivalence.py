"""torchrun --nproc-per-node N tools/check_ddp_equivalence.py

Sharded training step == single-GPU step on the whole batch: every rank runs DGCNN_cls
(SyncBatchNorm semantics, BatchNorm statistics exchanged between gather and finalize, FlatGradSync) on
its shard of a global batch; rank 0 also runs the same weights in one process on the full batch with
plain BatchNorm.  Loss, averaged gradients and BatchNorm running statistics must agree.

Environment:
  DDP_CHECK_BACKEND  nccl (default: one GPU per rank, statistics over NVLink peer memory) |
                     gloo (ranks may SHARE a GPU -- rank r uses cuda:(r % device_count) -- so the
                     world-size-N protocol is checked on a one-GPU box; statistics travel by gloo)
  DDP_CHECK_B        global batch (default 4 * world);  DDP_CHECK_N, DDP_CHECK_K, DDP_CHECK_EMB
  DDP_CHECK_ARBITER  1: rank 0 also runs the fp64 CPU oracle on the same graphs and reports the
                     deviation of BOTH the sharded and the full-batch GPU gradients from it
  DDP_CHECK_SEED     data seed (default: scan 5, 6, ... for the first batch without a knife-edge, below)
  DDP_CHECK_VERBOSE  1: per-parameter table
  DDP_CHECK_OUT      path of a JSON result file (rank 0)

Knife-edge rule.  LeakyReLU is not differentiable at 0: an activation whose pre-activation is within
fp32 rounding of 0 gets derivative 1 or 0.2 depending on the last bit, in ANY implementation (the
reference included), and the head's BatchNorm bias starts at 0, so its pre-activation IS the
normalised value.  With the historical data seed 5 and B = 32 one element of bn6's output has
|y| ~ 1e-8: single-GPU, sharded and fp64 runs then differ by 1-15 % in head.bn6.bias and everything
upstream of it while every forward value agrees to 2e-6 (profiles/r2_ddp_equivalence.md).  The check
therefore runs on the first seed whose head pre-activations all satisfy |y| > 1e-4 (reported as
"kink margin"); DDP_CHECK_SEED=5 reproduces the knife-edge case.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from types import SimpleNamespace  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dgcnn_pytorch_b200 as ec  # noqa: E402
import edgeconv_oracle as orc  # noqa: E402  (this is a checker, not the product)
from dgcnn_pytorch_b200.dist import FlatGradSync, PeerStatsExchange, shard_range  # noqa: E402

backend = os.environ.get("DDP_CHECK_BACKEND", "nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
ndev = torch.cuda.device_count()
if backend == "nccl":
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
else:
    dev = torch.device("cuda", local % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(os.environ.get("DDP_CHECK_B", 4 * world))
N = int(os.environ.get("DDP_CHECK_N", 256))
k = int(os.environ.get("DDP_CHECK_K", 12))
emb = int(os.environ.get("DDP_CHECK_EMB", 128))
verbose = bool(os.environ.get("DDP_CHECK_VERBOSE"))
args = SimpleNamespace(emb_dims=emb, k=k, dropout=0.0)
torch.manual_seed(3)
ref = ec.DGCNN_cls(args).to(dev).train()                      # identical on every rank (same seed)
sd = {n: v.clone() for n, v in ref.state_dict().items()}
off = int(os.environ.get("DDP_CHECK_OFFSET", "0"))     # clouds [off, off+B) of a larger batch


def batch(seed):
    xs = orc.synthetic_xyz(B + off, N, seed=seed)[off:].contiguous().to(dev)
    ys = torch.randint(0, 40, (B + off,), generator=torch.Generator().manual_seed(seed))[off:].to(dev)
    return xs, ys


def kink_margin(xs):
    """min |pre-activation| over the head's two LeakyReLUs in a full-batch forward (rank 0)."""
    seen = []
    hooks = [m.register_forward_hook(lambda _m, _i, o: seen.append(o.detach().abs().min().item()))
             for m in (ref.head.bn6, ref.head.bn7)]
    with torch.no_grad():
        ref(xs)
    for h in hooks:
        h.remove()
    ref.load_state_dict(sd)          # undo the running-statistics update of this probe
    return min(seen)


model = ec.DGCNN_cls(args).to(dev)
model.load_state_dict(sd)
model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).train()
sync = FlatGradSync(model.parameters(), overlap=os.environ.get("ECB200_GRAD_OVERLAP", "1") == "1"
                    and backend == "nccl")
mode = os.environ.get("ECB200_STATS_EXCHANGE", "peer") if backend == "nccl" else "gloo all-reduce"
if mode == "peer":
    PeerStatsExchange.enable()
b0, b1 = shard_range(B, world, rank)
STRICT, GROSS = 2e-4, 5e-2


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def pick_seed(first):
    """rank 0 scans for a batch whose head pre-activations keep clear of the LeakyReLU kink"""
    seed_t = torch.zeros(2, dtype=torch.float64)
    if rank == 0:
        if os.environ.get("DDP_CHECK_SEED"):
            seed = int(os.environ["DDP_CHECK_SEED"])
            margin = kink_margin(batch(seed)[0])
        else:
            for seed in range(first, first + 64):
                margin = kink_margin(batch(seed)[0])
                if margin > 1e-4:
                    break
        seed_t[0], seed_t[1] = seed, margin
    if backend == "nccl":
        seed_t = seed_t.to(dev)
    dist.broadcast(seed_t, 0)
    return int(seed_t[0].item()), float(seed_t[1].item())


def one_batch(seed, margin):
    """sharded step on every rank + full-batch step on rank 0 -> result dict (rank 0) or None"""
    x, y = batch(seed)
    model.load_state_dict(sd)
    ref.load_state_dict(sd)
    sync.zero()
    # mean over the local shard; shards are equal-sized, so averaging the rank gradients gives the
    # gradient of the global mean loss
    loss = ec.cal_loss(model(x[b0:b1]), y[b0:b1])
    loss.backward()
    sync.average()
    gl = loss.detach().clone()
    dist.all_reduce(gl)
    gl /= world
    if rank != 0:
        return None
    ref.backbone.record_idx = True
    ref.zero_grad(set_to_none=True)
    lr = ec.cal_loss(ref(x), y)
    lr.backward()
    strict = [("loss", abs(gl.item() - lr.item()) / abs(lr.item()), abs(lr.item()))]
    grads = []
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad.abs().max().item() < 1e-6:
            continue
        grads.append((f"grad {n}", rel(p.grad, q.grad), q.grad.abs().max().item()))
    for (n, bp), (_, bq) in zip(model.named_buffers(), ref.named_buffers()):
        if bp.dtype.is_floating_point:
            strict.append((f"buffer {n}", rel(bp, bq), bq.abs().max().item()))
    devs = sorted(strict + grads, key=lambda t: -t[1])
    worst = devs[0]
    if verbose:
        for n, r, sc in devs[:8]:
            print(f"    {n:40s} rel {r:.2e}  (scale {sc:.2e})", flush=True)
    off_grads = sum(1 for _, r, _ in grads if r >= STRICT)
    res = {"seed": seed, "kink_margin": margin, "worst": {"name": worst[0], "rel": worst[1]},
           # every tensor within 2e-4
           "strict_ok": worst[1] < STRICT,
           # loss and BatchNorm buffers within 2e-4, at least half of the gradient tensors within 2e-4 and
           # none grossly off: what one LeakyReLU-kink / tied-max flip inside the network can produce
           "loose_ok": all(r < STRICT for _, r, _ in strict) and off_grads * 2 <= len(grads)
                       and all(r < GROSS for _, r, _ in grads),
           "gradient_tensors_beyond_strict": off_grads, "gradient_tensors": len(grads)}
    if os.environ.get("DDP_CHECK_ARBITER"):
        # fp64 CPU oracle on the graphs the full-batch GPU run used: which side is off?
        o = orc.DGCNNClsOracle(args).double().train()
        o.load_state_dict({n: v.detach().cpu().double() if v.dtype.is_floating_point else v.detach().cpu()
                           for n, v in sd.items()})
        idx_list = [i.long().cpu() for i in ref.backbone.last_idx]
        lo = orc.smoothed_ce_oracle(o(x.cpu().double(), idx_list=idx_list), y.cpu())
        lo.backward()
        worst_s = worst_f = ("", 0.0)
        for (n, p), (_, q), (_, r) in zip(model.named_parameters(), ref.named_parameters(), o.named_parameters()):
            if r.grad.abs().max().item() < 1e-6:
                continue
            ds, df = rel(p.grad.cpu().double(), r.grad), rel(q.grad.cpu().double(), r.grad)
            if ds > worst_s[1]:
                worst_s = (n, ds)
            if df > worst_f[1]:
                worst_f = (n, df)
            if verbose:
                print(f"    vs fp64 oracle  {n:32s} sharded {ds:.2e}   full-batch {df:.2e}", flush=True)
        print(f"  fp64-oracle arbiter: sharded worst {worst_s[1]:.2e} ({worst_s[0]}), full-batch GPU worst "
              f"{worst_f[1]:.2e} ({worst_f[0]})", flush=True)
        res["vs_fp64_oracle"] = {"sharded": worst_s[1], "full_batch": worst_f[1]}
    print(f"  batch seed {seed} (head kink margin {margin:.1e}): worst relative deviation {worst[1]:.2e} at "
          f"{worst[0]}; {off_grads}/{len(grads)} gradient tensors beyond {STRICT}", flush=True)
    return res


# Verdict.  Inside the network ~3 million activations pass a LeakyReLU or a max per step; the sharded
# and the full-batch run differ by ~1e-7 relative in the BatchNorm affine (fp64 sums in another order),
# so in roughly one batch out of four some activation within 1e-7 of the kink (or a tied max) resolves
# the other way and moves a few gradient tensors by 1e-3..1e-2 -- the reference's own autograd has the
# same discontinuity.  Up to three batches are tried: at least one must agree STRICTLY (every tensor
# within 2e-4) and none may be worse than what a single flip explains.
batches = []
first = 5
ntries = 1 if os.environ.get("DDP_CHECK_SEED") else 3
flag = torch.zeros(1, dtype=torch.int32, device=dev if backend == "nccl" else "cpu")
for attempt in range(ntries):
    seed, margin = pick_seed(first)
    first = seed + 1
    r = one_batch(seed, margin)
    if rank == 0:
        batches.append(r)
        flag[0] = 1 if r["strict_ok"] else 0
    dist.broadcast(flag, 0)
    if int(flag.item()):
        break
ok = True
if rank == 0:
    ok = any(b["strict_ok"] for b in batches) and all(b["loose_ok"] for b in batches)
    worst = min(b["worst"]["rel"] for b in batches)
    result = {"world": world, "backend": backend, "stats_exchange": mode, "B": B, "N": N, "k": k, "emb": emb,
              "tolerance": STRICT, "batches": batches, "ok": ok,
              "vs_fp64_oracle": batches[-1].get("vs_fp64_oracle")}
    print(f"ddp equivalence ({world} ranks, backend {backend}, stats exchange = {mode}, B={B} N={N} k={k}): "
          f"{len(batches)} batch(es), best worst-case relative deviation {worst:.2e} -> {'OK' if ok else 'FAIL'}",
          flush=True)
    if os.environ.get("DDP_CHECK_OUT"):
        with open(os.environ["DDP_CHECK_OUT"], "a") as f:
            f.write(json.dumps(result) + "\n")
    flag[0] = 1 if ok else 0
dist.broadcast(flag, 0)
dist.barrier()
torch.cuda.synchronize()
os._exit(0 if int(flag.item()) else 1)
