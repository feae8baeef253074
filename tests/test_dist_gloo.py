"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the
BatchNorm-statistics exchange protocol of SURVEY.md §8(e).  The statistics themselves are
produced on the GPU by ecb200_edge_gather / ecb200_bwd_prep; here they are restated with
the oracle so that the PROTOCOL (what is summed, and that the sums reproduce single-process
BatchNorm over the whole batch) is checked without a GPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _edge_values(x, w, k, idx):
    import edgeconv_oracle as orc
    gf = orc.graph_feature_oracle(x, k=k, idx=idx)                    # [B,2C,N,k]
    return torch.nn.functional.conv2d(gf, w.view(w.shape[0], -1, 1, 1))   # e = W.[x_j; x_i]


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import dgcnn_pytorch_b200 as ec
    from dgcnn_pytorch_b200 import dist as ecd
    import edgeconv_oracle as orc
    torch.set_num_threads(1)
    r, _, w = ecd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # group handles are stable and non-zero (0 means "no exchange")
    h = ec.ops.register_group(None)
    assert h > 0 and ec.ops.register_group(None) == h

    B, C, N, k, Co = 5, 4, 48, 6, 8                     # 5 clouds over 2 ranks: 3 + 2
    x = orc.synthetic_features(B, C, N, seed=3).double()
    wgt = torch.randn(Co, 2 * C, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    gamma = torch.linspace(-1.0, 1.5, Co, dtype=torch.float64)
    beta = torch.linspace(0.3, -0.3, Co, dtype=torch.float64)
    gout = torch.randn(B, Co, N, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    lo, hi = ecd.shard_range(B, world, rank)
    xs = x[lo:hi]
    idx = orc.knn_oracle(xs, k)
    e = _edge_values(xs, wgt, k, idx)                   # this rank's edges [b,Co,N,k]
    # forward exchange: [sum e | sum e^2 | count]
    stats = torch.cat([e.sum((0, 2, 3)), (e * e).sum((0, 2, 3)), torch.tensor([float(e[:, 0].numel())],
                                                                              dtype=torch.float64)])
    ecd.allreduce_stats(stats)
    cnt = stats[-1]
    mean = stats[:Co] / cnt
    var = stats[Co:2 * Co] / cnt - mean * mean
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    a = gamma * invstd
    b = beta - a * mean
    sel = torch.where(gamma.view(1, -1, 1) >= 0, e.max(-1)[0], e.min(-1)[0])
    y = torch.nn.functional.leaky_relu(a.view(1, -1, 1) * sel + b.view(1, -1, 1), 0.2)
    # backward exchange: [sum g | sum g * xhat]
    g = gout[lo:hi] * torch.where(a.view(1, -1, 1) * sel + b.view(1, -1, 1) > 0, 1.0, 0.2)
    xhat = (sel - mean.view(1, -1, 1)) * invstd.view(1, -1, 1)
    bst = torch.cat([g.sum((0, 2)), (g * xhat).sum((0, 2))])
    local = bst.clone()
    ecd.allreduce_stats(bst)
    torch.save({"y": y, "lo": lo, "hi": hi, "cnt": cnt, "mean": mean, "var": var, "dbeta_local": local[:Co],
                "dgamma_local": local[Co:], "dbeta": bst[:Co], "dgamma": bst[Co:]}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from dgcnn_pytorch_b200 import dist as ecd
    for n, w in ((32, 8), (5, 2), (3, 4), (1, 1), (0, 2)):
        spans = [ecd.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ecd.shard_range(4, 2, 2)
    with pytest.raises(TypeError):
        ecd.allreduce_stats(torch.zeros(3))


def test_syncbn_exchange_matches_single_process(tmp_path):
    import edgeconv_oracle as orc
    world, port, out = 2, _free_port(), str(tmp_path / "r")
    mp.start_processes(_worker, args=(world, port, out), nprocs=world, join=True, start_method="spawn")
    parts = [torch.load(f"{out}.{r}") for r in range(world)]
    assert [(p["lo"], p["hi"]) for p in parts] == [(0, 3), (3, 5)]
    # single-process reference: BatchNorm2d over ALL clouds' edges (what SyncBatchNorm means)
    B, C, N, k, Co = 5, 4, 48, 6, 8
    x = orc.synthetic_features(B, C, N, seed=3).double()
    wgt = torch.randn(Co, 2 * C, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    gamma = torch.linspace(-1.0, 1.5, Co, dtype=torch.float64).requires_grad_(True)
    beta = torch.linspace(0.3, -0.3, Co, dtype=torch.float64).requires_grad_(True)
    gout = torch.randn(B, Co, N, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    ref = orc.edgeconv_block_oracle(x, wgt, gamma, beta, None, None, k, training=True)
    (ref * gout).sum().backward()
    y = torch.cat([p["y"] for p in parts])
    assert torch.allclose(y, ref.detach(), rtol=1e-10, atol=1e-12)
    assert float(parts[0]["cnt"]) == B * N * k
    # global sums are identical on both ranks and equal the single-process gradients;
    # the per-rank (local) sums add up to them (DDP then averages parameter gradients)
    for key, grad in (("dbeta", beta.grad), ("dgamma", gamma.grad)):
        assert torch.allclose(parts[0][key], parts[1][key], rtol=0, atol=0)
        assert torch.allclose(parts[0][key], grad, rtol=1e-7, atol=1e-9)
        assert torch.allclose(parts[0][key + "_local"] + parts[1][key + "_local"], grad, rtol=1e-7, atol=1e-9)


def _grad_sync_worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dgcnn_pytorch_b200 import dist as ecd
    torch.set_num_threads(1)
    ecd.init_from_env("gloo")
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 2)) for _ in range(15)]     # 12 "late" + 3 "early"
    sync = ecd.FlatGradSync(params)
    assert not sync.overlap                                  # CPU: no side stream, one flat all-reduce
    assert sync.flat_late.numel() == 12 * 6 and sync.flat_early.numel() == 3 * 6
    # a step starts with every .grad None: autograd stores its gradients, nothing accumulates in place
    sync.zero()
    assert all(p.grad is None for p in params)
    loss = sum(((rank + 1.0) * (i + 1) * p).sum() for i, p in enumerate(params[:-1]))   # the last one gets no gradient
    loss.backward()                                          # d/dp = (rank+1)*(i+1)
    sync.average()
    # afterwards every .grad is its slice of the flat buffer, late bucket first
    assert params[0].grad.data_ptr() == sync.flat.data_ptr()
    assert params[3].grad.data_ptr() == sync.flat.data_ptr() + 3 * 6 * 4
    assert params[12].grad.data_ptr() == sync.flat_early.data_ptr()
    assert torch.equal(params[-1].grad, torch.zeros(3, 2))   # no gradient this step -> zeros, not stale values
    # a second step reuses the buffer
    sync.zero()
    loss = sum(((rank + 1.0) * (i + 1) * p).sum() for i, p in enumerate(params))
    loss.backward()
    sync.average()
    torch.save([p.grad.clone() for p in params], f"{out}.{rank}")
    dist.destroy_process_group()


def test_flat_grad_sync_buckets_and_average(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "g")
    mp.start_processes(_grad_sync_worker, args=(world, port, out), nprocs=world, join=True, start_method="spawn")
    a, b = (torch.load(f"{out}.{r}") for r in range(world))
    for i, (ga, gb) in enumerate(zip(a, b)):
        assert torch.equal(ga, gb)
        assert torch.allclose(ga, torch.full((3, 2), 1.5 * (i + 1)))      # mean of (1, 2) * (i+1)
