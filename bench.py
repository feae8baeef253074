#!/usr/bin/env python
"""Headline benchmark: DGCNN-cls forward+backward clouds/sec (N=1024, k=20) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the training hot path over one batch of synthetic clouds:
zero_grad, DGCNN-cls forward (4 fused EdgeConv layers + conv5 + cls head),
label-smoothed CE, backward, SGD update.  Prints ONE JSON line (rank 0).

  value     whole-job clouds/s with inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e       the same through the public API from pinned HOST buffers (H2D of the batch and
            D2H of the loss inside the timed region)
  roofline  the dominant kernel of the EdgeConv path against the measured HBM peak
  cpu_baseline  the oracle (CPU restatement of the reference) on the host cores, bounded sample

--impl reference times the reference's CPU implementation of the same step (the oracle port:
the reference itself is Python under /root/reference and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "DGCNN-cls fwd+bwd clouds/sec (N=1024,k=20)"
UNIT = "clouds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU per step")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--emb", type=int, default=1024)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only")
    ap.add_argument("--cpu-clouds", type=int, default=8, help="clouds per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


def workload_name(a):
    return (f"DGCNN-cls train step (fwd+bwd+SGD), synthetic ModelNet40-shape xyz clouds, "
            f"B={a.batch}/GPU N={a.points} k={a.k} emb={a.emb}")


# ------------------------------------------------------------------- CPU reference arm
def cpu_step_fn(a, clouds):
    import edgeconv_oracle as orc
    torch.manual_seed(1)
    args = SimpleNamespace(emb_dim=a.emb, k=a.k, dropout=0.5)
    model = orc.DGCNNClsOracle(args).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = orc.synthetic_xyz(clouds, a.points, seed=1)
    y = torch.randint(0, 40, (clouds,), generator=torch.Generator().manual_seed(1))

    def step():
        opt.zero_grad(set_to_none=True)
        loss = orc.smoothed_ce_oracle(model(x), y)
        loss.backward()
        opt.step()
        return loss.item()
    return step


def time_cpu(a, clouds, steps, warmup):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_step_fn(a, clouds)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return clouds / statistics.median(ts), statistics.median(ts), cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clouds = a.cpu_clouds
    cps, sec, cores = time_cpu(a, clouds, max(1, a.steps), max(1, a.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3 * a.batch / clouds,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "device": "host CPU"},
        "cpu_baseline": {"value": cps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{clouds} clouds per step (same N, k, emb, same step), median of "
                                   f"{max(1, a.steps)} steps, torch CPU threads = {cores}"},
        "e2e": {"value": cps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for t, line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                mhz, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            smax.append(mx)
            if t0 <= t <= t1 + 0.2:
                sm.append(mhz)
                with contextlib.suppress(ValueError):
                    power.append(float(f[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------- per-entry event timing
class EntryTimer:
    """Brackets every C-ABI call with CUDA events on the launching stream."""

    def __init__(self):
        self.events = {}
        self.enabled = False

    @contextlib.contextmanager
    def __call__(self, name):
        if not self.enabled:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.events.setdefault(name, []).append((s, e))

    def summary(self):
        out = {}
        for name, evs in self.events.items():
            ms = [s.elapsed_time(e) for s, e in evs]
            out[name] = {"calls": len(ms), "total_ms": sum(ms), "avg_us": 1e3 * sum(ms) / len(ms)}
        return out


def note(msg):
    """progress notes on stderr (BENCH_DEBUG=1)"""
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def edge_layer_shapes(a):
    """(C, Co) of the four EdgeConv layers (models/dgcnn.py:54-73)."""
    return [(3, 64), (64, 64), (64, 128), (128, 256)]


def gather_bytes(M, k, Co, training=True):
    """Algorithmic bytes of one ecb200_edge_gather launch: idx + the k gathered
    U rows + the V row in; sel, arg (+ esum in training) out.  SpMM convention: every gathered
    row counts once wherever it is served from."""
    b = 4 * M * k + 4 * M * k * Co + 4 * M * Co + 4 * M * Co + M * Co
    if training:
        b += 4 * M * Co
    return b


# ------------------------------------------------------------------------ B200 arm
def run_b200(a):
    import torch.distributed as dist

    import dgcnn_pytorch_b200 as ec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. the NCCL
    # version banner) is routed to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); "
                           "use --impl reference for the CPU reference arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # collectives get graph-captured
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    B = a.batch if a.scaling == "weak" else max(1, a.batch // world)
    N, k = a.points, a.k

    torch.manual_seed(1)
    args = SimpleNamespace(emb_dim=a.emb, k=k, dropout=0.5)
    model = ec.DGCNN_cls(args).to(dev).train()
    sync = None
    stats_exchange = None
    if world > 1:
        # SyncBatchNorm semantics for every BN (the EdgeConv ones exchange their statistics
        # inside the fused op); gradients averaged by one flat all-reduce per step
        from dgcnn_pytorch_b200.dist import FlatGradSync
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        for p in model.parameters():          # identical replicas
            dist.broadcast(p.data, 0)
        sync = FlatGradSync(model.parameters())
        if os.environ.get("ECB200_STATS_EXCHANGE", "peer") == "peer":
            try:   # BatchNorm statistics over NVLink peer memory, one kernel per exchange
                from dgcnn_pytorch_b200.dist import PeerStatsExchange
                peer_exchange = PeerStatsExchange.enable()
                stats_exchange = "one-kernel push exchange over NVLink peer memory (symmetric memory)"
            except Exception as exc:  # noqa: BLE001 - transport fallback, reported in the JSON line
                stats_exchange = f"NCCL all-reduce (peer memory unavailable: {type(exc).__name__}: {exc})"[:160]
        else:
            stats_exchange = "NCCL all-reduce"
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)

    import edgeconv_oracle as orc   # only for the synthetic-input generator and the cpu_baseline leg
    npool = 4
    host_x = [orc.synthetic_xyz(B, N, seed=100 * rank + i).pin_memory() for i in range(npool)]
    host_y = [torch.randint(0, 40, (B,), generator=torch.Generator().manual_seed(7 * rank + i)).pin_memory()
              for i in range(npool)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(x, y):
        if sync is not None:
            sync.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss = ec.cal_loss(model(x), y)
        loss.backward()
        if sync is not None:
            sync.average()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    timer = EntryTimer()
    ec._lib.set_event_hook(timer)

    # ---- warm-up (also sizes the allocator pools and opts kernels into big smem)
    note("warm-up")
    for i in range(max(3, a.warmup)):
        step(dev_x[i % npool], dev_y[i % npool])
    barrier()
    note("warm-up done")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_clock0 = time.perf_counter()

    # ---- timed region A: eager launches, per-step CUDA events, L2 flushed between steps,
    #      every C-ABI entry bracketed by events (per-kernel durations for the roofline)
    def timed(fn, count):
        evs = []
        barrier()
        for i in range(count):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(i)
            e.record()
            evs.append((s, e))
        barrier()
        return sum(s.elapsed_time(e) for s, e in evs) / count

    timer.enabled = True
    launches0 = ec._lib.launch_count
    eager_ms = timed(lambda i: step(dev_x[i % npool], dev_y[i % npool]), a.steps)
    launches_per_step = (ec._lib.launch_count - launches0) / a.steps
    timer.enabled = False
    eager_ms = reduce_max(eager_ms)
    entries = timer.summary()
    note(f"eager timed: {eager_ms:.3f} ms/step")

    # ---- timed region B: the same step captured once into a CUDA graph and replayed
    #      (dgcnn_pytorch_b200.GraphedTrainStep, the package's public way to run a step)
    graph_ms, graph_err, gstep = None, None, None
    if not a.no_graph:
        try:
            gstep = ec.GraphedTrainStep(model, opt, ec.cal_loss, dev_x[0], dev_y[0], grad_sync=sync)
            note("graph captured")
            for i in range(3):
                gstep(dev_x[i % npool], dev_y[i % npool])
            graph_ms = reduce_max(timed(lambda i: gstep(dev_x[i % npool], dev_y[i % npool]), a.steps))
        except Exception as exc:  # noqa: BLE001 - reported, the eager number stands
            graph_err = f"{type(exc).__name__}: {exc}"[:200]
            gstep = None
            torch.cuda.synchronize()
    note(f"graph timed: {graph_ms} ms/step, error {graph_err}")
    use_graph = graph_ms is not None and graph_ms < eager_ms
    best_ms = graph_ms if use_graph else eager_ms

    # ---- end to end through the public API from pinned HOST buffers: per step the batch is
    #      copied host -> device and the loss is read back device -> host
    def e2e_step(i):
        hx, hy = host_x[i % npool], host_y[i % npool]
        if use_graph:
            return gstep(hx, hy).item()
        return step(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)).item()
    e2e_ms, e2e_err = float("nan"), None
    try:
        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            e2e_step(i)
        barrier()
        e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3 / a.steps)
    except Exception as exc:  # noqa: BLE001 - reported in the JSON line
        e2e_err = f"{type(exc).__name__}: {exc}"[:200]
    t_clock1 = time.perf_counter()
    clocks = sampler.stop(t_clock0, t_clock1) if rank == 0 else None

    if rank != 0:
        finish(world)
        return

    pk = peaks()
    clouds = B * world
    M = B * N
    # roofline of the dominant EdgeConv kernel: the neighbour gather (HBM-bound by the SpMM
    # convention).  Its launches differ per layer, so achieved = sum(bytes) / sum(time).
    gather = entries.get("ecb200_edge_gather")
    roof = None
    if gather:
        per_step_calls = gather["calls"] // a.steps
        by = sum(gather_bytes(M, k, co) for _, co in edge_layer_shapes(a))
        t_ms = gather["total_ms"] / a.steps
        ach = by / (t_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_gather_traffic.json")
        if (a.batch, a.points, a.k) == (32, 1024, 20) and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)["dram_bytes_per_step"]     # ncu --set full capture, per step
        roof = {"bound": "hbm", "kernel": "edge_gather_kernel", "achieved": ach, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic,
                "peak_source": pk["source"], "launches_per_step": per_step_calls,
                "algorithmic_bytes_per_step": by, "ms_per_step": t_ms}
    # second roofline: the tensor-core kNN of the feature-space layers against the 3xTF32 tensor
    # roofline (bf16 peak / 2 for TF32 / 3 MMAs per product); FLOPs = 2*M*N*C per layer
    roof_knn = None
    try:
        ktc = entries.get("ecb200_knn_tc")
        if ktc:
            flops = sum(2.0 * M * N * c for c, _ in edge_layer_shapes(a) if c % 32 == 0 and 32 <= c <= 128)
            t_ms = ktc["total_ms"] / a.steps
            peak = pk["bf16_tflops"] / 2.0 / 3.0
            ach = flops / (t_ms * 1e-3) / 1e12
            roof_knn = {"bound": "tensor", "kernel": "knn_tc_kernel", "achieved": ach, "peak": peak,
                        "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": pk["source"] + " bf16 / 2 (tf32) / 3 (3xTF32)",
                        "launches_per_step": ktc["calls"] // a.steps, "flops_per_step": flops,
                        "ms_per_step": t_ms}
    except Exception:  # noqa: BLE001 - an extra, never at the expense of the line
        roof_knn = None
    breakdown = {n: round(v["total_ms"] / a.steps, 4) for n, v in
                 sorted(entries.items(), key=lambda kv: -kv[1]["total_ms"])}

    cpu = None
    if not a.no_cpu_baseline and world >= 1:
        cps, sec, cores = time_cpu(a, a.cpu_clouds, 3, 1)
        cpu = {"value": cps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{a.cpu_clouds} clouds per step of the same workload, 1 warm-up + median of 3 "
                         f"steps ({sec:.2f} s/step)"}

    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    line = {
        "metric": METRIC, "value": clouds / (best_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": best_ms,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": clouds, "parallelism": f"dp{world}",
                   "l2": "256 MiB buffer rewritten between timed steps (L2 flush)",
                   "cuda_graph": use_graph,
                   "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
                   "graph_error": graph_err, "conv5_and_head": "conv5 GEMM in cuDNN (library default TF32) on the channels-last concat; its BatchNorm + LeakyReLU + max|avg pooling in own kernels (embed_pool); 3 head linears in torch",
                   "grad_sync": "one flat NCCL all-reduce per step" if world > 1 else None,
                   "bn_stats_exchange": stats_exchange},
        "clocks": clocks,
        "e2e": {"value": clouds / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "error": e2e_err},
        "gpu_launches": int(round(launches_per_step * a.steps)),
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "roofline_knn": roof_knn,
        "kernel_ms_per_step": breakdown,
        "cpu_baseline": cpu,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    finish(world)


def finish(world):
    """Leave without tearing NCCL down: communicators referenced by captured CUDA graphs can
    make destroy_process_group() wait forever, so multi-rank runs exit hard once every rank
    has finished (the JSON line is already flushed)."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
