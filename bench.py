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

  roofline  the dominant kernel of the step (tensor-core kNN) against the 3xTF32 tensor roofline;
            roofline_gather / roofline_edgeconv: the gather kernel against HBM and L2, and the whole
            EdgeConv forward against its combined bound
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, a verbatim copy made by
            baseline/install_ref.py) on the host cores, bounded sample
  gpu_eager_reference  the same unmodified reference run eagerly on this B200, TF32 off

--impl reference times the reference's own CPU implementation of the same step at the same batch
(baseline/_ref; falls back to the oracle port only if that copy is missing).
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
for p in (ROOT, os.path.join(ROOT, "baseline")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "DGCNN-cls fwd+bwd clouds/sec (N=1024,k=20)"
UNIT = "clouds/s"
CONV5_NOTE = ("conv5 GEMM: library convolution (cuDNN, TF32 -- torch.backends.cudnn.allow_tf32 is on, PyTorch's default) on the "
              "channels-last concat; an own tcgen05 GEMM with the BatchNorm statistics in its epilogue exists "
              "(ecb200_embed_gemm: 3xTF32 whenever allow_tf32 is off, TF32 by ECB200_CONV5=tf32) but its 128x128 tiles "
              "are 2.9x slower than the library's 2-SM 256x256 kernel; BatchNorm + LeakyReLU + max|avg pooling in own "
              "kernels (embed_pool); 3 head linears in torch")


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
    ap.add_argument("--no-gpu-eager-reference", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--profile-ranks", action="store_true",
                    help="multi-rank runs: also take the per-kernel profile of the replayed graph (rank 0 reports)")
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
def reference_model(a):
    """(model, loss_fn, kind): the unmodified reference (baseline/_ref) + upstream cls head, or --
    only when that copy is missing -- the oracle port of the same network."""
    import ref_cls
    args = SimpleNamespace(emb_dim=a.emb, k=a.k, dropout=0.5)
    torch.manual_seed(1)
    if ref_cls.available():
        return ref_cls.RefDGCNNCls(args), ref_cls.reference_loss, "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import edgeconv_oracle as orc            # the one other place bench.py may execute oracle/
    return orc.DGCNNClsOracle(args), orc.smoothed_ce_oracle, "port"


def time_cpu(a, clouds, steps, warmup, budget_s=None):
    """Reference train step on the host cores: (clouds/s, s/step, cores, kind, steps timed)."""
    from dgcnn_pytorch_b200.synthetic import synthetic_xyz
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    model, loss_fn, kind = reference_model(a)
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = synthetic_xyz(clouds, a.points, seed=1)
    y = torch.randint(0, 40, (clouds,), generator=torch.Generator().manual_seed(1))

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(x), y)
        loss.backward()
        opt.step()
        return loss.item()
    for _ in range(warmup):
        step()
    ts = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    sec = sum(ts) / len(ts)
    return clouds / sec, sec, cores, kind, len(ts)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same configuration as the b200 arm: B clouds per step, same N, k, emb, same train step.
    # One step is seconds of CPU work, so warm-up is capped at 1 and the timed loop stops early
    # if it would exceed ~4 minutes (the number of steps actually timed is reported).
    clouds = a.batch
    cps, sec, cores, kind, done = time_cpu(a, clouds, max(1, a.steps), min(max(1, a.warmup), 1), budget_s=240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": UNIT, "n_gpus": a.gpus,
        "steps": done, "warmup": min(max(1, a.warmup), 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "device": "host CPU", "same_config": True,
                   "global_batch": clouds,
                   "code": "unmodified reference models/dgcnn.py (baseline/_ref) + upstream cls head"
                           if kind == "reference" else "oracle port (baseline/_ref missing)"},
        "cpu_baseline": {"value": cps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{clouds} clouds per step (the full batch, same N, k, emb, same train step), "
                                   f"mean of {done} steps, torch CPU threads = {cores}"},
        "e2e": {"value": cps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_gpu_eager_reference(a, dev, B):
    """The unmodified reference (same network, same train step) run eagerly on this GPU with TF32
    off: the strongest existing implementation of the path."""
    import ref_cls
    from dgcnn_pytorch_b200.synthetic import synthetic_xyz
    if not ref_cls.available():
        return {"unavailable": "baseline/_ref missing"}
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(1)
        model = ref_cls.RefDGCNNCls(SimpleNamespace(emb_dim=a.emb, k=a.k, dropout=0.5)).to(dev).train()
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
        x = synthetic_xyz(B, a.points, seed=1).to(dev)
        y = torch.randint(0, 40, (B,), generator=torch.Generator().manual_seed(1)).to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = ref_cls.reference_loss(model(x), y)
            loss.backward()
            opt.step()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        n = 20
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        peak = torch.cuda.max_memory_allocated(dev)
        del model, opt, x, y
        torch.cuda.empty_cache()
        return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": n, "warmup": 5,
                "tf32": False, "peak_mem_bytes": int(peak),
                "what": "unmodified reference DGCNN (baseline/_ref) + cls head, eager torch on this GPU, same step"}
    except Exception as exc:  # noqa: BLE001 - a comparator leg never takes the line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


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


def q_edge_bytes(M, C, k, Co):
    """algorithmic bytes of the edge-MLP + max of one layer, SpMM convention (every
    gathered row counts once, wherever it is served from): x in + idx in + k gathered U rows + out."""
    return 4 * M * (C + k + k * Co + Co)


def gather_kernel_bytes(M, k, Co, training=True):
    """What ecb200_edge_gather itself moves per launch: idx + the k gathered U rows + the V row in;
    sel, arg (+ esum in training) out.  Reported beside the SURVEY figure."""
    b = 4 * M * k + 4 * M * k * Co + 4 * M * Co + 4 * M * Co + M * Co
    if training:
        b += 4 * M * Co
    return b


# forward kernels of the EdgeConv path (kNN + edge MLP + max), by name fragments of the kernels
EDGE_FWD_KERNELS = ("knn_tc_kernel<32, false", "knn_tc_kernel<32, 0", "knn_tc2_kernel", "knn_xyz_kernel", "knn_fma_kernel", "sqnorms_kernel",
                    "split_tf32_kernel", "split_f16_kernel", "pack_xyz_f16_kernel", "absmax_kernel",
                    "prepare_weights_kernel", "pack_weight_kernel", "gemm_tile_kernel", "knn_tc_kernel<32, true", "edge_gather_kernel",
                    "bn_finalize_kernel", "edge_apply_kernel", "bn_update_running_kernel")


def short_kernel_name(name):
    """'void <unnamed>::knn_tc_kernel<32, 0>(const float*, ...)' -> 'knn_tc_kernel<32, 0>'"""
    n = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if n.startswith("void "):
        n = n[5:]
    depth, out = 0, []
    for ch in n:                      # cut the argument list (first '(' at template depth 0)
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out)[:96]


def profile_graph_kernels(replay, nsteps):
    """Per-kernel device time INSIDE the replayed CUDA graph (CUPTI activity records through
    torch.profiler): {kernel: {"calls_per_step", "us_per_step"}}.  Not the timed region -- the
    headline comes from CUDA events without any profiler -- but the same graph, same inputs."""
    from torch.profiler import ProfilerActivity, profile
    replay(0)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(nsteps):
            replay(i)
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if str(getattr(ev, "device_type", "")).endswith("CUDA") and ev.name and not ev.name.startswith("Memcpy") \
                and not ev.name.startswith("Memset"):
            dur = getattr(ev, "device_time", None)
            if dur is None:
                dur = getattr(ev, "cuda_time", 0.0)
            if "FillFunctor<unsigned char>" in ev.name:
                continue          # the L2 flush between the replays, not part of the step
            d = agg.setdefault(short_kernel_name(ev.name), [0, 0.0])
            d[0] += 1
            d[1] += float(dur)
    return {n: {"calls_per_step": c / nsteps, "us_per_step": us / nsteps} for n, (c, us) in agg.items()}


# ------------------------------------------------------------------------ B200 arm
def run_b200(a):
    import torch.distributed as dist

    import dgcnn_pytorch_b200 as ec
    from dgcnn_pytorch_b200.synthetic import synthetic_xyz

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
        # inside the fused op, the head's through the same exchange); gradients averaged by one
        # flat all-reduce per step
        from dgcnn_pytorch_b200.dist import FlatGradSync, PeerStatsExchange
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        for p in model.parameters():          # identical replicas
            dist.broadcast(p.data, 0)
        sync = FlatGradSync(model.parameters())
        if os.environ.get("ECB200_STATS_EXCHANGE", "peer") == "peer":
            # BatchNorm statistics over NVLink peer memory, one kernel per exchange; every rank
            # must agree on the transport, or the pushing ranks would wait for words that never come
            stats_exchange = PeerStatsExchange.enable_collectively()
        else:
            stats_exchange = "NCCL all-reduce"
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)

    npool = 4
    host_x = [synthetic_xyz(B, N, seed=100 * rank + i).pin_memory() for i in range(npool)]
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
    #      every C-ABI entry bracketed by events (fallback per-kernel durations)
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

    # ---- per-kernel time inside the replayed graph (after the timed regions; single-rank runs)
    kprof, kprof_err = None, None
    # (multi-rank: every rank replays -- the graph holds the exchanges -- and rank 0 reports; the statistics
    # exchange kernels then show their wait for the slowest rank, i.e. where the scaling loss sits)
    if use_graph and not a.no_kernel_profile and (world == 1 or a.profile_ranks):
        try:
            def replay(i):
                flush.zero_()
                gstep(dev_x[i % npool], dev_y[i % npool])
            kprof = profile_graph_kernels(replay, min(a.steps, 10))
        except Exception as exc:  # noqa: BLE001
            kprof_err = f"{type(exc).__name__}: {exc}"[:200]

    if rank != 0:
        finish(world)
        return

    pk = peaks()
    clouds = B * world
    M = B * N
    tf32x3_peak = pk["bf16_tflops"] / 2.0 / 3.0
    layers = edge_layer_shapes(a)

    def kernel_us(fragment):
        """us per step of the kernels whose name contains `fragment`: in-graph profile if there is
        one, else the eager event brackets of the matching entry point."""
        if kprof:
            return sum(v["us_per_step"] for n, v in kprof.items() if fragment in n)
        return None

    def entry_ms(name):
        e = entries.get(name)
        return e["total_ms"] / a.steps if e else None

    # roofline 1 (dominant kernel of the step): tensor-core kNN of the feature-space layers against
    # the 3xTF32 tensor roofline (bf16 peak / 2 for TF32 / 3 MMAs per product); FLOPs = 2*M*N*C per
    # layer (norm terms and the error-compensation MMAs do not count)
    roof = None
    knn_flops = sum(2.0 * M * N * c for c, _ in layers if c % 32 == 0 and 32 <= c <= 128)
    f16x3_peak = pk["bf16_tflops"] / 3.0

    def knn_tc_us(xyz):
        """kNN tensor-core kernels of the graph profile: template tail '<.., F16, TERMS>'; TERMS == 1 is
        the xyz layer (one MMA per tile), TERMS == 3 the feature-space layers."""
        if not kprof:
            return None, None
        t, f16 = 0.0, False
        for n, v in kprof.items():
            if n.startswith("knn_tc2_kernel<"):          # 256-row variant: packed fp16 only, <TERMS>
                is_xyz, is_f16 = n.rstrip(">").split("<")[-1].strip() == "1", True
            elif n.startswith("knn_tc_kernel<32, false") or n.startswith("knn_tc_kernel<32, 0"):
                targs = [t.strip() for t in n[n.index("<") + 1:n.rindex(">")].split(",")]   # NBINS, DEBUG, CL, S, F16, TERMS, FOLD
                is_xyz = len(targs) >= 6 and targs[5] == "1"
                is_f16 = len(targs) >= 5 and targs[4] in ("true", "1")
            else:
                continue
            if is_xyz != xyz:
                continue
            t += v["us_per_step"]
            f16 = f16 or is_f16
        return (t or None), f16
    t_us, knn_f16 = knn_tc_us(False)
    src = "in-graph (CUPTI)"
    if not t_us and (entry_ms("ecb200_knn_tc_f16") or entry_ms("ecb200_knn_tc")):
        knn_f16 = bool(entry_ms("ecb200_knn_tc_f16"))
        t_us, src = (entry_ms("ecb200_knn_tc_f16") or entry_ms("ecb200_knn_tc")) * 1e3, "eager CUDA-event brackets"
    if t_us:
        ach = knn_flops / (t_us * 1e-6) / 1e12
        # `peak` / `frac` follow SURVEY 8d's definition of this kernel's roofline (fp32-equivalent products on the
        # tensor pipe = TF32 peak / 3), the yardstick of round 1; the kernel now reaches the same accuracy with
        # kind::f16 MMAs, whose pipe peak is twice that: `frac_of_f16x3_pipe` is the fraction of the pipe in use
        knn_traffic = None
        kpath = os.path.join(ROOT, "profiles", "r2c_knn_traffic.json")
        if (a.batch, a.points, a.k) == (32, 1024, 20) and os.path.exists(kpath):
            with open(kpath) as f:
                knn_traffic = json.load(f).get("dram_bytes_per_step")     # ncu --set full, DRAM read + write per step
        roof = {"bound": "tensor", "kernel": "knn_tc_kernel", "achieved": ach, "peak": tf32x3_peak,
                "unit": "TFLOP/s", "frac": ach / tf32x3_peak, "traffic": knn_traffic,
                "operands": ("packed fp16 hi/lo halves (kind::f16, 3 MMAs per product: same 11-bit significands and "
                             "error bound as 3xTF32 at twice the pipe rate)") if knn_f16 else "tf32 hi/lo halves (3xTF32)",
                "peak_source": pk["source"] + " bf16 burst / 2 (tf32) / 3 (3xTF32), SURVEY 8d",
                "frac_of_3xtf32_roofline": ach / tf32x3_peak,
                "frac_of_f16x3_pipe": (ach / f16x3_peak) if knn_f16 else None,
                "f16x3_pipe_peak": f16x3_peak if knn_f16 else None,
                "bound_note": ("measured: tensor pipe ~20 % active, issue slots 45 %; the kernel is bound by the issue "
                               "rate of its selection epilogue (profiles/r2c_knn_f16_full.txt, r2c_ubench_pipes.txt)"),
                "launches_per_step": 3, "flops_per_step": knn_flops, "us_per_step": t_us, "timing": src}
    # roofline 2: the neighbour gather against HBM (SURVEY's Q_edge and the kernel's own byte count)
    roof_gather = None
    t_us = kernel_us("edge_gather_kernel")
    src = "in-graph (CUPTI)"
    if not t_us and entry_ms("ecb200_edge_gather"):
        t_us, src = entry_ms("ecb200_edge_gather") * 1e3, "eager CUDA-event brackets"
    if t_us:
        q_survey = sum(q_edge_bytes(M, c, k, co) for c, co in layers)
        q_kernel = sum(gather_kernel_bytes(M, k, co) for _, co in layers)
        ach = q_survey / (t_us * 1e-6) / 1e9
        traffic = lts = None
        tpath = os.path.join(ROOT, "profiles", "r2_gather_traffic.json")
        if (a.batch, a.points, a.k) == (32, 1024, 20) and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, lts = tj.get("dram_bytes_per_step"), tj.get("lts_bytes_per_step")
        roof_gather = {"bound": "hbm", "kernel": "edge_gather_kernel", "achieved": ach, "peak": pk["hbm_gbs"],
                       "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic,
                       "lts_bytes_per_step": lts,
                       "note": "rows are served from L2 (DRAM traffic << algorithmic bytes): the binding resource is "
                               "L2 bandwidth, the HBM fraction only says the SpMM-convention bytes move faster than HBM could",
                       "algorithmic_bytes_per_step": q_survey, "kernel_bytes_per_step": q_kernel,
                       "achieved_kernel_bytes_gbs": q_kernel / (t_us * 1e-6) / 1e9,
                       "l2_gbs": (lts / (t_us * 1e-6) / 1e9) if lts else None,
                       "us_per_step": t_us, "timing": src}
    # roofline 3: xyz kNN against the FP32 FMA issue peak (3 FMAs per pair, one sweep is algorithmic)
    roof_xyz = None
    xyz_kernel = "knn_xyz_kernel"
    t_us = kernel_us("knn_xyz_kernel")
    if not t_us and knn_tc_us(True)[0]:
        t_us, xyz_kernel = knn_tc_us(True)[0], "tensor-core kNN, TERMS = 1 (one kind::f16 MMA per tile and sweep; knn_tc2_kernel<1> or knn_tc_kernel<.., true, 1>)"
    if not t_us and entry_ms("ecb200_knn_tc_xyz"):
        t_us, xyz_kernel = entry_ms("ecb200_knn_tc_xyz") * 1e3, "ecb200_knn_tc_xyz (eager brackets)"
    if not t_us and entry_ms("ecb200_knn"):
        t_us = entry_ms("ecb200_knn") * 1e3
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    fma_peak = 148 * 128 * sm_mhz * 1e6          # FMA/s
    if t_us:
        pairs = float(M) * N
        roof_xyz = {"bound": "fp32-fma issue", "kernel": xyz_kernel, "pairs_per_s": pairs / (t_us * 1e-6),
                    "fma_per_pair_algorithmic": 3, "frac_of_fma_peak": 3 * pairs / (t_us * 1e-6) / fma_peak,
                    "fma_peak_per_s": fma_peak, "us_per_step": t_us}
    # roofline 4: the whole fused EdgeConv forward (kNN + edge MLP + max, 4 layers) against its
    # combined bound: Q_edge / HBM + F_knn / (TF32/3) + xyz pairs * 3 FMA / FMA peak
    roof_edge = None
    if kprof:
        t_fwd = sum(v["us_per_step"] for n, v in kprof.items()
                    if any(f in n for f in EDGE_FWD_KERNELS) and "EpiD" not in n)
        q_survey = sum(q_edge_bytes(M, c, k, co) for c, co in layers)
        bound_us = (q_survey / (pk["hbm_gbs"] * 1e9) + knn_flops / (tf32x3_peak * 1e12)
                    + 3.0 * M * N / fma_peak) * 1e6
        roof_edge = {"bound_us": bound_us, "measured_us": t_fwd, "frac": bound_us / t_fwd if t_fwd else None,
                     "kernels": [f for f in EDGE_FWD_KERNELS],
                     "what": "sum of the forward EdgeConv kernels inside the replayed graph (4 layers) vs "
                             "SURVEY 8d: Q_edge/HBM + F_knn/(TF32/3) + xyz FMA bound"}
    if kprof:
        breakdown = {n: round(v["us_per_step"] / 1e3, 4) for n, v in
                     sorted(kprof.items(), key=lambda kv: -kv[1]["us_per_step"])[:40]}
        breakdown_src = "CUPTI kernel records of the replayed CUDA graph (separate pass after the timed region)"
        breakdown_sum = sum(v["us_per_step"] for v in kprof.values()) / 1e3
    else:
        breakdown = {n: round(v["total_ms"] / a.steps, 4) for n, v in
                     sorted(entries.items(), key=lambda kv: -kv[1]["total_ms"])}
        breakdown_src = "CUDA-event brackets around the C-ABI calls of the eager pass" + (f" ({kprof_err})" if kprof_err else "")
        breakdown_sum = sum(breakdown.values())

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        try:
            cps, sec, cores, kind, done = time_cpu(a, a.cpu_clouds, 3, 1, budget_s=60.0)
            cpu = {"value": cps, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{a.cpu_clouds} clouds per step of the same workload (the --impl reference arm runs "
                             f"the full batch), 1 warm-up + mean of {done} steps ({sec:.2f} s/step)"}
        except Exception as exc:  # noqa: BLE001
            cpu = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    eager_ref = None
    if world == 1 and not a.no_gpu_eager_reference:
        del flush
        torch.cuda.empty_cache()
        eager_ref = time_gpu_eager_reference(a, dev, B)

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
                   "graph_error": graph_err, "conv5_and_head": CONV5_NOTE,
                   "grad_sync": "one flat NCCL all-reduce per step" if world > 1 else None,
                   "bn_stats_exchange": stats_exchange},
        "clocks": clocks,
        "e2e": {"value": clouds / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "error": e2e_err},
        "gpu_launches": int(round(launches_per_step * a.steps)),
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "roofline_gather": roof_gather,
        "roofline_knn_xyz": roof_xyz,
        "roofline_edgeconv": roof_edge,
        "kernel_ms_per_step": breakdown,
        "kernel_ms_per_step_source": breakdown_src,
        "kernel_ms_per_step_sum": breakdown_sum,
        "cpu_baseline": cpu,
        "gpu_eager_reference": eager_ref,
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
