"""Multi-rank parity of the CUDA path (-m gpu): a sharded training step (SyncBatchNorm semantics,
statistics exchange between the gather and finalize kernels, flat gradient all-reduce) must equal
the single-process step on the whole batch (main_partseg_dist.py:189-196 semantics).

Each case launches tools/check_ddp_equivalence.py under torchrun:
  * always: world sizes 2 and 8 over the gloo backend with the ranks SHARING cuda:0 -- the
    world-size-N protocol (what is summed, who divides by what, gradient averaging) is exercised
    on the one-GPU box the driver runs the suite on, and the check also reports both sides against
    the fp64 CPU oracle;
  * with >= 2 GPUs: NCCL, one GPU per rank, BatchNorm statistics over NVLink peer memory
    (csrc/peer_exchange.cu) -- the transport the benchmark uses.
Tolerance: worst relative deviation (max|a-b| / max|b| per tensor) of loss, every averaged
gradient and every BatchNorm buffer <= 2e-4 on at least one of up to three batches, and no batch worse
than one LeakyReLU-kink / tied-max flip explains (see the verdict comment in the tool).
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "tools", "check_ddp_equivalence.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, backend, tmp_path, extra_env=None, timeout=600):
    out = tmp_path / f"ddp_{backend}_{world}.jsonl"
    env = dict(os.environ, DDP_CHECK_BACKEND=backend, DDP_CHECK_ARBITER="1", DDP_CHECK_OUT=str(out),
               OMP_NUM_THREADS="2", **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), TOOL]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, f"{backend} x{world} failed:\n{tail}"
    res = json.loads(out.read_text().strip().splitlines()[-1])
    assert res["ok"] and res["world"] == world, res
    # gross-error check of both sides against the fp64 oracle run on the same graphs (loose: against
    # fp64 an activation within fp32 rounding of a LeakyReLU kink, or a tied max, may resolve the other
    # way and move a few gradient tensors by up to a per cent)
    assert res["vs_fp64_oracle"]["sharded"] < 5e-2 and res["vs_fp64_oracle"]["full_batch"] < 5e-2, res
    return res


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_step_equals_full_batch_shared_gpu(world, tmp_path):
    _run(world, "gloo", tmp_path)


def test_sharded_step_equals_full_batch_b32_n1024(tmp_path):
    # the benchmark's per-cloud shape (N = 1024, k = 20): 4 ranks x 2 clouds on one GPU
    _run(4, "gloo", tmp_path, {"DDP_CHECK_B": "8", "DDP_CHECK_N": "1024", "DDP_CHECK_K": "20"})


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_equals_full_batch_nccl_peer(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, "nccl", tmp_path)
