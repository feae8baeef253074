"""CUDA-graph execution of a training step.

At the classification shapes one step is ~150 kernels of 5-100 us each; launched
eagerly from Python the GPU idles between them.  ``GraphedTrainStep`` captures
zero_grad + forward + loss + backward + optimizer update ONCE into a CUDA graph
(every kernel of libedgeconv_b200 only enqueues work on the caller's stream and never
allocates or synchronises, so the whole EdgeConv path is capturable) and afterwards
replays it with one launch per step.  Inputs are copied into static device buffers
(from pinned host memory if given host tensors); the returned loss is a static tensor.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer,
                 loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                 example_x: torch.Tensor, example_y: torch.Tensor, warmup: int = 3,
                 grad_sync=None):
        """``grad_sync``: optional dist.FlatGradSync for data-parallel runs (its all-reduce is
        captured with the step)."""
        if not example_x.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA example tensors (no CPU fallback)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.grad_sync = grad_sync
        self.x = example_x.clone()
        self.y = example_y.clone()
        # warm-up on a side stream: sizes allocator pools, opts kernels into large shared memory
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if grad_sync is None:
            self.optimizer.zero_grad(set_to_none=True)
        # thread_local: NCCL's watchdog thread may query events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = self._eager()
        torch.cuda.synchronize()

    def _eager(self) -> torch.Tensor:
        if self.grad_sync is not None:
            self.grad_sync.zero()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.x), self.y)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.average()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """x, y: device tensors or pinned host tensors of the captured shapes."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
