"""CUDA-graph step runners for the Chamfer hot path.

At the headline shape (B=32, N=M=2048) one ChamferLoss forward+backward is ~50 us of GPU work in two kernel
launches, less than the Python/ctypes cost of enqueuing it.  These helpers capture S steps -- exactly the calls a
user makes, `loss = ChamferLoss()(pred, target); loss.backward()` -- into one CUDA graph so the launch-bound
loop is replayed by the driver instead of re-issued from Python (train_rl_gan_net.py:220-249 is such a loop).

  ChamferStepGraph      S steps over S device-resident batches.
  HostChamferStepGraph  S steps over S pinned host batches: every step's H2D copy of its inputs and D2H copy of
                        its loss are inside the graph, double-buffered on a copy stream so transfers overlap
                        the previous step's kernels.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .chamfer import ChamferLoss, _use_tensor


def launches_per_step(n: int, m: int) -> int:
    """Kernels of this library in one ChamferLoss forward + backward: the fused tensor forward is ONE launch, the FP32
    path two (pair sweep + refinement); the backward is one (the forward zero-fills the gradient buffers)."""
    return 2 if _use_tensor(None, n, m) else 3


class ChamferStepGraph:
    """Capture `for pred, target in batches: loss = crit(pred, target); loss.backward()` in one CUDA graph.
    After replay(): self.losses[k] (0-dim tensors) and self.grads[k] (d loss / d pred) hold step k's results."""

    def __init__(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], bidirectional: bool = True,
                 grad_target: bool = False, after_step: Optional[Callable[[torch.Tensor], None]] = None):
        """after_step(loss): optional hook captured after every step ON A SIDE STREAM (forked after the step's kernels,
        joined at the end of the graph) -- e.g. the data-parallel all-reduce of the logged loss scalar, which then
        overlaps the next step's kernels instead of serialising with them."""
        assert len(batches) > 0 and batches[0][0].is_cuda
        self.crit = ChamferLoss(bidirectional)
        self.batches = [(a.detach().requires_grad_(True), b.detach().requires_grad_(grad_target)) for a, b in batches]
        self.device = self.batches[0][0].device
        self.losses: List[torch.Tensor] = []
        self.grads: List[torch.Tensor] = []
        self.after_step = after_step
        self.side = torch.cuda.Stream(self.device) if after_step is not None else None
        a0, b0 = self.batches[0]
        self.kernel_launches_per_replay = launches_per_step(a0.shape[1], b0.shape[1]) * len(self.batches)
        self._one = torch.ones((), dtype=torch.float32, device=self.device)   # d loss / d loss, made once (no fill per step)
        self.stream = torch.cuda.Stream(self.device)
        self.graph = torch.cuda.CUDAGraph()
        self._capture()

    def _one_step(self, k: int) -> torch.Tensor:
        a, b = self.batches[k]
        a.grad = None
        b.grad = None
        loss = self.crit(a, b)
        loss.backward(gradient=self._one)
        return loss

    def _capture(self) -> None:
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for k in range(min(3, len(self.batches))):       # warm-up on the capture stream (workspace, allocator)
                self._one_step(k)
        self.stream.synchronize()
        with torch.cuda.graph(self.graph, stream=self.stream):
            for k in range(len(self.batches)):
                self.losses.append(self._one_step(k))
                self.grads.append(self.batches[k][0].grad)
                if self.after_step is not None:
                    self.side.wait_stream(self.stream)
                    with torch.cuda.stream(self.side):
                        self.after_step(self.losses[-1])
            if self.after_step is not None:
                self.stream.wait_stream(self.side)
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def replay(self) -> None:
        self.graph.replay()


def pin_pair(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Copies of (a, b) as two views of ONE pinned host buffer, so a step's inputs cross PCIe as a single transfer."""
    flat = torch.empty(a.numel() + b.numel(), dtype=a.dtype).pin_memory()
    va, vb = flat[:a.numel()].view_as(a), flat[a.numel():].view_as(b)
    va.copy_(a)
    vb.copy_(b)
    return va, vb


def _adjacent(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype
            and a.data_ptr() + a.numel() * a.element_size() == b.data_ptr())


class HostChamferStepGraph:
    """S training steps whose inputs live in pinned HOST memory.  Per step, inside the graph: H2D copy of
    (pred, target) into one of two device staging slots (copy stream; ONE transfer when the two host tensors are
    adjacent views of one pinned buffer, see pin_pair), ChamferLoss forward+backward (compute stream), D2H copy of the
    loss into a pinned (S,) result vector (its own stream, so it neither waits behind the next H2D nor holds up the
    next step).  The copy of step k+1 overlaps the kernels of step k.  After replay() + synchronize:
    self.losses_host[k] is step k's loss."""

    def __init__(self, host_batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], device: torch.device,
                 bidirectional: bool = True):
        assert all(a.is_pinned() and b.is_pinned() for a, b in host_batches), "host batches must be pinned"
        self.host = list(host_batches)
        self.device = device
        self.S = len(self.host)
        self.crit = ChamferLoss(bidirectional)
        a0, b0 = self.host[0]
        self.flat = [torch.empty(a0.numel() + b0.numel(), dtype=a0.dtype, device=device) for _ in range(2)]
        self.stage = [(f[:a0.numel()].view_as(a0).requires_grad_(True), f[a0.numel():].view_as(b0)) for f in self.flat]
        self.single_copy = all(_adjacent(a, b) for a, b in self.host)
        self.losses_host = torch.zeros(self.S, dtype=torch.float32).pin_memory()
        self.h2d_bytes_per_step = (a0.numel() + b0.numel()) * 4
        self.d2h_bytes_per_step = 4
        self.kernel_launches_per_replay = launches_per_step(a0.shape[1], b0.shape[1]) * self.S
        self._one = torch.ones((), dtype=torch.float32, device=device)
        self.compute = torch.cuda.Stream(device)
        self.copy = torch.cuda.Stream(device)
        self.readback = torch.cuda.Stream(device)
        self.graph = torch.cuda.CUDAGraph()
        # device loss scalars of the captured steps, kept alive: each is read by a D2H copy on ANOTHER stream, so its
        # block must not go back to the capture pool (where the next step's outputs would overwrite it early)
        self._losses_dev: List[torch.Tensor] = []
        self._capture()

    def _copy_in(self, k: int) -> None:
        ha, hb = self.host[k]
        if self.single_copy:
            n = ha.numel() + hb.numel()
            host_flat = torch.as_strided(ha, (n,), (1,))          # the pinned buffer both views live in
            self.flat[k % 2].copy_(host_flat, non_blocking=True)
        else:
            da, db = self.stage[k % 2]
            da.detach().copy_(ha, non_blocking=True)
            db.copy_(hb, non_blocking=True)

    def _step(self, k: int) -> torch.Tensor:
        da, db = self.stage[k % 2]
        da.grad = None
        loss = self.crit(da, db)
        loss.backward(gradient=self._one)
        return loss.detach().reshape(1)

    def _capture(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        self.compute.wait_stream(cur)
        with torch.cuda.stream(self.compute):          # eager warm-up
            for k in range(min(2, self.S)):
                self._copy_in(k)
                self.losses_host[k:k + 1].copy_(self._step(k), non_blocking=True)
        self.compute.synchronize()
        with torch.cuda.graph(self.graph, stream=self.compute):
            copied = [None] * self.S
            computed = [None] * self.S

            def step_and_read_back(j: int) -> None:
                self.compute.wait_event(copied[j])
                loss = self._step(j)
                self._losses_dev.append(loss)
                computed[j] = torch.cuda.Event()
                computed[j].record(self.compute)
                self.readback.wait_event(computed[j])
                with torch.cuda.stream(self.readback):
                    self.losses_host[j:j + 1].copy_(loss, non_blocking=True)

            for k in range(self.S):
                # copy k may start once step k-2 (last user of this staging slot) has finished
                self.copy.wait_stream(self.compute) if k < 2 else self.copy.wait_event(computed[k - 2])
                with torch.cuda.stream(self.copy):
                    self._copy_in(k)
                    copied[k] = torch.cuda.Event()
                    copied[k].record(self.copy)
                if k >= 1:
                    step_and_read_back(k - 1)          # while copy k is in flight
            step_and_read_back(self.S - 1)
            self.compute.wait_stream(self.copy)
            self.compute.wait_stream(self.readback)
        cur.wait_stream(self.compute)

    def replay(self) -> None:
        self.graph.replay()
