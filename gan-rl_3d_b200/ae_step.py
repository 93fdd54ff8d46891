"""The autoencoder training step of phanich004/GAN-RL_3D on B200, as a CUDA graph.

train_rl_gan_net.py:220-249 runs, per batch:  optimizer.zero_grad(); recon = model(incomplete);
loss = ChamferLoss()(recon, complete); loss.backward(); optimizer.step()  with Adam (lr 1e-3, weight_decay 1e-5,
train_rl_gan_net.py:173-177) on models/autoencoder.py's PointCloudAutoencoder in train mode.  Here the encoder trunk
(forward and backward, BatchNorm batch statistics included) and the Chamfer loss (forward and backward) are the B200
kernels of this package; the tiny global MLP, the decoder MLP (3 GEMMs on (B,128..6144)) and Adam stay stock torch.
A step is ~60 kernel launches of a few microseconds each, so S steps are captured into one CUDA graph; with more than one
rank the gradient all-reduce (7.15 MB for the reference's dims, as two flat NCCL buckets: the decoder's 6.7 MB on a side stream
under the encoder's backward, the encoder's 0.45 MB after it) is captured before optimizer.step() -- every step, inside the
timed region.

  PointNetDecoder / PointCloudAutoencoder   same constructor arguments, module tree and state_dict keys as the reference
                                            (models/autoencoder.py:79-171); the encoder is this package's PointNetEncoder
  AEStepGraph                               S captured training steps over S device-resident batches
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .chamfer import ChamferLoss
from .encoder import PointNetEncoder


class PointNetDecoder(nn.Module):
    """GFV (B, latent) -> cloud (B, num_points, 3): Linear+BatchNorm1d+ReLU blocks, then a Linear (autoencoder.py:79-129)."""

    def __init__(self, latent_dim: int = 128, num_points: int = 2048, hidden_dims: Optional[List[int]] = None):
        super().__init__()
        hidden_dims = [256, 256, 6144] if hidden_dims is None else list(hidden_dims)
        if hidden_dims[-1] != num_points * 3:
            raise AssertionError(f"Last hidden dim should be {num_points * 3}, got {hidden_dims[-1]}")
        self.latent_dim, self.num_points, self.hidden_dims = latent_dim, num_points, hidden_dims
        layers: List[nn.Module] = []
        c_in = latent_dim
        for c in hidden_dims[:-1]:
            layers += [nn.Linear(c_in, c), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
            c_in = c
        layers.append(nn.Linear(c_in, hidden_dims[-1]))
        self.mlp = nn.Sequential(*layers)

    def forward(self, gfv: torch.Tensor) -> torch.Tensor:
        return self.mlp(gfv).view(-1, self.num_points, 3)


class PointCloudAutoencoder(nn.Module):
    """encoder + decoder with the reference's names (autoencoder.py:132-171)."""

    def __init__(self, input_dim: int = 3, latent_dim: int = 128, num_points: int = 2048,
                 encoder_dims: Optional[List[int]] = None, decoder_dims: Optional[List[int]] = None):
        super().__init__()
        self.input_dim, self.latent_dim, self.num_points = input_dim, latent_dim, num_points
        self.encoder = PointNetEncoder(input_dim, latent_dim, encoder_dims)
        self.decoder = PointNetDecoder(latent_dim, num_points, decoder_dims)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.encoder(x)

    def decode(self, gfv: torch.Tensor) -> torch.Tensor:
        return self.decoder(gfv)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        gfv = self.encode(x)
        return self.decode(gfv), gfv


class AEStepGraph:
    """S training steps of `model` (PointCloudAutoencoder-like: model(x) -> (reconstruction, gfv)) captured in one CUDA
    graph:  zero_grad -> forward -> loss_fn(reconstruction, complete) -> backward -> [all-reduce of the gradients over
    `group`] -> optimizer.step().  The optimizer must be capturable (torch.optim.Adam(..., capturable=True)).
    After replay(): self.losses[k] is step k's (local) loss."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer,
                 batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], loss_fn=None, world: int = 1, group=None,
                 warmup_steps: int = 3):
        assert len(batches) > 0 and batches[0][0].is_cuda
        self.model, self.opt, self.batches = model, optimizer, list(batches)
        self.loss_fn = loss_fn if loss_fn is not None else ChamferLoss()
        self.world, self.group = world, group
        self.device = batches[0][0].device
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.grad_bytes = sum(p.numel() * p.element_size() for p in self.params)
        self.losses: List[torch.Tensor] = []
        self.stream = torch.cuda.Stream(self.device)
        self.graph = torch.cuda.CUDAGraph()
        self._one = torch.ones((), dtype=torch.float32, device=self.device)
        self._flat = torch.empty(sum(p.numel() for p in self.params), dtype=torch.float32, device=self.device)
        self.side = torch.cuda.Stream(self.device) if world > 1 else None
        self.early_bytes = 0                               # gradient bytes all-reduced under the encoder's backward
        self._capture(warmup_steps)

    def _allreduce(self, params, flat: torch.Tensor) -> None:
        """Average the gradients of `params` over the ranks through the flat bucket `flat` (one NCCL all-reduce)."""
        import torch.distributed as dist
        views, off = [], 0
        for p in params:
            n = p.numel()
            views.append(flat[off:off + n].view_as(p))
            off += n
        grads = [p.grad for p in params]
        torch._foreach_copy_(views, grads)
        dist.all_reduce(flat[:off], op=dist.ReduceOp.SUM, group=self.group)
        flat[:off].mul_(1.0 / self.world)                  # every rank's loss is its own batch mean
        torch._foreach_copy_(grads, views)

    def step(self, k: int) -> torch.Tensor:
        x, y = self.batches[k]
        self.opt.zero_grad(set_to_none=True)
        recon, gfv = self.model(x)
        loss = self.loss_fn(recon, y)
        early: List[torch.nn.Parameter] = []
        if self.world > 1 and torch.is_tensor(gfv) and gfv.requires_grad:
            # The backward reaches the decoder first (94 % of the reference's parameters).  When the gradient of the GFV is
            # ready every decoder gradient is final: their bucket is all-reduced on a side stream while the encoder's
            # backward (the long part) runs; only the encoder's small bucket is left for afterwards.
            def gfv_ready(_g):
                early.extend(p for p in self.params if p.grad is not None)
                n = sum(p.numel() for p in early)
                main = torch.cuda.current_stream(self.device)
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    self._allreduce(early, self._flat[:n])
                self.early_bytes = 4 * n
            gfv.register_hook(gfv_ready)
        loss.backward(gradient=self._one)
        if self.world > 1:
            done = {id(p) for p in early}
            late = [p for p in self.params if id(p) not in done and p.grad is not None]
            n0 = sum(p.numel() for p in early)
            if late:
                self._allreduce(late, self._flat[n0:])
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        self.opt.step()
        return loss.detach()

    def _capture(self, warmup_steps: int) -> None:
        import copy
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            # eager warm-up on the capture stream (workspaces, cuBLAS handle, Adam state, NCCL), leaving no trace: the
            # model, its BatchNorm buffers and the optimizer are put back to where they were
            model_state = copy.deepcopy(self.model.state_dict())
            for k in range(max(1, warmup_steps)):
                self.step(k % len(self.batches))
            opt_state = copy.deepcopy(self.opt.state_dict())
            with torch.no_grad():
                self.model.load_state_dict(model_state)
                for st in opt_state["state"].values():     # Adam: step count and both moments back to zero
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
                self.opt.load_state_dict(opt_state)
            self.opt.zero_grad(set_to_none=True)
        self.stream.synchronize()
        with torch.cuda.graph(self.graph, stream=self.stream):
            for k in range(len(self.batches)):
                self.losses.append(self.step(k))
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def replay(self) -> None:
        self.graph.replay()
