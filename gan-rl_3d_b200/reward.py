"""Batched RL reward evaluation: the reference's RewardFunction (utils/losses.py:209-246) evaluated
for E episodes in one device-resident pass.

The reference's environment (models/rl_gan_net.py:303-339) runs one episode at a time: B=1 Chamfer of the decoder's
output against the complete cloud, an MSE between two GFVs, one discriminator logit, and a `.item()` host sync per
step.  Here E episodes are E cloud pairs of ONE batched Chamfer forward (no gradient, no (E,N,M) matrix) and the
reward comes back as an (E,) tensor with no host synchronisation:

    reward[e] = -( w_chamfer * CD[e] + w_gfv * mean_k (pred_gfv[e,k] - target_gfv[e,k])^2 + w_discriminator * (-D[e]) )

which is exactly what `RewardFunction.compute_reward` returns when called with the episode's B=1 tensors
(ChamferLoss = mean over the 1-pair batch of (dist1+dist2)/2, F.mse_loss over the (1,G) GFVs, -mean of the one logit).
"""
from __future__ import annotations

import torch

from .chamfer import ChamferLoss, chamfer_distance


def batched_rewards(pred_pc: torch.Tensor, target_pc: torch.Tensor, pred_gfv: torch.Tensor, target_gfv: torch.Tensor,
                    discriminator_output: torch.Tensor, w_chamfer: float = 100.0, w_gfv: float = 10.0,
                    w_discriminator: float = 0.01) -> torch.Tensor:
    """(E,N,3), (E,M,3), (E,G), (E,G), (E,) or (E,1)  ->  (E,) rewards, one per episode (see the module docstring)."""
    with torch.no_grad():
        cd = chamfer_distance(pred_pc, target_pc, bidirectional=True)                       # (E,)
        gfv = (pred_gfv.float() - target_gfv.float()).pow(2).mean(dim=1)                    # F.mse_loss per episode
        disc = -discriminator_output.float().reshape(discriminator_output.shape[0], -1).mean(dim=1)
        return -(w_chamfer * cd + w_gfv * gfv + w_discriminator * disc)


class RewardFunction:
    """Mirror of the reference class (same constructor, same `compute_reward` semantics: ONE scalar for the batch it is
    given), plus `compute_rewards` for E episodes at once."""

    def __init__(self, w_chamfer: float = 100.0, w_gfv: float = 10.0, w_discriminator: float = 0.01):
        self.w_chamfer = w_chamfer
        self.w_gfv = w_gfv
        self.w_discriminator = w_discriminator
        self.chamfer_loss = ChamferLoss()

    def compute_reward(self, pred_pc, target_pc, pred_gfv, target_gfv, discriminator_output) -> torch.Tensor:
        chamfer = self.chamfer_loss(pred_pc, target_pc)
        gfv = torch.nn.functional.mse_loss(pred_gfv, target_gfv)
        disc = -torch.mean(discriminator_output)
        return -(self.w_chamfer * chamfer + self.w_gfv * gfv + self.w_discriminator * disc)

    def compute_rewards(self, pred_pc, target_pc, pred_gfv, target_gfv, discriminator_output) -> torch.Tensor:
        return batched_rewards(pred_pc, target_pc, pred_gfv, target_gfv, discriminator_output, self.w_chamfer,
                               self.w_gfv, self.w_discriminator)
