"""Batched RL environment step: E episodes of RLGANNetEnvironment.step in one device-resident pass.

The reference's environment (models/rl_gan_net.py:267-339) plays ONE episode per call: z (1, z_dim) -> latent GAN generator
-> decoder -> completed cloud; the complete cloud's GFV from the encoder; the discriminator's logit for the generated GFV;
RewardFunction (utils/losses.py:209-246) = -(100 CD + 10 MSE(gfv) + 0.01 (-D)); then `reward.item()` (a host sync) and two
`.cpu().numpy()` copies.  train_rl_agent (train_rl_gan_net.py:406-429) loops that over episodes.

Here episode e of a step uses batch item e: the generator, decoder and discriminator (small MLPs, the model's own modules,
stock torch) run once on (E, .) tensors, the encoder runs once over the E complete clouds on this package's kernels, and the
E Chamfer distances are ONE batched launch; states and rewards come back as device tensors -- no `.item()`, no host copies.
With `capture=True` the whole step is one CUDA-graph replay.  All modules are used in eval mode, like the reference's
environment (its B=1 calls could not run BatchNorm in train mode).

  env = BatchedRLEnvironment.from_model(rl_gan_net)        # anything with the reference's RLGANNet methods
  states = env.reset({"incomplete": (E,N,3), "complete": (E,M,3)})          # (E, latent) device tensor
  next_states, rewards, dones, info = env.step(actions)                      # actions (E, z_dim) tensor / ndarray
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from .reward import batched_rewards


class BatchedRLEnvironment:
    def __init__(self, encode: Callable, generate: Callable, decode: Callable, discriminate: Callable, device,
                 w_chamfer: float = 100.0, w_gfv: float = 10.0, w_discriminator: float = 0.01, capture: bool = False):
        self.encode, self.generate, self.decode, self.discriminate = encode, generate, decode, discriminate
        self.device = torch.device(device)
        self.weights = (float(w_chamfer), float(w_gfv), float(w_discriminator))
        self.capture = capture
        self.current_batch: Optional[Dict[str, torch.Tensor]] = None
        self.current_step = 0
        self.target_pc: Optional[torch.Tensor] = None
        self.target_gfv: Optional[torch.Tensor] = None
        self._graph = None
        self._z = None
        self._out = None

    @classmethod
    def from_model(cls, model, capture: bool = False) -> "BatchedRLEnvironment":
        """`model`: the reference's RLGANNet (models/rl_gan_net.py:33) or anything with encode_point_cloud / generate_clean_gfv /
        decode_gfv / latent_gan.discriminate / reward_function.{w_chamfer,w_gfv,w_discriminator} / device."""
        rf = model.reward_function
        return cls(model.encode_point_cloud, model.generate_clean_gfv, model.decode_gfv, model.latent_gan.discriminate,
                   model.device, rf.w_chamfer, rf.w_gfv, rf.w_discriminator, capture)

    def reset(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """rl_gan_net.py:279-297 for every item of the batch: the states are the GFVs of the incomplete clouds, (E, latent).
        The complete clouds and their GFVs (which the reference re-encodes at every step, :318-319) are kept on the device."""
        self.current_batch = batch
        self.current_step = 0
        self._graph = None
        with torch.no_grad():
            incomplete = batch["incomplete"].to(self.device, non_blocking=True)
            self.target_pc = batch["complete"].to(self.device, non_blocking=True).contiguous()
            states = self.encode(incomplete)
            self.target_gfv = self.encode(self.target_pc)
        return states.detach()

    def _step_tensors(self, z: torch.Tensor):
        clean_gfv = self.generate(z)                                           # rl_gan_net.py:311
        completed_pc = self.decode(clean_gfv)                                  # :314
        disc = self.discriminate(clean_gfv)                                    # compute_reward, :196
        rewards = batched_rewards(completed_pc, self.target_pc, clean_gfv, self.target_gfv, disc, *self.weights)   # :199
        return clean_gfv, completed_pc, rewards

    def step(self, actions) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        """rl_gan_net.py:299-339 for E episodes at once.  Returns (next_states (E, latent), rewards (E,), dones (E,) bool,
        info) as device tensors; nothing is synchronised with the host."""
        if self.target_pc is None:
            raise RuntimeError("reset() must be called before step()")
        E = self.target_pc.shape[0]
        z = torch.as_tensor(actions, dtype=torch.float32).to(self.device, non_blocking=True).reshape(E, -1)
        with torch.no_grad():
            if not self.capture:
                clean_gfv, completed_pc, rewards = self._step_tensors(z)
            else:
                if self._graph is None:
                    self._capture(z)
                self._z.copy_(z, non_blocking=True)
                self._graph.replay()
                clean_gfv, completed_pc, rewards = self._out
                # the graph writes into fixed buffers: states and rewards are handed out as copies, the (large) clouds in
                # `info` are views that the next step overwrites
                clean_gfv, rewards = clean_gfv.clone(), rewards.clone()
        self.current_step += 1
        dones = torch.ones(E, dtype=torch.bool, device=self.device)           # every episode is one step (:327)
        info = {"completed_pc": completed_pc, "target_pc": self.target_pc, "clean_gfv": clean_gfv, "target_gfv": self.target_gfv}
        return clean_gfv, rewards, dones, info

    def _capture(self, z: torch.Tensor) -> None:
        self._z = z.clone()
        stream = torch.cuda.Stream(self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(stream):
            for _ in range(2):
                self._step_tensors(self._z)
        stream.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=stream):
            self._out = self._step_tensors(self._z)
        torch.cuda.current_stream(self.device).wait_stream(stream)
