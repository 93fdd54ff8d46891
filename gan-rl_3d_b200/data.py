"""Input pipeline on the device: a binary cache of the complete clouds + GPU-side incomplete-cloud creation,
augmentation, normalisation and duplicate-padding, behind rlg_batch_prepare (csrc/batch_prepare.cu).

The reference's loader (utils/dataset.py) parses a text file per sample with np.loadtxt (:234-236), makes the incomplete
cloud, augments and normalises in Python/numpy per sample (:135-187,252-297), pads with a Python loop in the collate function
(:393-421) and disables pinned memory (:440); its own log shows ~50 s per 1000-cloud epoch, all of it here.

  build_cache(files_or_arrays)     parse once -> (items, N, 3) float32 array (np.save it; np.load(mmap_mode="r") later)
  draw_plan(rng, B, N, ...)        the random decisions of one batch, with the reference's distributions
                                   (removal ratio U(0.2, 0.5), method 50/50, rotation / jitter / scale with p = .5/.5/.3)
  DeviceBatcher(cache, device)     keeps the cache in HBM (800 clouds = 19.7 MB) and turns a plan into
                                   {'complete_pc': (B,N,3), 'incomplete_pc': (B,max_len,3)} like shapenet_collate_fn
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib


def build_cache(sources: Iterable, num_points: int = 2048, seed: int = 0) -> np.ndarray:
    """Complete clouds -> one (items, num_points, 3) float32 array.  `sources`: paths of whitespace-separated text files
    (x y z ... per line, utils/dataset.py:234-236) or arrays (n, >=3).  Clouds with more / fewer points are resampled /
    padded with repeated points once, here (the reference redraws that per epoch, utils/dataset.py:150-158)."""
    rng = np.random.default_rng(seed)
    out = []
    for src in sources:
        pc = np.loadtxt(src)[:, :3] if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__") else np.asarray(src)[:, :3]
        n = len(pc)
        if n > num_points:
            pc = pc[rng.choice(n, num_points, replace=False)]
        elif n < num_points:
            pc = np.concatenate([pc, pc[rng.choice(n, num_points - n)]], axis=0)
        out.append(pc.astype(np.float32))
    return np.stack(out) if out else np.zeros((0, num_points, 3), np.float32)


def rotation_matrix(theta: Sequence[float]) -> np.ndarray:
    """Rz @ Ry @ Rx of utils/data_utils.py:74-99 for the three angles theta."""
    cx, sx, cy, sy, cz, sz = np.cos(theta[0]), np.sin(theta[0]), np.cos(theta[1]), np.sin(theta[1]), np.cos(theta[2]), np.sin(theta[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def draw_plan(rng: np.random.Generator, B: int, N: int, items: Optional[np.ndarray] = None, augment: bool = True,
              jitter_sigma: float = 0.01, jitter_clip: float = 0.05, host_jitter: bool = False,
              host_indices: bool = True) -> Dict[str, np.ndarray]:
    """The random decisions of one batch as host arrays (see struct rlg_prepare_plan), drawn with the reference's
    distributions: utils/dataset.py:255-266 (removal), :284-294 (augmentation, independently for the complete and the
    incomplete cloud), :408 (padding indices).  The jitter noise itself (2 x B x N x 3 normals, the bulk of the draws) is
    made on the device by DeviceBatcher from the per-cloud on/off flags `jitter_on` unless host_jitter=True puts the
    clipped noise into the plan (tests feed the same noise to the restated reference that way).  host_indices=False leaves
    the two bulk index draws -- the random subsets' permutations (:260) and the padding indices (:408) -- to the device as
    well (torch's CUDA generator: argsort of uniforms, randint), so the host draws a handful of scalars per cloud."""
    plan = {"item": (np.arange(B) if items is None else np.asarray(items)).astype(np.int32),
            "method": np.zeros(B, np.int32), "n_keep": np.zeros(B, np.int32), "keep_idx": np.zeros((B, N), np.int32),
            "center": np.zeros(B, np.int32), "q_index": np.zeros(B, np.int32), "q_gamma": np.zeros(B, np.float64),
            "ratio": np.zeros(B, np.float64),
            "pad_idx": rng.integers(0, 2 ** 31 - 1, (B, N), dtype=np.int64).astype(np.int32) if host_indices else None}
    if not host_indices:
        plan["keep_idx"] = None
    for b in range(B):
        ratio = rng.uniform(0.2, 0.5)
        plan["ratio"][b] = ratio
        if rng.random() < 0.5:
            n_keep = int(N * (1 - ratio))
            plan["n_keep"][b] = n_keep
            if host_indices:
                plan["keep_idx"][b, :n_keep] = rng.choice(N, n_keep, replace=False)
        else:
            plan["method"][b] = 1
            plan["center"][b] = rng.integers(N)
            vi = (N - 1) * np.true_divide(ratio * 100, 100.0)           # numpy's percentile, method 'linear'
            plan["q_index"][b] = int(np.floor(vi))
            plan["q_gamma"][b] = vi - np.floor(vi)
    if augment:
        rot = np.tile(np.eye(3, dtype=np.float32).reshape(1, 1, 9), (2, B, 1))
        scale = np.ones((2, B), np.float32)
        jitter = np.zeros((2, B, N, 3), np.float32) if host_jitter else None
        jitter_on = np.zeros((2, B), np.bool_)
        for which in range(2):
            for b in range(B):
                if rng.random() < 0.5:
                    rot[which, b] = rotation_matrix(rng.uniform(0, 2 * np.pi, 3)).astype(np.float32).reshape(9)
                if rng.random() < 0.5:
                    jitter_on[which, b] = True
                    if host_jitter:
                        jitter[which, b] = np.clip(rng.normal(0.0, jitter_sigma, (N, 3)), -jitter_clip, jitter_clip)
                if rng.random() < 0.3:
                    scale[which, b] = rng.uniform(0.8, 1.2)
        plan.update(rot=rot, scale=scale, jitter_on=jitter_on, jitter_sigma=jitter_sigma, jitter_clip=jitter_clip)
        if host_jitter:
            plan["jitter"] = jitter
    return plan


class DeviceBatcher:
    """The cache lives on the device; `make_batch(plan)` uploads the plan (a few hundred KB of draws), runs the two kernels of
    rlg_batch_prepare and returns {'complete_pc', 'incomplete_pc', 'lengths'} as device tensors.  One 4-byte read-back per
    batch (the padded length is data dependent, as in the reference's collate function)."""

    def __init__(self, cache: np.ndarray, device: torch.device):
        if cache.ndim != 3 or cache.shape[2] != 3:
            raise ValueError("cache must be (items, N, 3)")
        self.device = torch.device(device)
        self.cache = torch.as_tensor(np.ascontiguousarray(cache, dtype=np.float32)).to(self.device)
        self.items, self.N = int(cache.shape[0]), int(cache.shape[1])

    def make_batch(self, plan: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
        lib = _lib.load()
        B, N, dev = int(len(plan["method"])), self.N, self.device
        if B and (int(np.max(plan["item"])) >= self.items or int(np.min(plan["item"])) < 0):
            raise IndexError("plan['item'] outside the cache")
        keep = {}
        cp = _lib.RlgPreparePlan()
        if plan.get("keep_idx") is None:                   # a uniform random permutation per cloud: the first n_keep are the subset
            perm = torch.rand((B, N), device=dev).argsort(dim=1).to(torch.int32)
            keep["keep_idx"] = perm
            cp.keep_idx = perm.data_ptr()
        if plan.get("pad_idx") is None:
            pad = torch.randint(0, 2 ** 31 - 1, (B, N), device=dev, dtype=torch.int32)
            keep["pad_idx"] = pad
            cp.pad_idx = pad.data_ptr()
        if plan.get("jitter") is None and plan.get("jitter_on") is not None and bool(np.any(plan["jitter_on"])):
            # the jitter noise of utils/data_utils.py:140-142, drawn on the device (torch's CUDA generator)
            on = torch.as_tensor(np.ascontiguousarray(plan["jitter_on"])).to(dev, non_blocking=True)
            noise = torch.randn((2, B, N, 3), dtype=torch.float32, device=dev).mul_(float(plan["jitter_sigma"]))
            noise.clamp_(-float(plan["jitter_clip"]), float(plan["jitter_clip"])).mul_(on[:, :, None, None])
            keep["jitter"] = noise
            cp.jitter = noise.data_ptr()
        for name, dtype in (("item", np.int32), ("method", np.int32), ("n_keep", np.int32), ("keep_idx", np.int32),
                            ("center", np.int32), ("q_index", np.int32), ("q_gamma", np.float64), ("rot", np.float32),
                            ("scale", np.float32), ("jitter", np.float32), ("pad_idx", np.int32)):
            if plan.get(name) is None:
                continue
            t = torch.as_tensor(np.ascontiguousarray(plan[name], dtype=dtype)).to(dev, non_blocking=True)
            keep[name] = t
            setattr(cp, name, t.data_ptr())
        complete = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        incomplete = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        lengths = torch.empty(B, dtype=torch.int32, device=dev)
        max_len = torch.zeros(1, dtype=torch.int32, device=dev)
        if B == 0:
            return {"complete_pc": complete, "incomplete_pc": incomplete[:, :0], "lengths": lengths}
        with torch.cuda.device(dev):
            _lib.nvtx_push("rlg.batch_prepare")
            rc = lib.rlg_batch_prepare(self.cache.data_ptr(), self.items, N, B, ctypes.byref(cp), complete.data_ptr(),
                                       incomplete.data_ptr(), lengths.data_ptr(), max_len.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream)
            _lib.nvtx_pop()
            _lib.check("rlg_batch_prepare", rc)
        m = int(max_len.item())
        # "draws": what was uploaded or drawn on the device (keep_idx, pad_idx, jitter, ...), for inspection / tests
        return {"complete_pc": complete, "incomplete_pc": incomplete[:, :m].contiguous(), "lengths": lengths, "draws": keep}
