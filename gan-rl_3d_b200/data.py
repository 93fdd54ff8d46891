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
    """Rz @ Ry @ Rx of utils/data_utils.py:74-99 for the three angles theta; theta may carry leading batch dimensions
    (..., 3) -> (..., 3, 3)."""
    th = np.asarray(theta, np.float64)
    c, s = np.cos(th), np.sin(th)
    cx, cy, cz, sx, sy, sz = c[..., 0], c[..., 1], c[..., 2], s[..., 0], s[..., 1], s[..., 2]
    R = np.empty(th.shape[:-1] + (3, 3))
    R[..., 0, 0] = cz * cy; R[..., 0, 1] = cz * sy * sx - sz * cx; R[..., 0, 2] = cz * sy * cx + sz * sx
    R[..., 1, 0] = sz * cy; R[..., 1, 1] = sz * sy * sx + cz * cx; R[..., 1, 2] = sz * sy * cx - cz * sx
    R[..., 2, 0] = -sy;     R[..., 2, 1] = cy * sx;                R[..., 2, 2] = cy * cx
    return R


def draw_plan(rng: np.random.Generator, B: int, N: int, items: Optional[np.ndarray] = None, augment: bool = True,
              jitter_sigma: float = 0.01, jitter_clip: float = 0.05, host_jitter: bool = False,
              host_indices: bool = True) -> Dict[str, np.ndarray]:
    """The random decisions of one batch as host arrays (see struct rlg_prepare_plan), drawn with the reference's
    distributions: utils/dataset.py:255-266 (removal), :284-294 (augmentation, independently for the complete and the
    incomplete cloud), :408 (padding indices).  Every per-cloud scalar is drawn for the whole batch at once (no Python loop
    per cloud; the reference's stream of draws is not reproduced, its distributions are).  The jitter noise itself
    (2 x B x N x 3 normals, the bulk of the draws) is made on the device by DeviceBatcher from the per-cloud on/off flags
    `jitter_on` unless host_jitter=True puts the clipped noise into the plan (tests feed the same noise to the restated
    reference that way).  host_indices=False leaves the two bulk index draws -- the random subsets' permutations (:260) and
    the padding indices (:408) -- to the device as well (torch's CUDA generator: argsort of uniforms, randint), so the host
    draws a handful of scalars per cloud."""
    ratio = rng.uniform(0.2, 0.5, B)
    random_subset = rng.random(B) < 0.5                                   # else: the points nearest to a random centre go
    vi = (N - 1) * np.true_divide(ratio * 100, 100.0)                      # numpy's percentile, method 'linear'
    plan = {"item": (np.arange(B) if items is None else np.asarray(items)).astype(np.int32),
            "method": np.where(random_subset, 0, 1).astype(np.int32),
            "n_keep": np.where(random_subset, (N * (1 - ratio)).astype(np.int64), 0).astype(np.int32),
            "keep_idx": np.zeros((B, N), np.int32) if host_indices else None,
            "center": np.where(random_subset, 0, rng.integers(0, N, B)).astype(np.int32),
            "q_index": np.where(random_subset, 0, np.floor(vi)).astype(np.int32),
            "q_gamma": np.where(random_subset, 0.0, vi - np.floor(vi)),
            "ratio": ratio,
            "pad_idx": rng.integers(0, 2 ** 31 - 1, (B, N), dtype=np.int64).astype(np.int32) if host_indices else None}
    if host_indices:
        for b in np.flatnonzero(random_subset):
            k = int(plan["n_keep"][b])
            plan["keep_idx"][b, :k] = rng.choice(N, k, replace=False)
    if augment:
        rotate = rng.random((2, B)) < 0.5
        jitter_on = rng.random((2, B)) < 0.5
        rescale = rng.random((2, B)) < 0.3
        R = rotation_matrix(rng.uniform(0, 2 * np.pi, (2, B, 3)))
        rot = np.where(rotate[..., None, None], R, np.eye(3)).astype(np.float32).reshape(2, B, 9)
        scale = np.where(rescale, rng.uniform(0.8, 1.2, (2, B)), 1.0).astype(np.float32)
        plan.update(rot=rot, scale=scale, jitter_on=jitter_on, jitter_sigma=jitter_sigma, jitter_clip=jitter_clip)
        if host_jitter:
            noise = np.clip(rng.normal(0.0, jitter_sigma, (2, B, N, 3)), -jitter_clip, jitter_clip)
            plan["jitter"] = np.where(jitter_on[..., None, None], noise, 0.0).astype(np.float32)
    return plan


# per-cloud scalars of struct rlg_prepare_plan, packed into one upload by DeviceBatcher.make_batch
_PLAN_SCALARS = (("item", np.int32), ("method", np.int32), ("n_keep", np.int32), ("center", np.int32), ("q_index", np.int32),
                 ("q_gamma", np.float64), ("rot", np.float32), ("scale", np.float32))


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
        self._stage: Optional[torch.Tensor] = None

    def _staging(self, nbytes: int) -> torch.Tensor:
        """Pinned host buffer for the plan's scalars.  Reuse is safe: make_batch ends with a read-back of the padded length,
        so the previous batch's transfer out of this buffer has completed."""
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes, 4096), dtype=torch.uint8, pin_memory=True)
        return self._stage

    def make_batch(self, plan: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
        lib = _lib.load()
        B, N, dev = int(len(plan["method"])), self.N, self.device
        if B and (int(np.max(plan["item"])) >= self.items or int(np.min(plan["item"])) < 0):
            raise IndexError("plan['item'] outside the cache")
        keep = {}
        cp = _lib.RlgPreparePlan()
        # the per-cloud scalars of the plan (a few KB) travel as ONE transfer from a pinned staging buffer; the optional bulk
        # arrays of a host-drawn plan (subset / padding indices, jitter noise) as one transfer each
        small = [(name, np.ascontiguousarray(plan[name], dtype=dtype)) for name, dtype in _PLAN_SCALARS if plan.get(name) is not None]
        device_jitter = plan.get("jitter") is None and plan.get("jitter_on") is not None and bool(np.any(plan["jitter_on"]))
        if device_jitter:
            small.append(("jitter_on", np.ascontiguousarray(plan["jitter_on"], dtype=np.float32)))     # not a field of the C struct
        offsets, total = {}, 0
        for name, arr in small:
            offsets[name] = total
            total += (arr.nbytes + 15) & ~15
        if total:
            stage = self._staging(total)
            view = stage.numpy()
            for name, arr in small:
                view[offsets[name]: offsets[name] + arr.nbytes] = arr.reshape(-1).view(np.uint8)
            packed = stage[:total].to(dev, non_blocking=True)
            keep["plan_scalars"] = packed
            for name, _ in small:
                if name != "jitter_on":
                    setattr(cp, name, packed.data_ptr() + offsets[name])
        if plan.get("keep_idx") is None:                   # a uniform random permutation per cloud: the first n_keep are the subset
            perm = torch.rand((B, N), device=dev).argsort(dim=1).to(torch.int32)
            keep["keep_idx"] = perm
            cp.keep_idx = perm.data_ptr()
        if plan.get("pad_idx") is None:
            pad = torch.randint(0, 2 ** 31 - 1, (B, N), device=dev, dtype=torch.int32)
            keep["pad_idx"] = pad
            cp.pad_idx = pad.data_ptr()
        if device_jitter:
            # the jitter noise of utils/data_utils.py:140-142, drawn on the device (torch's CUDA generator)
            on = packed[offsets["jitter_on"]: offsets["jitter_on"] + 8 * B].view(torch.float32).view(2, B)
            noise = torch.randn((2, B, N, 3), dtype=torch.float32, device=dev).mul_(float(plan["jitter_sigma"]))
            noise.clamp_(-float(plan["jitter_clip"]), float(plan["jitter_clip"])).mul_(on[:, :, None, None])
            keep["jitter"] = noise
            cp.jitter = noise.data_ptr()
        for name in ("keep_idx", "pad_idx", "jitter"):
            if plan.get(name) is None:
                continue
            t = torch.as_tensor(np.ascontiguousarray(plan[name], dtype=np.float32 if name == "jitter" else np.int32)).to(dev, non_blocking=True)
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
