"""gan-rl_3d_b200: the B200-native (sm_100a) hot path of RL-GAN-Net point-cloud completion
(phanich004/GAN-RL_3D) -- batched Chamfer distance and the PointNet encoder trunk + max-pool -- as
torch.autograd.Function drop-ins over a plain C ABI (include/rlg_b200.h, librlg_b200.so).

    import gan_rl_3d_b200 as rlg          # alias module at the repo root (the package dir has a hyphen)
    rlg.install()                         # rebinding the reference's two choke points; its files stay unchanged
    ...
    rlg.uninstall()

Standalone use (no reference checkout needed): rlg.chamfer_distance_l2 / chamfer_distance / ChamferLoss and
rlg.PointNetEncoder mirror the reference's names, arguments and return shapes.
"""
from __future__ import annotations

import importlib
import sys

import torch
from typing import Optional

from . import _lib
from .chamfer import (ChamferFn, ChamferLoss, ChamferLossFn, chamfer_backward, chamfer_distance, chamfer_distance_l2,
                      chamfer_nearest, deterministic_backward, get_default_sweep, get_reserved_sms, set_default_sweep,
                      set_deterministic_backward, set_reserved_sms)
from .chamfer import is_hot_path_input as _chamfer_hot
from .reward import RewardFunction, batched_rewards
from .encoder import (EncoderTrunkFn, PointNetEncoder, encoder_path_of, encoder_pool, encoder_pool_gemm, fold_trunk,
                      folded_trunk_cached, fused_forward, get_encoder_precision, pack_bf16, pack_gemm,
                      packed_trunk_cached, resolve_path, set_encoder_precision, set_train_path)
from .train import EncoderTrainFn, train_supported, trunk_pool_autograd
from .ae_step import AEStepGraph, PointCloudAutoencoder, PointNetDecoder
from . import train
from .environment import BatchedRLEnvironment
from .data import DeviceBatcher, build_cache, draw_plan

__all__ = ["install", "uninstall", "is_installed", "chamfer_distance_l2", "chamfer_distance", "ChamferLoss",
           "ChamferFn", "ChamferLossFn", "chamfer_nearest", "chamfer_backward", "set_default_sweep", "get_default_sweep", "set_reserved_sms", "get_reserved_sms", "set_deterministic_backward", "deterministic_backward", "PointNetEncoder", "EncoderTrunkFn", "encoder_pool",
           "fold_trunk", "folded_trunk_cached", "fused_forward", "set_encoder_precision", "get_encoder_precision",
           "pack_bf16", "pack_gemm", "packed_trunk_cached", "encoder_pool_gemm", "encoder_path_of", "resolve_path", "RewardFunction", "batched_rewards", "library_path", "abi_version", "set_nvtx",
           "EncoderTrainFn", "train_supported", "trunk_pool_autograd", "set_train_path",
           "AEStepGraph", "PointCloudAutoencoder", "PointNetDecoder", "BatchedRLEnvironment", "DeviceBatcher", "build_cache", "draw_plan"]

_installed = {}


def set_nvtx(enabled: bool) -> None:
    """NVTX ranges (rlg.chamfer_fwd, rlg.chamfer_bwd, rlg.encoder_*, rlg.batch_prepare) around the library calls, for nsys /
    ncu --nvtx timelines; also RLG_NVTX=1."""
    _lib.set_nvtx(enabled)


def library_path() -> str:
    return _lib.LIB_PATH


def abi_version() -> int:
    return int(_lib.load().rlg_version())


def _find(modname: str):
    if modname in sys.modules:
        return sys.modules[modname]
    try:
        return importlib.import_module(modname)
    except Exception:
        return None


def install(losses_module=None, autoencoder_module=None, check_finite: bool = False) -> dict:
    """Rebind the reference's two choke points:

      utils.losses.chamfer_distance_l2             (utils/losses.py:13; resolved by global lookup from
                                                    chamfer_distance :54, so ChamferLoss, RewardFunction, the
                                                    trainer and the RL environment all follow)
      models.autoencoder.PointNetEncoder.forward   (models/autoencoder.py:56; class attribute, so the module
                                                    tree, parameters and state_dict keys stay intact)

    Pass the module objects explicitly, or let them be found as `utils.losses` / `models.autoencoder`.
    Inputs outside the CUDA contract (CPU tensors, fp64, 4-D broadcast input of validate_joint, train-mode
    BatchNorm) are handed to the saved original function, so reference behaviour -- including its errors --
    is preserved for everything that is not the hot path.  Loads the CUDA library eagerly: a missing
    librlg_b200.so is an ImportError here, not a silent fallback later.

    check_finite=True: clouds containing NaN/Inf are handed to the original function too (the reference lets a NaN
    candidate win torch.min; the kernels' filter may skip it).  The check is a device reduction and a host
    synchronisation per call, so it is off by default."""
    _lib.load()
    losses_module = losses_module or _find("utils.losses")
    autoencoder_module = autoencoder_module or _find("models.autoencoder")
    done = {}
    if losses_module is not None and "losses" not in _installed:
        original = losses_module.chamfer_distance_l2

        def chamfer_distance_l2_b200(pc1, pc2):
            if _chamfer_hot(pc1, pc2):
                if check_finite and not bool((torch.isfinite(pc1).all() & torch.isfinite(pc2).all()).item()):
                    return original(pc1, pc2)
                return ChamferFn.apply(pc1, pc2)
            return original(pc1, pc2)

        chamfer_distance_l2_b200.__wrapped__ = original
        losses_module.chamfer_distance_l2 = chamfer_distance_l2_b200
        _installed["losses"] = (losses_module, original)
        done["chamfer_distance_l2"] = losses_module.__name__
    if autoencoder_module is not None and "encoder" not in _installed:
        cls = autoencoder_module.PointNetEncoder
        original_fwd = cls.forward

        def forward_b200(self, x):
            return fused_forward(self, x, _original=original_fwd)

        forward_b200.__wrapped__ = original_fwd
        cls.forward = forward_b200
        _installed["encoder"] = (cls, original_fwd)
        done["PointNetEncoder.forward"] = autoencoder_module.__name__
    return done


def uninstall() -> None:
    if "losses" in _installed:
        mod, original = _installed.pop("losses")
        mod.chamfer_distance_l2 = original
    if "encoder" in _installed:
        cls, original = _installed.pop("encoder")
        cls.forward = original


def is_installed() -> bool:
    return bool(_installed)
