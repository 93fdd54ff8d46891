"""PointNet encoder trunk on B200: fused per-point MLP + global max-pool behind the C ABI.

Mirrors models/autoencoder.py:13-76 of phanich004/GAN-RL_3D.  `fused_forward` has the signature of
`PointNetEncoder.forward(self, x (B,N,3)) -> (B, latent_dim)` and works on any module with the reference's
layout (`point_mlp` = [Conv1d(k=1), BatchNorm1d, ReLU] x L, `global_mlp` = [Linear, BatchNorm1d, ReLU]), so
rebinding the class attribute keeps parameters, buffers and state_dict keys untouched.

Eval mode only: BatchNorm with running statistics is an affine map and is folded into the conv weights
(cached; invalidated when any parameter/buffer changes).  In train mode BatchNorm needs batch statistics over
B*N points and updates its buffers, so the stock submodules run instead.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


# ---- BatchNorm folding ------------------------------------------------------------------------------
def _trunk_layers(point_mlp: nn.Sequential):
    mods = list(point_mlp)
    if len(mods) % 3 != 0 or not mods:
        raise ValueError("point_mlp is not [Conv1d, BatchNorm1d, ReLU] x L")
    out = []
    for k in range(0, len(mods), 3):
        conv, bn, act = mods[k], mods[k + 1], mods[k + 2]
        if not (isinstance(conv, nn.Conv1d) and conv.kernel_size == (1,) and conv.stride == (1,)
                and conv.padding == (0,) and conv.groups == 1
                and isinstance(bn, nn.BatchNorm1d) and isinstance(act, nn.ReLU)):
            raise ValueError("point_mlp is not [Conv1d(k=1), BatchNorm1d, ReLU] x L")
        out.append((conv, bn))
    return out


def fold_trunk(point_mlp: nn.Sequential) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Fold every Conv1d(k=1)+BatchNorm1d(eval) of the trunk into (W' (c_out,c_in), b' (c_out)) fp32:
    W' = W * gamma/sqrt(var+eps), b' = (b - mean) * gamma/sqrt(var+eps) + beta  (computed in float64)."""
    folded = []
    with torch.no_grad():
        for conv, bn in _trunk_layers(point_mlp):
            w = conv.weight.detach().double().squeeze(-1)
            b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64,
                                                                                     device=w.device)
            if bn.track_running_stats and bn.running_mean is not None:
                mean, var = bn.running_mean.double(), bn.running_var.double()
            else:
                raise ValueError("BatchNorm1d without running statistics cannot be folded")
            gamma = bn.weight.detach().double() if bn.affine else torch.ones_like(mean)
            beta = bn.bias.detach().double() if bn.affine else torch.zeros_like(mean)
            scale = gamma / torch.sqrt(var + bn.eps)
            folded.append(((w * scale[:, None]).float().contiguous(), ((b - mean) * scale + beta).float().contiguous()))
    return folded


def _trunk_state_version(point_mlp: nn.Sequential) -> tuple:
    sig = []
    for t in list(point_mlp.parameters()) + list(point_mlp.buffers()):
        sig.append((t.data_ptr(), t._version, t.device.index))
    return tuple(sig)


def folded_trunk_cached(module: nn.Module) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Folded weights are a derived cache, rebuilt when a parameter or buffer was modified in place
    (optimizer step, load_state_dict, .to()) -- tracked through tensor versions and storage pointers."""
    ver = _trunk_state_version(module.point_mlp)
    cache = module.__dict__.get("_rlg_folded")
    if cache is None or cache[0] != ver:
        cache = (ver, fold_trunk(module.point_mlp))
        module.__dict__["_rlg_folded"] = cache
        module.__dict__.pop("_rlg_packed", None)
    return cache[1]


def packed_trunk_cached(module: nn.Module) -> torch.Tensor:
    """bf16 images of the folded weights, cached next to (and invalidated with) the folded fp32 weights."""
    layers = folded_trunk_cached(module)
    packed = module.__dict__.get("_rlg_packed")
    if packed is None:
        packed = pack_bf16(layers)
        module.__dict__["_rlg_packed"] = packed
    return packed


# ---- the CUDA trunk ---------------------------------------------------------------------------------
def is_hot_path_input(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
            and x.shape[2] == 3 and x.shape[1] >= 1)


_PRECISION = "fp32"


def set_encoder_precision(precision: str) -> None:
    """Default arithmetic of the fused trunk: "fp32" (CUDA cores, GFVs within 1e-5 of the reference) or "bf16"
    (tcgen05/TMEM tensor-core GEMMs with fp32 accumulation, GFVs within 2e-2; widths must fit the tensor path).
    A module can override it with an attribute `rlg_precision`."""
    global _PRECISION
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION = precision


def get_encoder_precision() -> str:
    return _PRECISION


def _layer_array(layers):
    L = len(layers)
    arr = (_lib.RlgLayer * L)()
    keep = []
    for l, (w, b) in enumerate(layers):
        w = w.contiguous()
        b = b.contiguous()
        if not (w.is_cuda and b.is_cuda and w.dtype == torch.float32 and b.dtype == torch.float32 and w.dim() == 2):
            raise ValueError("folded layers must be CUDA float32 (c_out,c_in) / (c_out)")
        keep += [w, b]
        arr[l].w, arr[l].b = w.data_ptr(), b.data_ptr()
        arr[l].c_out, arr[l].c_in = w.shape[0], w.shape[1]
    return arr, keep


def pack_bf16(layers: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """bf16 weight images of layers >= 1 in the swizzled shared-memory layout the tcgen05 kernel loads with TMA
    bulk copies.  Raises RlgError (RLG_ERR_UNSUPPORTED) if the widths do not fit the tensor path."""
    lib = _lib.load()
    arr, keep = _layer_array(layers)
    dev = layers[0][0].device
    with torch.cuda.device(dev):
        nbytes = lib.rlg_encoder_pack_bytes(arr, len(layers))
        if nbytes == 0:
            _lib.check("rlg_encoder_pack_bytes", -4)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = lib.rlg_encoder_pack_bf16(arr, len(layers), packed.data_ptr(), packed.numel(),
                                       torch.cuda.current_stream(dev).cuda_stream)
        _lib.check("rlg_encoder_pack_bf16", rc)
    return packed


def encoder_pool(x: torch.Tensor, layers: List[Tuple[torch.Tensor, torch.Tensor]], want_argmax: bool = False,
                 precision: str = "fp32", packed: Optional[torch.Tensor] = None):
    """pooled (B, C_last) = max over points of the folded per-point MLP (ReLU after every layer), i.e.
    torch.max(point_mlp(x.transpose(2,1)), dim=2)[0] of models/autoencoder.py:65-71 in eval mode.
    Returns (pooled fp32, argmax int32 or None).  precision="bf16" runs layers >= 1 on the tensor cores."""
    if not is_hot_path_input(x):
        raise ValueError("gan-rl_3d_b200 encoder needs a CUDA float32 tensor (B,N,3) with N >= 1; there is no CPU path")
    lib = _lib.load()
    x = x.contiguous()
    B, N, _ = x.shape
    L = len(layers)
    arr, keep = _layer_array(layers)
    c_last = layers[-1][0].shape[0]
    if precision == "bf16":
        if want_argmax:
            raise ValueError("the bf16 tensor-core path does not report argmax indices")
        if packed is None:
            packed = pack_bf16(layers)
        pooled = torch.empty((B, c_last), dtype=torch.float32, device=x.device)
        if B == 0:
            return pooled, None
        with torch.cuda.device(x.device):
            rc = lib.rlg_encoder_fwd_bf16(x.data_ptr(), B, N, arr, L, packed.data_ptr(), packed.numel(),
                                          pooled.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
            _lib.check("rlg_encoder_fwd_bf16", rc)
        return pooled, None
    pooled = torch.empty((B, c_last), dtype=torch.float32, device=x.device)
    argmax = torch.empty((B, c_last), dtype=torch.int32, device=x.device) if want_argmax else None
    if B == 0:
        return pooled, argmax
    with torch.cuda.device(x.device):
        nbytes = lib.rlg_encoder_ws_bytes(B, N, arr, L)
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = lib.rlg_encoder_fwd(x.data_ptr(), B, N, arr, L, pooled.data_ptr(),
                                 argmax.data_ptr() if want_argmax else None,
                                 ws.data_ptr(), ws.numel(), stream)
        _lib.check("rlg_encoder_fwd", rc)
    return pooled, argmax


class EncoderTrunkFn(torch.autograd.Function):
    """x (B,N,3) -> pooled (B,C_last) through the fused kernel.  The backward (rare: eval mode with autograd
    on) recomputes the trunk with the module's own stock layers and differentiates that, so gradients reach
    the input and the original parameters exactly as in the reference graph."""

    @staticmethod
    def forward(ctx, x, module, *params):
        pooled = _trunk_pool(module, x)
        ctx.module = module
        ctx.n_params = len(params)
        ctx.save_for_backward(x)
        return pooled

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        module = ctx.module
        params = list(module.point_mlp.parameters())
        with torch.enable_grad():
            xin = x.detach().requires_grad_(True)
            feat = torch.max(module.point_mlp(xin.transpose(2, 1)), dim=2)[0]
            wanted = [xin] + params
            grads = torch.autograd.grad(feat, wanted, g, allow_unused=True)
        gx = grads[0] if ctx.needs_input_grad[0] else None
        gps = tuple(gp if ctx.needs_input_grad[2 + k] else None for k, gp in enumerate(grads[1:]))
        return (gx, None) + gps


def _trunk_pool(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    precision = getattr(module, "rlg_precision", _PRECISION)
    layers = folded_trunk_cached(module)
    if precision == "bf16":
        return encoder_pool(x, layers, precision="bf16", packed=packed_trunk_cached(module))[0]
    return encoder_pool(x, layers)[0]


def fused_forward(self: nn.Module, x: torch.Tensor, _original=None) -> torch.Tensor:
    """Drop-in for PointNetEncoder.forward (models/autoencoder.py:56-76)."""
    if self.training or not is_hot_path_input(x) or x.shape[0] == 0:
        if _original is not None:
            return _original(self, x)
        # stock path with the module's own layers (train-mode BatchNorm needs batch statistics)
        feat = torch.max(self.point_mlp(x.transpose(2, 1)), dim=2)[0]
        return self.global_mlp(feat)
    params = list(self.point_mlp.parameters())
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        pooled = EncoderTrunkFn.apply(x, self, *params)
    else:
        pooled = _trunk_pool(self, x)
    return self.global_mlp(pooled)


class PointNetEncoder(nn.Module):
    """Same constructor, module tree and state_dict keys as the reference's PointNetEncoder
    (models/autoencoder.py:13-54); forward runs the fused B200 trunk in eval mode."""

    def __init__(self, input_dim: int = 3, latent_dim: int = 128,
                 hidden_dims: Optional[List[int]] = None):
        super().__init__()
        hidden_dims = [64, 128, 128, 256, 128] if hidden_dims is None else list(hidden_dims)
        self.input_dim, self.latent_dim, self.hidden_dims = input_dim, latent_dim, hidden_dims
        seq: List[nn.Module] = []
        c_in = input_dim
        for c_out in hidden_dims:
            seq += [nn.Conv1d(c_in, c_out, 1), nn.BatchNorm1d(c_out), nn.ReLU(inplace=True)]
            c_in = c_out
        self.point_mlp = nn.Sequential(*seq)
        self.global_mlp = nn.Sequential(nn.Linear(c_in, latent_dim), nn.BatchNorm1d(latent_dim),
                                        nn.ReLU(inplace=True))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return fused_forward(self, x)
