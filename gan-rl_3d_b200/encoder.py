"""PointNet encoder trunk on B200: fused per-point MLP + global max-pool behind the C ABI.

Mirrors models/autoencoder.py:13-76 of phanich004/GAN-RL_3D.  `fused_forward` has the signature of
`PointNetEncoder.forward(self, x (B,N,3)) -> (B, latent_dim)` and works on any module with the reference's
layout (`point_mlp` = [Conv1d(k=1), BatchNorm1d, ReLU] x L, `global_mlp` = [Linear, BatchNorm1d, ReLU]), so
rebinding the class attribute keeps parameters, buffers and state_dict keys untouched.

Eval mode: BatchNorm with running statistics is an affine map and is folded into the conv weights (cached;
invalidated when any parameter/buffer changes); the trunk then runs on the tensor cores at fp32 accuracy by default
(layer-wise tcgen05 GEMMs on fp16 hi+lo operand pairs), see set_encoder_precision.  In train mode BatchNorm needs
batch statistics over B*N points and updates its buffers: see train.py.
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
    """Changes whenever a parameter or buffer of the trunk is modified in place, moved or replaced."""
    return tuple((id(t), t.data_ptr(), t._version) for m in point_mlp for t in (*m._parameters.values(), *m._buffers.values())
                 if t is not None)


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


def _packed_cache(module: nn.Module) -> dict:
    folded_trunk_cached(module)                      # invalidates the packed images together with the folded weights
    return module.__dict__.setdefault("_rlg_packed", {})


def packed_trunk_cached(module: nn.Module) -> torch.Tensor:
    """bf16 image of the folded weights for the single fused tcgen05 kernel, cached next to the folded fp32 weights."""
    cache = _packed_cache(module)
    if "fused" not in cache:
        cache["fused"] = pack_bf16(folded_trunk_cached(module))
    return cache["fused"]


# ---- the CUDA trunk ---------------------------------------------------------------------------------
def is_hot_path_input(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
            and x.shape[2] == 3 and x.shape[1] >= 1)


PRECISIONS = ("auto", "fp32", "fp32x", "bf16")
_PRECISION = "auto"


def set_encoder_precision(precision: str) -> None:
    """Arithmetic of the fused trunk (a module can override it with an attribute `rlg_precision`):
      "fp32x"  tensor cores at fp32 accuracy: layer-wise tcgen05 GEMMs on fp16 hi+lo operand pairs (22 bits), GFVs within
               1e-5 of the reference.  Widths must be multiples of 64 with hidden widths <= 256.
      "fp32"   the CUDA-core kernel (any widths, reports argmax indices), GFVs within 1e-5.
      "bf16"   bf16 tensor-core GEMMs with fp32 accumulation, GFVs within 2e-2: the single fused tcgen05 kernel when the
               widths fit it (hidden <= 128, last a multiple of 128), else the layer-wise bf16 GEMMs.
      "auto"   (default) "fp32x" where the widths allow, else "fp32".
    A precision whose kernels do not cover the module's widths falls back to "fp32" (never raises)."""
    global _PRECISION
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}")
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
    """bf16 weight images of layers >= 1 in the swizzled shared-memory layout the fused tcgen05 kernel loads with TMA
    bulk copies.  Raises RlgError (RLG_ERR_UNSUPPORTED) if the widths do not fit that kernel."""
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


def fused_tc_supported(layers) -> bool:
    """Widths the single fused tcgen05 kernel (encoder_tc.cu) covers."""
    dims = [w.shape[0] for w, _ in layers]
    return (len(dims) >= 2 and all(d in (64, 128) for d in dims[:-1]) and dims[-1] % 128 == 0)


def gemm_supported(layers) -> bool:
    """Widths the layer-wise tcgen05 GEMM path (encoder_layers.cu) covers."""
    dims = [w.shape[0] for w, _ in layers]
    return (2 <= len(dims) <= 8 and all(d % 64 == 0 and d >= 64 for d in dims) and all(d <= 256 for d in dims[:-1])
            and layers[0][0].shape[1] == 3)


def pack_gemm(layers, mode: int):
    """Operand image of the folded weights for the layer-wise GEMM path + the per-layer power-of-two weight scales
    (largest power of two with max|w| * scale <= 2^14; a host synchronisation, paid when the weights change)."""
    lib = _lib.load()
    arr, keep = _layer_array(layers)
    dev = layers[0][0].device
    L = len(layers)
    scales = (ctypes.c_float * L)()
    scales[0] = 1.0
    for l in range(1, L):
        m = float(layers[l][0].abs().max().item()) if mode == _lib.ENC_FP32X else 1.0
        e = 0 if mode != _lib.ENC_FP32X or not (m > 0.0) or m != m else max(-20, min(20, int(torch.floor(torch.log2(torch.tensor(16384.0 / m))).item())))
        scales[l] = float(2.0 ** e)
    with torch.cuda.device(dev):
        nbytes = lib.rlg_encoder_gemm_pack_bytes(arr, L, mode)
        if nbytes == 0:
            _lib.check("rlg_encoder_gemm_pack_bytes", -4)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = lib.rlg_encoder_gemm_pack(arr, L, mode, scales, packed.data_ptr(), packed.numel(),
                                       torch.cuda.current_stream(dev).cuda_stream)
        _lib.check("rlg_encoder_gemm_pack", rc)
    return packed, scales


_gemm_ws = {}


def _gemm_workspace(dev, stream, nbytes):
    key = (dev.index, stream)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing() and ws is not None:
            raise RuntimeError("the encoder workspace cannot grow while its stream is being captured: run the shape once eagerly first")
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        _gemm_ws[key] = ws
    return ws


def encoder_pool_gemm(x: torch.Tensor, layers, mode: int, packed=None) -> torch.Tensor:
    """pooled (B, C_last) through the layer-wise tcgen05 GEMM path (mode = _lib.ENC_BF16 or _lib.ENC_FP32X)."""
    if not is_hot_path_input(x):
        raise ValueError("gan-rl_3d_b200 encoder needs a CUDA float32 tensor (B,N,3) with N >= 1; there is no CPU path")
    lib = _lib.load()
    x = x.contiguous()
    B, N, _ = x.shape
    L = len(layers)
    arr, keep = _layer_array(layers)
    if packed is None:
        packed = pack_gemm(layers, mode)
    image, scales = packed
    pooled = torch.empty((B, layers[-1][0].shape[0]), dtype=torch.float32, device=x.device)
    if B == 0:
        return pooled
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        ws = _gemm_workspace(x.device, stream, lib.rlg_encoder_gemm_ws_bytes(B, N, arr, L, mode))
        _lib.nvtx_push("rlg.encoder_gemm_fwd")
        rc = lib.rlg_encoder_gemm_fwd(x.data_ptr(), B, N, arr, L, mode, scales, image.data_ptr(), image.numel(),
                                      pooled.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.nvtx_pop()
        _lib.check("rlg_encoder_gemm_fwd", rc)
    return pooled


def encoder_pool(x: torch.Tensor, layers: List[Tuple[torch.Tensor, torch.Tensor]], want_argmax: bool = False,
                 precision: str = "fp32", packed: Optional[torch.Tensor] = None):
    """pooled (B, C_last) = max over points of the folded per-point MLP (ReLU after every layer), i.e.
    torch.max(point_mlp(x.transpose(2,1)), dim=2)[0] of models/autoencoder.py:65-71 in eval mode.
    Returns (pooled fp32, argmax int32 or None).  precision: "fp32" (CUDA cores, optional argmax), "bf16" (the single
    fused tcgen05 kernel), "fp32x" / "bf16_layers" (layer-wise tcgen05 GEMMs, see set_encoder_precision)."""
    if not is_hot_path_input(x):
        raise ValueError("gan-rl_3d_b200 encoder needs a CUDA float32 tensor (B,N,3) with N >= 1; there is no CPU path")
    if precision in ("fp32x", "bf16_layers"):
        if want_argmax:
            raise ValueError("the tensor-core paths do not report argmax indices")
        return encoder_pool_gemm(x, layers, _lib.ENC_FP32X if precision == "fp32x" else _lib.ENC_BF16, packed), None
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
            _lib.nvtx_push("rlg.encoder_fwd_bf16")
            rc = lib.rlg_encoder_fwd_bf16(x.data_ptr(), B, N, arr, L, packed.data_ptr(), packed.numel(),
                                          pooled.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
            _lib.nvtx_pop()
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
        _lib.nvtx_push("rlg.encoder_fwd")
        rc = lib.rlg_encoder_fwd(x.data_ptr(), B, N, arr, L, pooled.data_ptr(),
                                 argmax.data_ptr() if want_argmax else None,
                                 ws.data_ptr(), ws.numel(), stream)
        _lib.nvtx_pop()
        _lib.check("rlg_encoder_fwd", rc)
    return pooled, argmax


def resolve_path(layers, precision: str) -> str:
    """Which kernel family runs a layer list under a requested precision: "fp32", "fp32x", "bf16" (fused) or
    "bf16_layers".  Unsupported widths fall back to "fp32"."""
    if precision in ("auto", "fp32x"):
        return "fp32x" if gemm_supported(layers) else "fp32"
    if precision == "bf16":
        if fused_tc_supported(layers):
            return "bf16"
        return "bf16_layers" if gemm_supported(layers) else "fp32"
    return "fp32"


_PATH_NAMES = {"fp32": "fp32 CUDA-core fused trunk (encoder_fp32_kernel)",
               "fp32x": "layer-wise tcgen05 GEMMs, fp16 hi+lo operands = fp32-grade (encoder_layer_kernel<2>)",
               "bf16": "single fused tcgen05/TMEM bf16 kernel (encoder_tc_kernel)",
               "bf16_layers": "layer-wise tcgen05 GEMMs, bf16 operands (encoder_layer_kernel<1>)"}


def encoder_path_of(module: nn.Module) -> str:
    """Human-readable name of the kernel family `module` runs in eval mode under its current precision setting."""
    return _PATH_NAMES[resolve_path(folded_trunk_cached(module), getattr(module, "rlg_precision", _PRECISION))]


class _bn_eval:
    """Run a trunk with its BatchNorm layers in eval semantics regardless of module.training (restored on exit)."""

    def __init__(self, seq: nn.Sequential):
        self.bns = [m for m in seq if isinstance(m, nn.BatchNorm1d) and m.training]

    def __enter__(self):
        for m in self.bns:
            m.training = False

    def __exit__(self, *exc):
        for m in self.bns:
            m.training = True


class EncoderTrunkFn(torch.autograd.Function):
    """x (B,N,3) -> pooled (B,C_last) through the fused kernel.  The backward (rare: eval mode with autograd
    on) recomputes the trunk with the module's own stock layers -- BatchNorm forced to the eval semantics of the forward,
    whatever mode the module is in by then -- and differentiates that, so gradients reach the input and the original
    parameters exactly as in the reference graph."""

    @staticmethod
    def forward(ctx, x, module, *params):
        pooled = _trunk_pool(module, x)
        ctx.point_mlp = module.point_mlp
        ctx.n_params = len(params)
        ctx.save_for_backward(x)
        return pooled

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        point_mlp = ctx.point_mlp
        params = list(point_mlp.parameters())
        with torch.enable_grad(), _bn_eval(point_mlp):
            xin = x.detach().requires_grad_(True)
            feat = torch.max(point_mlp(xin.transpose(2, 1)), dim=2)[0]
            wanted = [xin] + params
            grads = torch.autograd.grad(feat, wanted, g, allow_unused=True)
        gx = grads[0] if ctx.needs_input_grad[0] else None
        gps = tuple(gp if ctx.needs_input_grad[2 + k] else None for k, gp in enumerate(grads[1:]))
        return (gx, None) + gps


class _Plan:
    """Everything derived from a trunk's weights that a forward needs, built once per state version: folded layers, the
    ctypes layer array, the resolved kernel family and its packed operand image."""
    __slots__ = ("layers", "arr", "keep", "path", "packed", "c_last", "device")

    def __init__(self, module: nn.Module, precision: str):
        self.layers = folded_trunk_cached(module)
        self.arr, self.keep = _layer_array(self.layers)
        self.path = resolve_path(self.layers, precision)
        self.c_last = self.layers[-1][0].shape[0]
        self.device = self.layers[0][0].device
        cache = _packed_cache(module)
        if self.path == "fp32":
            self.packed = None
        elif self.path == "bf16":
            self.packed = packed_trunk_cached(module)
        else:
            if self.path not in cache:
                cache[self.path] = pack_gemm(self.layers, _lib.ENC_FP32X if self.path == "fp32x" else _lib.ENC_BF16)
            self.packed = cache[self.path]


def _plan_of(module: nn.Module) -> Optional["_Plan"]:
    """The cached plan of `module`, or None when its trunk does not have the reference's layout."""
    precision = getattr(module, "rlg_precision", _PRECISION)
    try:
        ver = (_trunk_state_version(module.point_mlp), precision)
    except AttributeError:
        return None
    hit = module.__dict__.get("_rlg_plan")
    if hit is not None and hit[0] == ver:
        return hit[1]
    try:
        pairs = _trunk_layers(module.point_mlp)
    except (ValueError, AttributeError):
        return None
    w0 = pairs[0][0].weight
    if not w0.is_cuda or w0.dtype != torch.float32:          # CPU / fp64 / half modules are not the hot path
        return None
    plan = _Plan(module, precision)
    module.__dict__["_rlg_plan"] = (ver, plan)
    return plan


def _run_plan(plan: "_Plan", x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    x = x.contiguous()
    B, N, _ = x.shape
    L = len(plan.layers)
    pooled = torch.empty((B, plan.c_last), dtype=torch.float32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        if plan.path == "fp32":
            ws = torch.empty(max(lib.rlg_encoder_ws_bytes(B, N, plan.arr, L), 256), dtype=torch.uint8, device=x.device)
            _lib.nvtx_push("rlg.encoder_fwd")
            rc = lib.rlg_encoder_fwd(x.data_ptr(), B, N, plan.arr, L, pooled.data_ptr(), None, ws.data_ptr(), ws.numel(), stream)
            _lib.nvtx_pop()
            _lib.check("rlg_encoder_fwd", rc)
        elif plan.path == "bf16":
            _lib.nvtx_push("rlg.encoder_fwd_bf16")
            rc = lib.rlg_encoder_fwd_bf16(x.data_ptr(), B, N, plan.arr, L, plan.packed.data_ptr(), plan.packed.numel(),
                                          pooled.data_ptr(), stream)
            _lib.nvtx_pop()
            _lib.check("rlg_encoder_fwd_bf16", rc)
        else:
            mode = _lib.ENC_FP32X if plan.path == "fp32x" else _lib.ENC_BF16
            image, scales = plan.packed
            ws = _gemm_workspace(x.device, stream, lib.rlg_encoder_gemm_ws_bytes(B, N, plan.arr, L, mode))
            _lib.nvtx_push("rlg.encoder_gemm_fwd")
            rc = lib.rlg_encoder_gemm_fwd(x.data_ptr(), B, N, plan.arr, L, mode, scales, image.data_ptr(), image.numel(),
                                          pooled.data_ptr(), ws.data_ptr(), ws.numel(), stream)
            _lib.nvtx_pop()
            _lib.check("rlg_encoder_gemm_fwd", rc)
    return pooled


def _trunk_pool(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    plan = _plan_of(module)
    if plan is None:
        raise ValueError("point_mlp is not [Conv1d(k=1), BatchNorm1d, ReLU] x L")
    return _run_plan(plan, x)


def _module_is_hot(self: nn.Module, x) -> bool:
    """The module has the reference's trunk layout and its (float32) weights live on the device of x."""
    plan = _plan_of(self)
    return plan is not None and plan.device == x.device


_TRAIN_PATH = True


def set_train_path(enabled: bool) -> None:
    """Whether train-mode forwards (BatchNorm batch statistics) and eval-mode forwards under autograd run the B200
    train kernels (train.py) where the widths allow; off = the module's own stock layers."""
    global _TRAIN_PATH
    _TRAIN_PATH = bool(enabled)


def _train_plan_ok(self: nn.Module, x: torch.Tensor) -> bool:
    """The B200 train kernels cover this module and input (cached per parameter-storage layout)."""
    if not _TRAIN_PATH or x.requires_grad or x.shape[0] * x.shape[1] < 2:
        return False
    from .train import train_supported
    try:
        key = tuple((id(t), t.data_ptr()) for m in self.point_mlp for t in (*m._parameters.values(), *m._buffers.values())
                    if t is not None)
    except AttributeError:
        return False
    hit = self.__dict__.get("_rlg_train_ok")
    if hit is None or hit[0] != key:
        hit = (key, train_supported(self.point_mlp))
        self.__dict__["_rlg_train_ok"] = hit
    if not hit[1]:
        return False
    return self.point_mlp[0].weight.device == x.device


def fused_forward(self: nn.Module, x: torch.Tensor, _original=None) -> torch.Tensor:
    """Drop-in for PointNetEncoder.forward (models/autoencoder.py:56-76)."""
    hot = is_hot_path_input(x) and x.shape[0] > 0
    if hot and self.training and _train_plan_ok(self, x):
        from .train import trunk_pool_autograd
        return self.global_mlp(trunk_pool_autograd(self, x, batch_stats=True))
    if self.training or not hot or not _module_is_hot(self, x):
        if _original is not None:
            return _original(self, x)
        # stock path with the module's own layers
        feat = torch.max(self.point_mlp(x.transpose(2, 1)), dim=2)[0]
        return self.global_mlp(feat)
    params = list(self.point_mlp.parameters())
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        if _train_plan_ok(self, x):
            from .train import trunk_pool_autograd
            pooled = trunk_pool_autograd(self, x, batch_stats=False)
        else:
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
