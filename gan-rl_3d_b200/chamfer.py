"""Batched Chamfer distance on B200: torch.autograd.Function over the C ABI, plus mirrors of the reference's
loss API (utils/losses.py:13-75 of phanich004/GAN-RL_3D) with the same names, arguments and return shapes.

    chamfer_distance_l2(pc1, pc2) -> (dist1 (B,), dist2 (B,))       utils/losses.py:13-39
    chamfer_distance(pc1, pc2, bidirectional=True) -> (B,)           utils/losses.py:42-59
    ChamferLoss(bidirectional=True)(pred, target) -> 0-dim           utils/losses.py:62-75

Inputs must be CUDA fp32 (B,N,3)/(B,M,3) with N,M >= 1; there is no CPU implementation here.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import os

import torch
import torch.nn as nn

from . import _lib


def is_hot_path_input(pc1, pc2) -> bool:
    """The input contract of the CUDA path.  Anything else is not the hot path."""
    return (isinstance(pc1, torch.Tensor) and isinstance(pc2, torch.Tensor)
            and pc1.is_cuda and pc2.is_cuda and pc1.device == pc2.device
            and pc1.dtype == torch.float32 and pc2.dtype == torch.float32
            and pc1.dim() == 3 and pc2.dim() == 3 and pc1.shape[2] == 3 and pc2.shape[2] == 3
            and pc1.shape[0] == pc2.shape[0] and pc1.shape[1] >= 1 and pc2.shape[1] >= 1)


def _require_hot_path(pc1, pc2) -> None:
    if not is_hot_path_input(pc1, pc2):
        def d(t):
            return f"{tuple(t.shape)} {t.dtype} {t.device}" if isinstance(t, torch.Tensor) else repr(type(t))
        raise ValueError("gan-rl_3d_b200 Chamfer needs CUDA float32 tensors (B,N,3) and (B,M,3) on one device with "
                         f"N,M >= 1; got {d(pc1)} and {d(pc2)}. There is no CPU path.")


class _Workspace:
    """Caller-owned workspace of rlg_chamfer_fwd, cached per (device, stream, shape).  The forward leaves it
    in the all-ones state it needs on entry, so repeat calls skip the memset (RLG_CHAMFER_WS_CLEAN).

    A call made while its stream is being captured pins the entry for the life of the process: the CUDA graph holds the
    raw pointer, so the buffer must never go back to the allocator.  Unpinned entries are evicted one at a time, least
    recently used first.  `clean` tracks eager execution only: a captured call that finds the workspace not yet clean
    records the memset in the graph and leaves `clean` alone (nothing has run yet)."""
    _cache: "OrderedDict[tuple, _Workspace]" = OrderedDict()
    MAX_ENTRIES = 64

    def __init__(self, nbytes: int, device: torch.device):
        self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        self.clean = False
        self.pinned = False

    @classmethod
    def get(cls, device: torch.device, stream: int, B: int, N: int, M: int, capturing: bool = False) -> "_Workspace":
        key = (device.index, stream, B, N, M)
        ws = cls._cache.get(key)
        if ws is None:
            if not capturing:               # never free memory during a capture
                for old_key in [k for k, v in cls._cache.items() if not v.pinned]:
                    if len(cls._cache) < cls.MAX_ENTRIES:
                        break
                    del cls._cache[old_key]
            nbytes = _lib.load().rlg_chamfer_ws_bytes(B, N, M)
            ws = cls(nbytes, device)
            cls._cache[key] = ws
        else:
            cls._cache.move_to_end(key)
        if capturing:
            ws.pinned = True
        return ws


_DEFAULT_SWEEP = os.environ.get("RLG_CHAMFER_SWEEP", "auto").lower()
if _DEFAULT_SWEEP not in ("fp32", "tensor", "auto"):
    raise ValueError("RLG_CHAMFER_SWEEP must be fp32, tensor or auto")


_RESERVED_SMS = 0


def set_reserved_sms(n: int) -> None:
    """SMs the persistent Chamfer forward leaves free (RLG_CHAMFER_RESERVE_SMS): 1 when NCCL kernels run beside it (one
    process per GPU with per-step collectives on a side stream; distributed.init_from_env sets it), 0 otherwise."""
    global _RESERVED_SMS
    if not 0 <= int(n) <= 255:
        raise ValueError("reserved SMs must be in 0..255")
    _RESERVED_SMS = int(n)


def get_reserved_sms() -> int:
    return _RESERVED_SMS


_DETERMINISTIC: Optional[bool] = {"1": True, "0": False}.get(os.environ.get("RLG_DETERMINISTIC", ""))


def set_deterministic_backward(mode: Optional[bool]) -> None:
    """True: every Chamfer backward runs the run-to-run reproducible kernels (rlg_chamfer_bwd_det: fixed-point integer
    atomics instead of float atomics, about twice the time of a 7 us kernel); False: never; None (default): follow
    torch.are_deterministic_algorithms_enabled().  The forward is reproducible either way."""
    global _DETERMINISTIC
    _DETERMINISTIC = None if mode is None else bool(mode)


def deterministic_backward() -> bool:
    return torch.are_deterministic_algorithms_enabled() if _DETERMINISTIC is None else _DETERMINISTIC


def _bwd_workspace(lib, B: int, N: int, M: int, device: torch.device) -> torch.Tensor:
    # from torch's caching allocator on the current stream (inside a capture: the graph's private pool)
    return torch.empty(max(int(lib.rlg_chamfer_bwd_ws_bytes(B, N, M)), 16), dtype=torch.uint8, device=device)


def set_default_sweep(kind: str) -> None:
    """Which kernel sweeps the N x M pairs when chamfer_nearest() is not told explicitly:
    'fp32'   the FP32-pipe filter + refinement kernel (chamfer_filter.cu),
    'tensor' the contraction on tcgen05 with split-tf32 operands, refinement fused in (chamfer_tcsweep.cu),
    'auto'   (default) tensor when both clouds have at least 64 points, else fp32 (a 128 x 256 tensor tile is mostly
             padding for tiny clouds).
    Both give the same bits."""
    global _DEFAULT_SWEEP
    if kind not in ("fp32", "tensor", "auto"):
        raise ValueError("kind must be 'fp32', 'tensor' or 'auto'")
    _DEFAULT_SWEEP = kind


def get_default_sweep() -> str:
    return _DEFAULT_SWEEP


def _use_tensor(tensor: Optional[bool], n: int, m: int) -> bool:
    if tensor is not None:
        return bool(tensor)
    return _DEFAULT_SWEEP == "tensor" or (_DEFAULT_SWEEP == "auto" and n >= 64 and m >= 64)


def chamfer_nearest(pc1: torch.Tensor, pc2: torch.Tensor, want_means: bool = True, simple: bool = False,
                    loss_weights: Optional[Tuple[float, float]] = None, tensor: Optional[bool] = None,
                    track_two: bool = False,
                    zero_grads: Optional[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = None):
    """Nearest neighbours in both directions (no autograd).

    Returns (d1 (B,N) fp32, d2 (B,M) fp32, i1 (B,N) int32, i2 (B,M) int32, mean1 (B,), mean2 (B,)):
    exactly torch.min(torch.cdist(pc1,pc2), 2) / (…, 1) of utils/losses.py:29-33 with cdist in direct mode
    (ties as torch.min sees them on the sqrt-ed distances: lowest index), and torch.mean(…, dim=1) of :36-37.
    With loss_weights=(w1,w2) a 7th element is appended: the 0-dim batch loss sum_b(w1*mean1[b] + w2*mean2[b])
    reduced inside the same launches (utils/losses.py:54-59,75).
    simple selects the cross-check kernel (one thread per query, every candidate evaluated); tensor=True/False the
    pair sweep with the contraction on the tensor cores (tcgen05, split-tf32) or on the FP32 pipe (None: the module
    default, see set_default_sweep); track_two forces the tensor sweep's runner-up-group report (automatic beyond
    4096 points).  All paths return the same bits.
    zero_grads=(g1, g2): (B,N,3)/(B,M,3) buffers (each may be None) the forward zero-fills on the way, so
    chamfer_backward(..., out=(g1, g2), accumulate=True) is a single launch."""
    _require_hot_path(pc1, pc2)
    lib = _lib.load()
    pc1 = pc1.contiguous()
    pc2 = pc2.contiguous()
    B, N, _ = pc1.shape
    M = pc2.shape[1]
    dev = pc1.device
    d1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, M), dtype=torch.float32, device=dev)
    i1 = torch.empty((B, N), dtype=torch.int32, device=dev)
    i2 = torch.empty((B, M), dtype=torch.int32, device=dev)
    m1 = torch.empty((B,), dtype=torch.float32, device=dev) if want_means else None
    m2 = torch.empty((B,), dtype=torch.float32, device=dev) if want_means else None
    loss = torch.empty((), dtype=torch.float32, device=dev) if loss_weights is not None else None   # fully written
    if B == 0:
        return (d1, d2, i1, i2, m1, m2) if loss is None else (d1, d2, i1, i2, m1, m2, loss.zero_())
    w1, w2 = loss_weights if loss_weights is not None else (0.0, 0.0)
    gz1, gz2 = zero_grads if zero_grads is not None else (None, None)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        capturing = torch.cuda.is_current_stream_capturing()
        ws = _Workspace.get(dev, stream, B, N, M, capturing)
        flags = 0
        if simple:
            flags |= _lib.CHAMFER_ALGO_SIMPLE
        else:
            if ws.clean:
                flags |= _lib.CHAMFER_WS_CLEAN
            if _use_tensor(tensor, N, M):
                flags |= _lib.CHAMFER_ALGO_TENSOR | _lib.chamfer_reserve_sms(_RESERVED_SMS)
                if track_two:
                    flags |= _lib.CHAMFER_TRACK_TWO
        was_clean = ws.clean
        if not capturing and not simple:
            ws.clean = False                     # until the call has gone through
        _lib.nvtx_push("rlg.chamfer_fwd")
        rc = lib.rlg_chamfer_loss_fwd(pc1.data_ptr(), pc2.data_ptr(), B, N, M,
                                      d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                      m1.data_ptr() if want_means else None, m2.data_ptr() if want_means else None,
                                      loss.data_ptr() if loss is not None else None, w1, w2,
                                      gz1.data_ptr() if gz1 is not None else None,
                                      gz2.data_ptr() if gz2 is not None else None,
                                      ws.buf.data_ptr(), ws.buf.numel(), flags, stream)
        _lib.nvtx_pop()
        _lib.check("rlg_chamfer_loss_fwd", rc)
        if not simple:
            # eager: the forward restored the all-ones state.  Captured: nothing ran; the recorded sequence keeps the
            # invariant on every replay, and the host-side flag stays what eager execution last established.
            ws.clean = was_clean if capturing else True
    return (d1, d2, i1, i2, m1, m2) if loss is None else (d1, d2, i1, i2, m1, m2, loss)


def chamfer_backward(pc1, pc2, d1, d2, i1, i2, g1, g2, out=None, accumulate: bool = False,
                     deterministic: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gradient of (mean1, mean2) w.r.t. (pc1, pc2) for upstream (g1 (B,), g2 (B,)); None = zero.
    out=(gpc1, gpc2) with accumulate=True adds into buffers that already hold zeros (or a gradient).
    deterministic: None = the process-wide setting (set_deterministic_backward)."""
    lib = _lib.load()
    B, N, _ = pc1.shape
    M = pc2.shape[1]
    gpc1, gpc2 = out if out is not None else (torch.empty_like(pc1), torch.empty_like(pc2))
    if B == 0:
        return gpc1, gpc2
    g1 = g1.contiguous().float() if g1 is not None else None
    g2 = g2.contiguous().float() if g2 is not None else None
    with torch.cuda.device(pc1.device):
        stream = torch.cuda.current_stream(pc1.device).cuda_stream
        flags = _lib.CHAMFER_BWD_ACCUMULATE if (accumulate and out is not None) else 0
        args = (pc1.data_ptr(), pc2.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                g1.data_ptr() if g1 is not None else None, g2.data_ptr() if g2 is not None else None,
                B, N, M, gpc1.data_ptr(), gpc2.data_ptr())
        _lib.nvtx_push("rlg.chamfer_bwd")
        if deterministic_backward() if deterministic is None else deterministic:
            ws = _bwd_workspace(lib, B, N, M, pc1.device)
            rc = lib.rlg_chamfer_bwd_det(*args, ws.data_ptr(), ws.numel(), flags, stream)
        else:
            rc = lib.rlg_chamfer_bwd(*args, flags, stream)
        _lib.nvtx_pop()
        _lib.check("rlg_chamfer_bwd", rc)
    return gpc1, gpc2


class ChamferFn(torch.autograd.Function):
    """(pc1 (B,N,3), pc2 (B,M,3)) -> (mean1 (B,), mean2 (B,)), the two outputs of the reference's
    chamfer_distance_l2 (utils/losses.py:13-39).  Saves the per-point distances and int32 argmin indices
    instead of the reference's (B,N,M) matrix."""

    @staticmethod
    def forward(ctx, pc1, pc2):
        pc1c, pc2c = pc1.contiguous(), pc2.contiguous()
        need = any(ctx.needs_input_grad)
        gbuf = (torch.empty_like(pc1c), torch.empty_like(pc2c)) if need else None   # zero-filled by the forward
        d1, d2, i1, i2, m1, m2 = chamfer_nearest(pc1c, pc2c, want_means=True, zero_grads=gbuf)
        if need:
            ctx.save_for_backward(pc1c, pc2c, d1, d2, i1, i2)
            ctx.gbuf = gbuf
        return m1, m2

    @staticmethod
    def backward(ctx, g1, g2):
        pc1, pc2, d1, d2, i1, i2 = ctx.saved_tensors
        gbuf, ctx.gbuf = ctx.gbuf, None          # a second backward through the same graph allocates afresh
        gpc1, gpc2 = chamfer_backward(pc1, pc2, d1, d2, i1, i2, g1, g2, out=gbuf, accumulate=True)
        return (gpc1 if ctx.needs_input_grad[0] else None, gpc2 if ctx.needs_input_grad[1] else None)


def _loss_weights(B: int, bidirectional: bool) -> Tuple[float, float]:
    # torch.mean over the batch of (dist1+dist2)/2 (bidirectional) or dist1   (utils/losses.py:56-59, :75)
    return (0.5 / B, 0.5 / B) if bidirectional else (1.0 / B, 0.0)


class ChamferLossFn(torch.autograd.Function):
    """(pred (B,N,3), target (B,M,3)) -> the 0-dim ChamferLoss of utils/losses.py:62-75 with the batch reduction
    fused into the forward launch and its scaling fused into the backward: 2 + 1 kernel launches per
    training step (the forward also zero-fills the gradient buffers), no elementwise torch kernels in between."""

    @staticmethod
    def forward(ctx, pc1, pc2, bidirectional: bool):
        pc1c, pc2c = pc1.contiguous(), pc2.contiguous()
        w = _loss_weights(max(pc1c.shape[0], 1), bidirectional)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gbuf = (torch.empty_like(pc1c), torch.empty_like(pc2c)) if need else None   # zero-filled by the forward
        d1, d2, i1, i2, m1, m2, loss = chamfer_nearest(pc1c, pc2c, want_means=True, loss_weights=w, zero_grads=gbuf)
        if need:
            ctx.save_for_backward(pc1c, pc2c, d1, d2, i1, i2)
            ctx.w = w
            ctx.gbuf = gbuf
        return loss

    @staticmethod
    def backward(ctx, gloss):
        pc1, pc2, d1, d2, i1, i2 = ctx.saved_tensors
        lib = _lib.load()
        B, N, _ = pc1.shape
        M = pc2.shape[1]
        gbuf, ctx.gbuf = ctx.gbuf, None          # a second backward through the same graph allocates afresh
        flags = _lib.CHAMFER_BWD_ACCUMULATE if gbuf is not None else 0
        gpc1, gpc2 = gbuf if gbuf is not None else (torch.empty_like(pc1), torch.empty_like(pc2))
        gloss = gloss.contiguous().float()
        with torch.cuda.device(pc1.device):
            stream = torch.cuda.current_stream(pc1.device).cuda_stream
            args = (pc1.data_ptr(), pc2.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                    gloss.data_ptr(), ctx.w[0], ctx.w[1], B, N, M, gpc1.data_ptr(), gpc2.data_ptr())
            _lib.nvtx_push("rlg.chamfer_loss_bwd")
            if deterministic_backward():
                ws = _bwd_workspace(lib, B, N, M, pc1.device)
                rc = lib.rlg_chamfer_loss_bwd_det(*args, ws.data_ptr(), ws.numel(), flags, stream)
            else:
                rc = lib.rlg_chamfer_loss_bwd(*args, flags, stream)
            _lib.nvtx_pop()
            _lib.check("rlg_chamfer_loss_bwd", rc)
        return (gpc1 if ctx.needs_input_grad[0] else None, gpc2 if ctx.needs_input_grad[1] else None, None)


# ---- mirrors of the reference API -------------------------------------------------------------------
def chamfer_distance_l2(pc1: torch.Tensor, pc2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for utils/losses.py:13-39.  Returns (dist1 (B,), dist2 (B,)): the mean over points of the
    (non-squared) L2 distance to the nearest point of the other cloud, in each direction."""
    _require_hot_path(pc1, pc2)
    return ChamferFn.apply(pc1, pc2)


def chamfer_distance(pc1: torch.Tensor, pc2: torch.Tensor, bidirectional: bool = True) -> torch.Tensor:
    """Drop-in for utils/losses.py:42-59."""
    dist1, dist2 = chamfer_distance_l2(pc1, pc2)
    if bidirectional:
        return (dist1 + dist2) / 2.0
    return dist1


class ChamferLoss(nn.Module):
    """Drop-in for utils/losses.py:62-75."""

    def __init__(self, bidirectional: bool = True):
        super().__init__()
        self.bidirectional = bidirectional

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        _require_hot_path(pred, target)
        if pred.shape[0] == 0:
            return torch.mean(chamfer_distance(pred, target, self.bidirectional))   # nan, as the reference
        return ChamferLossFn.apply(pred, target, self.bidirectional)
