"""Train-mode PointNet encoder trunk on B200: forward and backward of [Conv1d(k=1), BatchNorm1d, ReLU] x L + max over the
points, with BatchNorm's batch statistics and running-stat updates, behind the C ABI (csrc/encoder_train.cu).

This is what every autoencoder training step runs (train_rl_gan_net.py:220-249: `self.model.train()`, forward,
`loss.backward()`) through models/autoencoder.py:32-47,65-71.  The same kernels with running statistics give the
eval-mode backward (a frozen encoder inside an autograd graph).

`trunk_pool_autograd(module, x)` returns pooled (B, C_last) attached to the autograd graph of the module's own parameters
(conv weight (c_out,c_in,1) / bias, BatchNorm weight / bias), updates running_mean / running_var / num_batches_tracked in
train mode exactly as torch.nn.BatchNorm1d does, and never touches the module tree (state_dict keys stay byte-compatible).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib

MAX_WIDTH = 256


def train_supported(point_mlp: nn.Sequential) -> bool:
    """Widths / options the train-mode kernels cover: c_in 3, every width a multiple of 64 in [64, 256], 2..8 blocks,
    BatchNorm with running statistics and a fixed momentum, float32 CUDA parameters."""
    from .encoder import _trunk_layers
    try:
        pairs = _trunk_layers(point_mlp)
    except (ValueError, AttributeError):
        return False
    if not 2 <= len(pairs) <= 8 or pairs[0][0].in_channels != 3:
        return False
    for conv, bn in pairs:
        c = conv.out_channels
        if c % 64 != 0 or c < 64 or c > MAX_WIDTH:
            return False
        if not bn.track_running_stats or bn.running_mean is None or bn.momentum is None:
            return False
        tensors = [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        if any(t is not None and (not t.is_cuda or t.dtype != torch.float32) for t in tensors):
            return False
    return True


def _param_list(pairs) -> List[Optional[torch.Tensor]]:
    out: List[Optional[torch.Tensor]] = []
    for conv, bn in pairs:
        out += [conv.weight, conv.bias, bn.weight if bn.affine else None, bn.bias if bn.affine else None]
    return out


_ws_cache = {}


def _workspace(dev: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    key = (dev.index, stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the encoder workspace cannot grow while its stream is being captured: run the shape once eagerly first")
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def _layer_structs(pairs, tensors):
    """ctypes rlg_bn_layer[] over the given (weight, bias, gamma, beta) tensors and the modules' running statistics."""
    L = len(pairs)
    arr = (_lib.RlgBnLayer * L)()
    keep = []
    for l, (conv, bn) in enumerate(pairs):
        w, b, ga, be = tensors[4 * l: 4 * l + 4]
        w2 = w.detach().reshape(w.shape[0], w.shape[1]).contiguous()
        keep.append(w2)
        arr[l].w = w2.data_ptr()
        for name, t in (("b", b), ("gamma", ga), ("beta", be)):
            if t is not None:
                tc = t.detach().contiguous()
                keep.append(tc)
                setattr(arr[l], name, tc.data_ptr())
        arr[l].running_mean = bn.running_mean.data_ptr()
        arr[l].running_var = bn.running_var.data_ptr()
        arr[l].eps = float(bn.eps)
        arr[l].momentum = float(bn.momentum)
        arr[l].c_in, arr[l].c_out = w2.shape[1], w2.shape[0]
    return arr, keep


class EncoderTrainFn(torch.autograd.Function):
    """x (B,N,3) [no gradient], then (conv.weight, conv.bias, bn.weight, bn.bias) per block -> pooled (B, C_last)."""

    @staticmethod
    def forward(ctx, x, pairs, batch_stats, *params):
        lib = _lib.load()
        x = x.contiguous()
        B, N, _ = x.shape
        arr, keep = _layer_structs(pairs, params)
        L = len(pairs)
        dev = x.device
        from .chamfer import get_reserved_sms
        flags = (_lib.ENC_BATCH_STATS if batch_stats else 0) | ((get_reserved_sms() & 0xff) << 16)
        pooled = torch.empty((B, pairs[-1][0].out_channels), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            nsaved = lib.rlg_encoder_train_saved_bytes(B, N, arr, L)
            nws = lib.rlg_encoder_train_ws_bytes(B, N, arr, L)
            if nsaved == 0 or nws == 0:
                _lib.check("rlg_encoder_train_saved_bytes", -4)
            saved = torch.empty(nsaved, dtype=torch.uint8, device=dev)
            ws = _workspace(dev, stream, nws)
            _lib.nvtx_push("rlg.encoder_train_fwd")
            rc = lib.rlg_encoder_train_fwd(x.data_ptr(), B, N, arr, L, flags, pooled.data_ptr(), saved.data_ptr(), saved.numel(),
                                           ws.data_ptr(), ws.numel(), stream)
            _lib.nvtx_pop()
            _lib.check("rlg_encoder_train_fwd", rc)
        if batch_stats:
            with torch.no_grad():
                counters = [bn.num_batches_tracked for _, bn in pairs if bn.num_batches_tracked is not None]
                if counters:
                    torch._foreach_add_(counters, 1)            # one launch for all blocks
        ctx.pairs, ctx.flags, ctx.shape = pairs, flags, (B, N)
        ctx.save_for_backward(x, saved, *[p for p in params if p is not None])
        ctx.present = [p is not None for p in params]
        return pooled

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        x, saved, *present = ctx.saved_tensors
        it = iter(present)
        params = [next(it) if has else None for has in ctx.present]
        pairs = ctx.pairs
        B, N = ctx.shape
        L = len(pairs)
        arr, keep = _layer_structs(pairs, params)
        dev = x.device
        g = g.contiguous().float()
        grads = (_lib.RlgBnGrads * L)()
        outs: List[Optional[torch.Tensor]] = []
        for l in range(L):
            for k, name in enumerate(("dw", "db", "dgamma", "dbeta")):
                p = params[4 * l + k]
                if p is not None and ctx.needs_input_grad[3 + 4 * l + k]:
                    t = torch.empty(p.shape, dtype=torch.float32, device=dev)
                    setattr(grads[l], name, t.data_ptr())
                    outs.append(t)
                else:
                    outs.append(None)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = _workspace(dev, stream, lib.rlg_encoder_train_ws_bytes(B, N, arr, L))
            _lib.nvtx_push("rlg.encoder_train_bwd")
            rc = lib.rlg_encoder_train_bwd(x.data_ptr(), B, N, arr, L, ctx.flags, g.data_ptr(), saved.data_ptr(), saved.numel(),
                                           grads, ws.data_ptr(), ws.numel(), stream)
            _lib.nvtx_pop()
            _lib.check("rlg_encoder_train_bwd", rc)
        return (None, None, None) + tuple(outs)


def trunk_pool_autograd(module: nn.Module, x: torch.Tensor, batch_stats: Optional[bool] = None) -> torch.Tensor:
    """pooled (B, C_last) = torch.max(module.point_mlp(x.transpose(2, 1)), dim=2)[0] with autograd to the trunk's parameters.
    batch_stats defaults to module.training (train mode: batch statistics + running-stat update)."""
    from .encoder import _trunk_layers
    pairs = _trunk_layers(module.point_mlp)
    if batch_stats is None:
        batch_stats = bool(module.training)
    return EncoderTrainFn.apply(x, pairs, batch_stats, *_param_list(pairs))


def inspect_saved(module: nn.Module, x_shape, saved: torch.Tensor):
    """Inspection helper (tests, debugging): the piecewise-linear branch a forward took, read back from its `saved`
    buffer (layout: rlg_encoder_train_saved_layout).  Returns (masks, argmax, positive): masks[l] (B,N,C_l) bool = the
    ReLU of block l let the value through (l < L-1); argmax (B,C_last) int64 = the point the max-pool picked;
    positive (B,C_last) bool = that maximum is > 0.  `saved` of a forward is `pooled.grad_fn.saved_tensors[1]`."""
    import ctypes
    from .encoder import _trunk_layers
    lib = _lib.load()
    pairs = _trunk_layers(module.point_mlp)
    B, N = int(x_shape[0]), int(x_shape[1])
    L = len(pairs)
    arr, keep = _layer_structs(pairs, _param_list(pairs))
    offs = (ctypes.c_size_t * (5 * L + 1))()
    _lib.check("rlg_encoder_train_saved_layout", lib.rlg_encoder_train_saved_layout(B, N, arr, L, offs, 5 * L + 1))
    masks = []
    for l in range(L - 1):
        # the decision the kernels take is fmaf(z, scale, shift) > 0 in fp32; float64 evaluates its sign exactly
        C = pairs[l][0].out_channels
        z = saved[offs[5 * l]: offs[5 * l] + B * N * C * 4].view(torch.float32).view(B, N, C).double()
        bnp = saved[offs[5 * l + 3]: offs[5 * l + 3] + 4 * C * 4].view(torch.float32).view(4, C).double()
        masks.append(z * bnp[0] + bnp[1] > 0)
    C = pairs[-1][0].out_channels
    keys = saved[offs[5 * L]: offs[5 * L] + B * C * 8].view(torch.int64).view(B, C)
    argmax = 0xFFFFFFFF - (keys & 0xFFFFFFFF)
    positive = (keys >> 32) > 0
    return masks, argmax, positive
