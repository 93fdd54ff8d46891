"""Batch-sharded data parallelism for the hot path: one process per GPU, cloud pairs split
along the batch dimension, no data-path collective in the Chamfer forward/backward or the encoder.  NCCL
(over NVLink 5 / NVSwitch) is used only for what the reference's single-process loop computes over the
whole batch: the logged/optimised loss scalar (train_rl_gan_net.py:236-242) and the autoencoder gradients.

The reference itself has no distributed code; these helpers are what a data-parallel launch of its
train_autoencoder_epoch (train_rl_gan_net.py:220-249) needs around the drop-ins.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, local_rank, world_size); a single process without those variables is (0, 0, 1) and does
    not create a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    if world > 1 and torch.cuda.is_available() and os.environ.get("RLG_NUMA_BIND", "1") != "0":
        bind_to_local_numa_node(local_rank)
    if world > 1 and torch.cuda.is_available():
        # NCCL kernels of the per-step collectives run beside the persistent Chamfer forward: leave them an SM
        from .chamfer import set_reserved_sms
        set_reserved_sms(int(os.environ.get("RLG_RESERVED_SMS", "1")))
    return rank, local_rank, world


def bind_to_local_numa_node(local_rank: int) -> Optional[int]:
    """Best effort: pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host
    buffers (Linux places pages on the node of the first touch).  With eight ranks feeding their GPUs from one host every
    step, pinned rings on the wrong socket send each transfer across the inter-socket link (the end-to-end path at 8 GPUs
    is host-bound: ~26 GB/s per GPU).  Returns the node, or None when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # nvml prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` cloud pairs for `rank`; the first n_items % world ranks get one
    extra pair."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def sharded_mean_loss(per_pair_local: torch.Tensor, global_batch: int,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Batch mean of a per-pair quantity whose pairs are sharded over ranks.

    Value: sum over ALL ranks' pairs / global_batch (one all-reduce of one fp32), identical on every rank,
    i.e. torch.mean(chamfer_distance(pred, target)) of utils/losses.py:75 over the un-sharded batch.
    Gradient: each rank back-propagates only its own pairs' share (d loss / d per_pair_local = 1/global_batch),
    so after `allreduce_gradients` the parameter gradients equal the single-process ones."""
    local = per_pair_local.sum() / float(global_batch)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    total = local.detach().clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return local + (total - local.detach())


def sharded_chamfer_loss(pred_local: torch.Tensor, target_local: torch.Tensor, global_batch: int,
                         bidirectional: bool = True, group: Optional[dist.ProcessGroup] = None,
                         pair_fn: Optional[Callable] = None) -> torch.Tensor:
    """ChamferLoss (utils/losses.py:62-75) over a batch whose pairs are sharded across ranks.
    `pair_fn(pred, target, bidirectional) -> (B_local,)` defaults to the CUDA chamfer_distance."""
    if pair_fn is None:
        from .chamfer import chamfer_distance as pair_fn
    return sharded_mean_loss(pair_fn(pred_local, target_local, bidirectional), global_batch, group)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        bucket_bytes: int = 32 << 20) -> int:
    """Sum the .grad of `params` across ranks in flat buckets (AE gradients: 7-8 MB fp32, SURVEY.md 8e -> one
    bucket; NVSwitch all-reduce cost is latency- not link-bound, so buckets are sized for launch count).
    Returns the number of all-reduce calls issued."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        calls += 1
        bucket, size = [], 0

    for g in grads:
        nb = g.numel() * g.element_size()
        if bucket and (size + nb > bucket_bytes or g.dtype != bucket[0].dtype):
            flush()
        bucket.append(g)
        size += nb
    flush()
    return calls


def gather_rows(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather equally sized per-rank row blocks (GFVs (B/G,latent) or rewards (B/G,)) along dim 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
