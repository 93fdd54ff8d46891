#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): batched Chamfer distance forward + backward, B=32 pairs per
GPU, N=M=2048, synthetic unit-sphere clouds, fp32.  One step = ChamferLoss(pred, target) and its backward over one
batch (two kernel launches of this library).  Multi-GPU: every rank runs its own batch of 32 pairs (weak scaling, no
data-path collective); the logged loss scalar is all-reduced (NCCL) EVERY step inside the timed region, captured in the
CUDA graph on a side stream.

Prints ONE JSON line on rank 0: metric/value (device-resident inputs, CUDA events, max over ranks), roofline of the
dominant kernel, cpu_baseline, e2e (pinned host buffers through the public API: H2D of every step's inputs + D2H of
the loss), clocks, gpu_launches, and extra keys: `uniform` (same workload on uniform clouds), `strong_scaling`
(the 32-pair batch split over the ranks), `torch_cuda` (the reference's op sequence on stock torch CUDA kernels on the
same GPU: the kernels to beat), `encoder` / `encoder_config_dims` (clouds/s), `reward_loop` and `env_step` (cfg4: the
reward formula alone and the whole batched environment step), `large_cloud` (cfg5), `ae_step_b16` / `ae_step_b32` (cfg1 at
the reference's dims: the autoencoder training step, with the gradient all-reduce inside the step at n_gpus > 1 and the same
step on stock torch CUDA kernels beside it), `input_pipeline` (device-side batch preparation) -- all aggregated over the
ranks (barrier + MAX time).  Progress goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B, N, M = 32, 2048, 2048                 # cfg2
ENC_B, ENC_N, ENC_DIMS = 256, 2048, [64, 128, 1024]   # cfg3
CFG_DIMS = [64, 128, 128, 256, 128]      # the reference's own encoder_dims (configs/config.yaml:9-11)
RING_BYTES = 288 << 20                   # inputs cycled through a ring larger than the 126 MB L2
FLOP_PER_PAIR = 8.0 * N * M              # every pairwise squared distance counted once
BWD_BYTES_PER_PAIR = 56.0 * (N + M)
WORKLOAD = "chamfer_fwd_bwd B=32 N=M=2048 sphere (BASELINE configs[1])"


_T0 = time.perf_counter()


def log(msg: str) -> None:
    """Progress on stderr (stdout carries the one JSON line): shows where a multi-GPU run is when it is cut off."""
    print(f"[bench r{os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (encoder, cfg4, cfg5, torch_cuda)")
    ap.add_argument("--no-encoder", action="store_true", help="alias of --no-extras")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def base_config(world: int) -> dict:
    """The workload description shared, key for key, by both arms."""
    return {"workload": WORKLOAD, "B_per_gpu": B, "N": N, "M": M, "global_batch": B * world,
            "parallelism": f"batch-sharded dp{world}"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def profile_traffic(kernel: str):
    """dram__bytes_read+write per launch of `kernel` from the committed ncu summary (profiles/r2_traffic.json, written by
    tools/export_profile.py from an `ncu --set full` capture); None when no capture of this build is committed."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# synthetic inputs: CPU generator -> identical bits everywhere
# ----------------------------------------------------------------------------------------------------
def sphere(gen, b, n):
    x = torch.randn(b, n, 3, generator=gen)
    return (x / x.norm(dim=2, keepdim=True).clamp_min(1e-12)).contiguous()


def uniform(gen, b, n):
    return (torch.rand(b, n, 3, generator=gen) * 2.0 - 1.0).contiguous()


def make_ring(rank: int, slots: int, kind: str = "sphere", b: int = B):
    gen = torch.Generator(device="cpu").manual_seed(1234 + 2 + 1000 * rank + (500 if kind == "uniform" else 0))
    mk = sphere if kind == "sphere" else uniform
    return [(mk(gen, b, N), mk(gen, b, M)) for _ in range(slots)]


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU op sequence (oracle port) on this host's cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step(pc1, pc2):
    from oracle import oracle as O
    a = pc1.clone().requires_grad_(True)
    b = pc2.clone().requires_grad_(True)
    loss = O.ref_port_chamfer_loss(a, b)       # utils/losses.py:62-75 as written (torch.cdist -> min -> mean)
    loss.backward()
    return float(loss.item())


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port of utils/losses.py:13-75 run with
    all host threads).  EXACTLY K timed steps; each step is ChamferLoss forward+backward on a bounded sample of the
    B=32 batch (b pairs, b chosen from a calibration step so the whole run stays within ~2 minutes)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    gen = torch.Generator(device="cpu").manual_seed(1234 + 2)
    pc1, pc2 = sphere(gen, B, N), sphere(gen, B, M)
    K, W = max(1, args.steps), max(1, args.warmup)
    cpu_reference_step(pc1[:2], pc2[:2])
    t0 = time.perf_counter()
    cpu_reference_step(pc1[:4], pc2[:4])
    per_pair = (time.perf_counter() - t0) / 4
    b = int(max(1, min(B, 100.0 / ((K + W) * per_pair))))
    for k in range(W):
        o = (k * b) % (B - b + 1)
        cpu_reference_step(pc1[o:o + b], pc2[o:o + b])
    t0 = time.perf_counter()
    for k in range(K):
        o = (k * b) % (B - b + 1)
        cpu_reference_step(pc1[o:o + b], pc2[o:o + b])
    dt = time.perf_counter() - t0
    value = K * b / dt
    sample = (f"{K} steps, each ChamferLoss fwd+bwd on {b} of the {B} pairs of the batch, N=M={N} "
              f"(torch CPU ops of the reference, {torch.get_num_threads()} threads)")
    print(json.dumps({
        "impl": "reference", "metric": "chamfer_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(base_config(1), sampled_pairs_per_step=b),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------
# helpers of our arm
# ----------------------------------------------------------------------------------------------------
class Dist:
    """World bookkeeping: barrier + MAX-over-ranks of a device-timed duration."""

    def __init__(self, world, dev):
        import torch.distributed as dist
        self.dist, self.world, self.dev = dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_ms(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps: int, warm: int = 3) -> float:
        """ms per call of fn(k): warm-up, barrier, CUDA events around `reps` calls, MAX over ranks."""
        for k in range(warm):
            fn(k)
        torch.cuda.synchronize()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            fn(k)
        e1.record()
        torch.cuda.synchronize()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)) / reps


def flush_l2(dev):
    """Write a buffer larger than the 126 MB L2 so that nothing timed afterwards starts cache-resident."""
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    buf.zero_()
    torch.cuda.synchronize()
    del buf


def timed_step_graphs(P, D, ring, K, W, world, loss_hook_factory):
    """EXACTLY K timed ChamferLoss forward+backward steps over the ring, in CUDA graphs of up to len(ring) steps.
    Every graph that will be timed is replayed once beforehand (upload + warm-up), then L2 is flushed.
    Returns (ms, n_launches, last_loss, info)."""
    slots = len(ring)
    S = min(K, slots)
    n_full, rem = divmod(K, S)
    hook = loss_hook_factory() if world > 1 else None
    full = P.ChamferStepGraph(ring[:S], after_step=hook)
    tail = P.ChamferStepGraph(ring[:rem], after_step=hook) if rem else None
    n_launches = n_full * full.kernel_launches_per_replay + (tail.kernel_launches_per_replay if tail else 0)
    warm_replays = max(1, -(-W // S))
    for _ in range(warm_replays):
        full.replay()
    if tail is not None:
        tail.replay()
    torch.cuda.synchronize()
    flush_l2(ring[0][0].device)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_full):
        full.replay()
    if tail is not None:
        tail.replay()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_ms(e0.elapsed_time(e1))
    last = float((tail.losses[-1] if tail else full.losses[-1]).item())
    touched = min(K, slots)
    info = {"steps_per_graph": S, "graph_replays": n_full + (1 if tail else 0), "warmup_steps": warm_replays * S + rem,
            "ring_slots": slots, "slots_touched_per_pass": touched,
            "l2_policy": (f"every timed graph replayed once beforehand, then L2 flushed (256 MiB written); steps read "
                          f"{touched} distinct batches = {touched * (ring[0][0].numel() + ring[0][1].numel()) * 4 >> 20} MiB per "
                          f"pass through the ring" + (" (> 126 MB L2, cycled)" if touched == slots else
                                                      " (each batch read once: never cache-resident)"))}
    return ms, n_launches, last, info


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import gan_rl_3d_b200 as rlg
    import importlib
    Dm = importlib.import_module("gan-rl_3d_b200.distributed")
    _lib = importlib.import_module("gan-rl_3d_b200._lib")
    import torch.distributed as dist

    rank, local_rank, world = Dm.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    peaks, peaks_src = load_peaks()
    K, W = max(1, args.steps), max(3, args.warmup)
    D = Dist(world, dev)
    extras_on = not (args.no_extras or args.no_encoder)

    P = importlib.import_module("gan-rl_3d_b200.pipeline")
    slot_bytes = (B * N + B * M) * 3 * 4
    slots = max(8, RING_BYTES // slot_bytes)
    ring_host = make_ring(rank, slots)
    ring = [(a.to(dev), b.to(dev)) for a, b in ring_host]

    # The logged loss scalar (train_rl_gan_net.py:241-247) is the only cross-rank quantity of this step: one fp32
    # all-reduce per step, captured in the graph on a side stream.
    red_bufs = []

    def loss_hook_factory():
        def hook(loss):
            buf = loss.detach().reshape(1).clone()
            red_bufs.append(buf)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return hook

    collective = {"kind": "none (1 GPU)", "us_per_step": 0.0}
    if world > 1:
        for _ in range(3):                                  # NCCL communicator warm-up outside any capture
            t_ = torch.zeros(1, device=dev)
            dist.all_reduce(t_)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    log("timed step graphs")
    try:
        ms, n_launches, last_loss, info = timed_step_graphs(P, D, ring, K, W, world, loss_hook_factory)
        if world > 1:
            collective["kind"] = "ncclAllReduce(sum) of the fp32 loss scalar, EVERY step, captured in the CUDA graph on a side stream"
    except Exception as e:                                  # NCCL capture unsupported: per-replay all-reduce instead
        if world == 1:
            raise
        red = torch.zeros(1, device=dev)

        def factory_none():
            return None
        ms, n_launches, last_loss, info = timed_step_graphs(P, D, ring, K, W, 1, factory_none)
        dist.all_reduce(red)
        collective["kind"] = f"capture of the NCCL all-reduce failed ({type(e).__name__}); NO collective inside the timed region"
    value = K * B * world / (ms * 1e-3)

    log(f"headline done: {value:.0f} pairs/s")
    if world > 1:
        # the collective alone: S captured all-reduces back to back
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        buf = torch.zeros(1, device=dev)
        s.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(s):
                dist.all_reduce(buf)
                s.synchronize()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(64):
                        dist.all_reduce(buf)
            torch.cuda.synchronize()
            cms = D.timed(lambda k: g.replay(), 5, 2)
            collective["us_per_step"] = cms * 1e3 / 64
            collective["note"] = ("latency of one 4-byte all-reduce when issued back to back alone; inside the step it runs "
                                  "on a side stream and overlaps the next step's kernels")
        except Exception as e:
            collective["us_per_step"] = None
            collective["note"] = f"stand-alone timing failed: {type(e).__name__}"

    log("e2e")
    # ---- e2e: host buffers through the public API, H2D + D2H of every step inside the timed region --
    nb = min(slots, 32)
    pinned = [P.pin_pair(a, b) for a, b in ring_host[:nb]]      # (pred, target) of a step: one pinned buffer, one transfer
    host_graph = P.HostChamferStepGraph(pinned, dev)
    Ke_replays = max(1, min(K, 2000) // nb)
    Ke = Ke_replays * nb
    host_graph.replay()
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke_replays):
        host_graph.replay()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = D.max_ms(e2e_s * 1e3) * 1e-3
    e2e_value = Ke * B * world / e2e_s
    e2e_loss_check = float(host_graph.losses_host[0])
    clocks = sampler.stop() if rank == 0 else None
    del host_graph, pinned

    extra = {}
    log("extras")
    # ---- same workload on uniform clouds (SURVEY.md 8d names both distributions) --------------------
    if extras_on:
        uring = [(a.to(dev), b.to(dev)) for a, b in make_ring(rank, 64, "uniform")]
        ums, _, uloss, _ = timed_step_graphs(P, D, uring, min(K, 640), 3, 1, lambda: None)
        extra["uniform"] = {"metric": "chamfer_pairs_per_s", "value": min(K, 640) * B * world / (ums * 1e-3), "unit": "pairs/s",
                            "config": {"workload": "chamfer_fwd_bwd B=32 N=M=2048 uniform [-1,1]^3 clouds, per GPU"},
                            "ms_per_step": ums / min(K, 640), "last_loss": uloss}
        del uring
        # ---- strong scaling of the headline batch: the SAME 32 pairs split over the ranks -----------
        bs = max(1, B // world)
        sring = [(a[:bs].contiguous(), b[:bs].contiguous()) for a, b in ring[:64]]
        sms, _, _, _ = timed_step_graphs(P, D, sring, min(K, 640), 3, 1, lambda: None)
        extra["strong_scaling"] = {"metric": "chamfer_pairs_per_s", "value": min(K, 640) * bs * world / (sms * 1e-3),
                                   "unit": "pairs/s", "scaling": "strong", "pairs_per_gpu": bs,
                                   "config": {"workload": f"the headline batch of {B} pairs split over {world} GPU(s): "
                                                          f"{bs} pairs per GPU per step"},
                                   "ms_per_step": sms / min(K, 640),
                                   "note": "launch- and latency-bound below ~16 pairs per GPU: a 148-CTA persistent kernel "
                                           "over 2*pairs*16 query blocks has less than one wave of work"}
        del sring

    if extras_on:
        log("encoder_measurement")
        extra.update(encoder_measurement(rlg, dev, D, peaks, peaks_src, ENC_DIMS, "encoder",
                                         "PointNet encoder 3->64->128->1024 + max-pool + GFV head, B=256 per GPU, N=2048 "
                                         "(BASELINE configs[2])"))
        log("encoder_measurement")
        extra.update(encoder_measurement(rlg, dev, D, peaks, peaks_src, CFG_DIMS, "encoder_config_dims",
                                         "PointNet encoder with the reference's own encoder_dims 3->64->128->128->256->128 "
                                         "(configs/config.yaml:9-11) + max-pool + GFV head, B=256 per GPU, N=2048"))
        log("reward_measurement")
        extra.update(reward_measurement(rlg, dev, D))
        log("env_step_measurement")
        extra.update(env_step_measurement(rlg, dev, D))
        log("large_cloud_measurement")
        extra.update(large_cloud_measurement(rlg, dev, D))
        log("input_pipeline_measurement")
        extra.update(input_pipeline_measurement(rlg, dev, D, not args.no_cpu_baseline))
        log("ae_step_measurement")
        extra.update(ae_step_measurement(rlg, dev, D, rank))

    if rank != 0:
        # the other ranks have nothing left to measure: meet rank 0 at its final barrier and leave WITHOUT tearing the
        # communicator down (graphs holding captured NCCL collectives are still alive; destroy_process_group hangs under them)
        log("done, waiting for rank 0")
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    # ---- roofline of the dominant kernel: the fused Chamfer forward, timed alone with CUDA events ----
    a, b = ring[0]
    a = a.detach(); b = b.detach()
    d1, d2, i1, i2, m1, m2 = rlg.chamfer_nearest(a, b)
    ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_forward(algo_flag, tile_only):
        """ms per launch sequence of the forward on ring inputs; the workspace is re-initialised (memset node) after a
        call that leaves it dirty, and that memset is timed separately and subtracted."""
        def call(x, y, extra_flags):
            rc = lib.rlg_chamfer_fwd(x.data_ptr(), y.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                                     i2.data_ptr(), m1.data_ptr(), m2.data_ptr(), ws.data_ptr(), ws.numel(),
                                     _lib.CHAMFER_WS_CLEAN | algo_flag | extra_flags, stream)
            _lib.check("rlg_chamfer_fwd", rc)

        def loop(reps_, flags, refill):
            ws.fill_(0xFF)
            for k in range(10):
                call(*[t_.detach() for t_ in ring[k % slots]], flags)
                if refill:
                    ws.fill_(0xFF)
            torch.cuda.synchronize()
            e0.record()
            for k in range(reps_):
                x, y = ring[k % slots]
                call(x.detach(), y.detach(), flags)
                if refill:
                    ws.fill_(0xFF)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps_
        if not tile_only:
            return loop(400, 0, False)
        e0.record()
        for _ in range(400):
            ws.fill_(0xFF)
        e1.record()
        torch.cuda.synchronize()
        fill_ms = e0.elapsed_time(e1) / 400
        return loop(400, _lib.CHAMFER_TILE_ONLY, True) - fill_ms

    use_tensor = rlg.get_default_sweep() == "tensor" or (rlg.get_default_sweep() == "auto" and N >= 64 and M >= 64)
    tc_fwd_ms = time_forward(_lib.CHAMFER_ALGO_TENSOR, False)
    fp_sweep_ms = time_forward(0, True)
    fp_fwd_ms = time_forward(0, False)
    ws.fill_(0xFF)
    peak = ctypes_float6(lib, dev)
    fp32_theory = peak[2]
    peak_src = (f"theoretical FP32 FMA: {int(peak[3])} SMs x 128 lanes x 2 flop x {peak[1]:.0f} MHz "
                "(MEASURED_PEAKS.json has no FP32 entry; north_star names the FFMA peak)")
    flops = FLOP_PER_PAIR * B
    # what the fused tensor kernel can at best reach: its minima cost one dispatch slot per (query, candidate, direction)
    # value and lane -- 4 sub-partitions x 32 lanes per SM and clock -- i.e. 64 pairs per SM and clock
    issue_bound_tflops = int(peak[3]) * 64 * peak[1] * 1e6 * 8.0 / 1e12
    fp32_sweep = {"bound": "fp32", "kernel": "chamfer_filter_kernel<16> (pair sweep of the FP32 path, alone)",
                  "achieved": flops / (fp_sweep_ms * 1e-3) / 1e12, "peak": fp32_theory, "unit": "TFLOP/s",
                  "frac": flops / (fp_sweep_ms * 1e-3) / 1e12 / fp32_theory, "traffic": profile_traffic("chamfer_filter_kernel"),
                  "launch_us": fp_sweep_ms * 1e3, "forward_us": fp_fwd_ms * 1e3,
                  "note": "RLG_CHAMFER_SWEEP=fp32: 3 FFMA + 1 FADD + 2 min slots per pair through one dispatch port per SM "
                          "sub-partition cap the FFMA pipe near 60 % (DESIGN.md 3.1)"}
    if use_tensor:
        tf = flops / (tc_fwd_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32", "kernel": "chamfer_tcsweep_kernel (the whole forward: pair sweep + exact refinement + means + loss, ONE launch)",
                    "achieved": tf, "peak": fp32_theory, "unit": "TFLOP/s", "frac": tf / fp32_theory,
                    "traffic": profile_traffic("chamfer_tcsweep_kernel"), "peak_source": peak_src,
                    "peak_measured_ffma": peak[0], "frac_of_measured_ffma": tf / peak[0] if peak[0] else None,
                    "launch_us": tc_fwd_ms * 1e3, "algorithmic_flop_per_launch": flops,
                    "issue_bound_tflops": issue_bound_tflops, "frac_of_issue_bound": tf / issue_bound_tflops,
                    "note": "algorithmic 8 flop per point pair against the FP32 FFMA peak north_star names.  The kernel runs the "
                            "3-term contraction on the tensor pipe (tcgen05 kind::tf32, split-tf32 operands) and only the "
                            "minima and the exact refinement on the CUDA cores; the bound it can reach is the dispatch rate of "
                            "the min reduction (issue_bound_tflops: one slot per value and lane), not the FFMA pipe",
                    "fp32_sweep": fp32_sweep}
        roofline_fwd = dict(roofline, kernel="chamfer_distance_l2 forward (one launch)")
    else:
        tf = flops / (fp_fwd_ms * 1e-3) / 1e12
        roofline = dict(fp32_sweep, peak_source=peak_src, peak_measured_ffma=peak[0], algorithmic_flop_per_launch=flops)
        roofline_fwd = {"bound": "fp32", "kernel": "chamfer_distance_l2 forward (pair sweep + refinement kernel)", "achieved": tf,
                        "peak": fp32_theory, "unit": "TFLOP/s", "frac": tf / fp32_theory, "traffic": None,
                        "launch_us": fp_fwd_ms * 1e3}

    # backward: HBM-bound by bytes, latency-bound at this size.  One stand-alone call per ring slot (two memset nodes + the
    # kernel; inside a training step the forward zero-fills and the backward is the kernel alone), the calls of 48 slots
    # (inputs + saved forward results + gradients = 3.6 MB each, 173 MB > L2) captured in one CUDA graph and replayed.
    g = torch.full((B,), 0.5 / B, device=dev)
    n_bwd = min(48, len(ring))
    slots = []
    for k in range(n_bwd):
        x, y = ring[k]
        x, y = x.detach(), y.detach()
        slots.append((x, y) + tuple(rlg.chamfer_nearest(x, y)[:4]) + (torch.empty_like(x), torch.empty_like(y)))
    bwd_us = {}
    cap = torch.cuda.Stream()
    for name, det in (("atomics", False), ("deterministic", True)):
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            x, y, s1, s2, j1, j2, gx, gy = slots[0]
            rlg.chamfer_backward(x, y, s1, s2, j1, j2, g, g, out=(gx, gy), deterministic=det)
            bwd_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(bwd_graph, stream=cap):
                for x, y, s1, s2, j1, j2, gx, gy in slots:
                    rlg.chamfer_backward(x, y, s1, s2, j1, j2, g, g, out=(gx, gy), deterministic=det)
        torch.cuda.synchronize()
        for _ in range(3):
            bwd_graph.replay()
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            bwd_graph.replay()
        e1.record()
        torch.cuda.synchronize()
        bwd_us[name] = e0.elapsed_time(e1) * 1e3 / (reps * n_bwd)
        del bwd_graph
    bwd_ms = bwd_us["atomics"] * 1e-3
    bwd_gbs = BWD_BYTES_PER_PAIR * B / (bwd_ms * 1e-3) / 1e9
    roofline_bwd = {"bound": "hbm", "kernel": "memset x2 + chamfer_bwd_kernel (stand-alone call; inside a training step the "
                    "forward zero-fills and the backward is the single kernel)", "achieved": bwd_gbs,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bwd_gbs / peaks["hbm_gbs"],
                    "traffic": profile_traffic("chamfer_bwd_kernel"),
                    "peak_source": f"{peaks_src} MEASURED_PEAKS.json hbm_gbs", "launch_us": bwd_us["atomics"],
                    "deterministic_launch_us": bwd_us["deterministic"],
                    "note": "7.3 MB per call: latency bound at this shape (two dependent memory round trips per point); "
                            f"{n_bwd} calls on distinct slots per CUDA graph, graph replays timed; deterministic = the run-to-run "
                            "reproducible fixed-point variant (memset + 2 kernels)"}
    del slots

    if extras_on:
        extra.update(torch_cuda_measurement(dev, ring))

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        pc1, pc2 = ring_host[0]
        cpu_reference_step(pc1, pc2)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0 and n < 40:
            cpu_reference_step(pc1, pc2)
            n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n * B / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} steps of ChamferLoss fwd+bwd, B={B}, N=M={N}, torch CPU ops as the reference "
                                  f"runs them (oracle port), {dt:.1f} s"}

    cfg = base_config(world)
    cfg.update(info)
    cfg.update({"step": "ChamferLoss forward + backward (autograd), S steps per CUDA graph; at n_gpus > 1 the loss scalar is "
                        "all-reduced every step inside the graph (side stream)", "last_loss": last_loss,
                "collective": collective,
                "pair_sweep": "tensor (tcgen05 kind::tf32 contraction + CUDA-core minima, refinement fused)" if use_tensor
                else "fp32 (FFMA pipe)"})
    out = {
        "metric": "chamfer_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "roofline": roofline, "roofline_fwd": roofline_fwd, "roofline_bwd": roofline_bwd, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": slot_bytes, "d2h_bytes_per_step": 4,
                "steps": Ke, "loss_step0": e2e_loss_check,
                "how": "HostChamferStepGraph: per step ONE H2D transfer of the pinned (pred,target) buffer -> ChamferLoss -> "
                       "backward -> D2H of the loss (own stream), 32 steps per CUDA-graph replay, copies double-buffered "
                       "against the previous step's kernels; wall clock around replays + synchronize, max over ranks"},
        "gpu_launches": n_launches, "clocks": clocks,
    }
    out.update(extra)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL collectives are still alive here; tearing the communicator down under them hangs
        # (ProcessGroupNCCL watchdog, observed on 2 GPUs).  Everything has been measured and printed: meet once more and
        # leave without the teardown.
        log("done, leaving")
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def ctypes_float6(lib, dev):
    import ctypes
    out = (ctypes.c_float * 6)()
    scratch = torch.zeros(64, device=dev)
    rc = lib.rlg_fp32_peak(out, 6, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"rlg_fp32_peak rc={rc}: {lib.rlg_last_error()}")
    return [float(v) for v in out]


def randomize_bn(module, seed=0):
    """Non-trivial BatchNorm statistics/affine -- a fresh BN is identity-like."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.5)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 1.7 + 0.3)
                m.weight.copy_(1.0 + torch.randn(m.num_features, generator=g) * 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.3)


def encoder_measurement(rlg, dev, D, peaks, peaks_src, dims, key, workload):
    """Encoder clouds/s through the public module call `enc(x)` in eval mode, every rank on its own 256 clouds (weak
    scaling), aggregated over ranks: the fastest available path (tensor cores where the widths allow) as the headline
    of the key, the fp32 path beside it, and an e2e figure from pinned host clouds."""
    torch.manual_seed(0)
    enc = rlg.PointNetEncoder(3, 128, dims)
    randomize_bn(enc, 0)
    enc = enc.eval().to(dev)
    gen = torch.Generator(device="cpu").manual_seed(1234 + 3 + 1000 * D.dist.get_rank() if D.world > 1 else 1234 + 3)
    xs_host = [sphere(gen, ENC_B, ENC_N).pin_memory() for _ in range(24)]      # 24 x 6.3 MB = 151 MB > L2
    xs = [x.to(dev) for x in xs_host]
    widths = [3] + list(dims)
    flop = 2.0 * ENC_N * sum(widths[i] * widths[i + 1] for i in range(len(dims))) * ENC_B
    res = {}
    for precision, reps in (("bf16", 120), ("auto", 60), ("fp32", 12)):
        enc.rlg_precision = precision

        def call(k):
            with torch.no_grad():
                return enc(xs[k % len(xs)])
        try:
            call(0)
            path = rlg.encoder_path_of(enc) if hasattr(rlg, "encoder_path_of") else precision
        except Exception as e:
            res[precision] = {"error": f"{type(e).__name__}: {e}"[:200]}
            continue
        ms_eager = D.timed(call, reps)
        # the same calls captured in CUDA graphs of 24 (one per input batch), like the headline's step graphs: a GFV-extraction
        # loop is launch-bound when issued call by call (trunk kernel(s) + the three kernels of the global MLP head)
        ms, ms_trunk = ms_eager, None
        try:
            import importlib
            Enc = importlib.import_module("gan-rl_3d_b200.encoder")
            gs = torch.cuda.Stream(dev)
            gs.wait_stream(torch.cuda.current_stream(dev))
            graphs = []
            for fn in (lambda x: enc(x), lambda x: Enc._trunk_pool(enc, x)):
                with torch.cuda.stream(gs), torch.no_grad():
                    fn(xs[0])
                gs.synchronize()
                g = torch.cuda.CUDAGraph()
                outs = []
                with torch.cuda.graph(g, stream=gs), torch.no_grad():
                    for x in xs:
                        outs.append(fn(x))
                graphs.append((g, outs))
            torch.cuda.current_stream(dev).wait_stream(gs)
            n_rep = max(1, reps // len(xs))
            ms = D.timed(lambda k: graphs[0][0].replay(), n_rep, 1) / len(xs)
            ms_trunk = D.timed(lambda k: graphs[1][0].replay(), n_rep, 1) / len(xs)
            del graphs
        except Exception as e:                                  # capture unsupported for this path: the eager number stands
            res.setdefault("notes", []).append(f"{precision}: graph capture failed ({type(e).__name__})")
        res[precision] = {"ms": ms, "ms_eager": ms_eager, "ms_trunk_only": ms_trunk,
                          "tflops": flop / ((ms_trunk or ms) * 1e-3) / 1e12, "path": path}
    best = "bf16" if "ms" in res.get("bf16", {}) else "fp32"
    enc.rlg_precision = best
    # e2e: pinned host clouds -> H2D -> enc(x) -> GFV D2H, every step, double-buffered on two streams
    gfv_host = torch.empty(len(xs_host), ENC_B, 128).pin_memory()
    copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage = [torch.empty_like(xs[0]) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_pass(n):
        for k in range(n):
            s = k % 2
            with torch.cuda.stream(copy_s):
                if k >= 2:
                    copy_s.wait_event(consumed[s])
                stage[s].copy_(xs_host[k % len(xs_host)], non_blocking=True)
                copied[s].record(copy_s)
            with torch.cuda.stream(comp_s), torch.no_grad():
                comp_s.wait_event(copied[s])
                gfv = enc(stage[s])
                consumed[s].record(comp_s)
                gfv_host[k % len(xs_host)].copy_(gfv, non_blocking=True)
        copy_s.synchronize()
        comp_s.synchronize()

    e2e_pass(4)
    D.barrier()
    n = 48 if best == "bf16" else 8
    t0 = time.perf_counter()
    e2e_pass(n)
    e2e_s = D.max_ms((time.perf_counter() - t0) * 1e3) * 1e-3
    ms, tf = res[best]["ms"], res[best]["tflops"]                # tf: the trunk kernels alone (algorithmic flop / their time)
    tensor = best == "bf16"
    peak = peaks["bf16_tflops_sustained"] if tensor else None
    out = {"metric": "encoder_clouds_per_s", "value": ENC_B * D.world / (ms * 1e-3), "unit": "clouds/s", "n_gpus": D.world,
           "scaling": "weak", "config": {"workload": workload, "l2_policy": "24 input batches = 151 MB cycled",
                                        "step": "enc(x) in eval mode, 24 calls per CUDA graph (ms_eager: issued call by call)"},
           "dtype": best, "path": res[best]["path"], "ms_per_step": ms,
           "roofline": {"bound": "tensor", "kernel": res[best]["path"], "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                        "frac": tf / peak if peak else None, "traffic": profile_traffic("encoder_tc_kernel"),
                        "peak_source": f"{peaks_src} MEASURED_PEAKS.json bf16_tflops_sustained" if peak else
                        "fp32 CUDA-core path: no tensor roofline applies", "algorithmic_flop_per_launch": flop},
           "e2e": {"value": n * ENC_B * D.world / e2e_s, "unit": "clouds/s", "h2d_bytes_per_step": ENC_B * ENC_N * 12,
                   "d2h_bytes_per_step": ENC_B * 128 * 4, "steps": n},
           "paths": res,
           "note": "paths: bf16 = bf16 tensor-core GEMMs (2e-2 class); auto = the default precision of the drop-in (fp32-grade "
                   "tensor-core GEMMs on fp16 hi+lo operand pairs where the widths allow); fp32 = the CUDA-core kernel"}
    del xs, xs_host
    return {key: out}


def reward_measurement(rlg, dev, D):
    """BASELINE configs[3]: reward evaluation for E=1024 episodes (decoder output vs complete cloud, N=M=2048) in one
    batched forward-only pass, the 1024 episodes sharded over the ranks (strong scaling: 1024/n_gpus per GPU)."""
    E_total, n = 1024, 2048
    E = max(1, E_total // D.world)
    rank = D.dist.get_rank() if D.world > 1 else 0
    gen = torch.Generator(device="cpu").manual_seed(1234 + 4 + 1000 * rank)
    batches = []
    for _ in range(max(3, (140 << 20) // (E * n * 24) + 1)):        # > L2 in total
        batches.append((sphere(gen, E, n).to(dev), sphere(gen, E, n).to(dev), torch.rand(E, 128, generator=gen).to(dev),
                        torch.rand(E, 128, generator=gen).to(dev), torch.randn(E, 1, generator=gen).to(dev)))
    out = {}

    def call(k):
        out["r"] = rlg.batched_rewards(*batches[k % len(batches)])
    ms = D.timed(call, 12)
    tf = 8.0 * n * n * E * D.world / (ms * 1e-3) / 1e12
    return {"reward_loop": {"metric": "reward_evals_per_s", "value": E * D.world / (ms * 1e-3), "unit": "episodes/s",
                            "n_gpus": D.world, "scaling": "strong", "episodes_per_gpu": E, "ms_per_step": ms,
                            "config": {"workload": "RewardFunction for E=1024 episodes sharded over the GPUs, Chamfer N=M=2048 forward "
                                                   "only + GFV MSE + discriminator term (BASELINE configs[3])"},
                            "chamfer_tflops_algorithmic": tf, "last_reward_mean": float(out["r"].mean().item())}}


def env_step_measurement(rlg, dev, D):
    """SURVEY 8(f)-1 / BASELINE configs[3]: the WHOLE environment step for E=1024 episodes (sharded over the GPUs): action z ->
    latent-GAN generator -> decoder -> 2048-point completion; discriminator logit; Chamfer against the complete cloud; GFV MSE
    against the complete cloud's GFV (encoder, once per reset) -> (E,) rewards, device-resident, one CUDA-graph replay per step.
    Reference-shaped modules at the reference's dims (models/latent_gan.py, models/autoencoder.py), random weights, eval."""
    import importlib
    Env = importlib.import_module("gan-rl_3d_b200.environment").BatchedRLEnvironment
    AE = importlib.import_module("gan-rl_3d_b200.ae_step")
    nn = torch.nn
    E_total, n = 1024, 2048
    E = max(1, E_total // D.world)
    rank = D.dist.get_rank() if D.world > 1 else 0
    torch.manual_seed(0)
    ae = AE.PointCloudAutoencoder().to(dev).eval()
    gen_dims, disc_dims = [256, 512, 512, 256, 128], [128, 256, 512, 256, 1]
    seq, c_in = [], 1
    for c in gen_dims[:-1]:
        seq += [nn.Linear(c_in, c), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
        c_in = c
    generator = nn.Sequential(*seq, nn.Linear(c_in, gen_dims[-1]), nn.Tanh()).to(dev).eval()
    seq, c_in = [], 128
    for c in disc_dims[:-1]:
        seq += [nn.utils.spectral_norm(nn.Linear(c_in, c)), nn.LayerNorm(c), nn.LeakyReLU(0.2, inplace=True), nn.Dropout(0.3)]
        c_in = c
    discriminator = nn.Sequential(*seq, nn.utils.spectral_norm(nn.Linear(c_in, disc_dims[-1]))).to(dev).eval()
    env = Env(ae.encode, generator, ae.decode, discriminator, dev, capture=True)
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    t0 = time.perf_counter()
    states = env.reset({"incomplete": sphere(g, E, 1400), "complete": sphere(g, E, n)})
    torch.cuda.synchronize()
    reset_ms = (time.perf_counter() - t0) * 1e3
    actions = [torch.randn(E, 1, generator=g).to(dev) for _ in range(4)]
    out = {}

    def call(k):
        out["r"] = env.step(actions[k % 4])[1]
    ms = D.timed(call, 12)
    return {"env_step": {"metric": "env_steps_per_s", "value": E * D.world / (ms * 1e-3), "unit": "episodes/s", "n_gpus": D.world,
                         "scaling": "strong", "episodes_per_gpu": E, "ms_per_step": ms, "reset_ms_first_call": reset_ms,
                         "config": {"workload": "batched RLGANNetEnvironment.step, E=1024 episodes sharded over the GPUs: generator + "
                                                "decoder + discriminator (stock torch MLPs) + Chamfer 2048x2048 forward + reward, one graph "
                                                "replay per step, no host synchronisation (models/rl_gan_net.py:299-339)"},
                         "last_reward_mean": float(out["r"].mean().item()), "state_shape": list(states.shape)}}


def input_pipeline_measurement(rlg, dev, D, cpu_leg: bool):
    """batches of 32 (complete, incomplete) cloud pairs out of a binary cache of 800 clouds (the reference's
    training split) -- incomplete-cloud creation, augmentation, normalisation, padding on the device (data.DeviceBatcher).
    value = clouds/s through make_batch() including the host's plan drawing and upload (what a training loop pays);
    device_ms = the two kernels alone.  cpu_baseline (rank 0): the same per-sample numeric work as the reference does it
    (oracle ports of utils/dataset.py:252-297,393-421 + data_utils.py:15-60, one process, text parsing NOT included)."""
    bsz, n = 32, 2048
    rank = D.dist.get_rank() if D.world > 1 else 0
    rng = np.random.default_rng(77 + rank)
    cache = rng.normal(size=(800, n, 3)).astype(np.float32)
    batcher = rlg.DeviceBatcher(cache, dev)
    plans = [rlg.draw_plan(rng, bsz, n, items=rng.integers(0, 800, bsz)) for _ in range(4)]
    for p in plans[:2]:
        batcher.make_batch(p)
    for _ in range(3):                 # the device-side draws (argsort of uniforms, randint, randn) load their kernels here
        batcher.make_batch(rlg.draw_plan(rng, bsz, n, items=rng.integers(0, 800, bsz), host_indices=False))
    torch.cuda.synchronize()
    D.barrier()
    reps = 40
    t0 = time.perf_counter()
    for k in range(reps):
        out = batcher.make_batch(rlg.draw_plan(rng, bsz, n, items=rng.integers(0, 800, bsz), host_indices=False))
    torch.cuda.synchronize()
    wall_ms = D.max_ms((time.perf_counter() - t0) * 1e3) / reps
    dev_ms = D.timed(lambda k: batcher.make_batch(plans[k % 4]), 8)        # pre-drawn plans: upload + kernels + 4-byte read-back
    entry = {"metric": "input_batches_clouds_per_s", "value": bsz * D.world / (wall_ms * 1e-3), "unit": "clouds/s", "n_gpus": D.world,
             "scaling": "weak", "ms_per_batch": wall_ms, "ms_per_batch_predrawn_plan": dev_ms, "batch": bsz,
             "incomplete_shape": list(out["incomplete_pc"].shape),
             "config": {"workload": "32-cloud training batches from an 800-cloud binary cache in HBM: incomplete-cloud creation + "
                                    "augmentation + normalisation + duplicate padding (utils/dataset.py:135-187,393-421), host plan "
                                    "drawing included"}}
    if cpu_leg and rank == 0:
        from oracle import oracle as O
        plan = plans[0]
        t0 = time.perf_counter()
        n_b = 0
        while time.perf_counter() - t0 < 3.0:
            inc = []
            for b in range(bsz):
                raw = cache[plan["item"][b]].astype(np.float64)
                draws = ({"method": 0, "keep_idx": plan["keep_idx"][b, :plan["n_keep"][b]]} if plan["method"][b] == 0 else
                         {"method": 1, "center": int(plan["center"][b]), "ratio": float(plan["ratio"][b])})
                part = O.ref_port_create_incomplete(raw, draws)
                for which, pc in enumerate((raw, part)):
                    noise = (np.clip(rng.normal(0.0, 0.01, (len(pc), 3)), -0.05, 0.05).astype(np.float32)
                             if plan["jitter_on"][which, b] else None)
                    aug = O.ref_port_normalize(O.ref_port_augment(pc, plan["rot"][which, b].reshape(3, 3), noise,
                                                                  float(plan["scale"][which, b])))
                inc.append(aug)
            m = max(len(p) for p in inc)
            O.ref_port_pad(inc, [plan["pad_idx"][b, :m - len(inc[b])] for b in range(bsz)])
            n_b += 1
        dt = time.perf_counter() - t0
        entry["cpu_baseline"] = {"value": n_b * bsz / dt, "unit": "clouds/s", "cores": 1, "kind": "port",
                                 "sample": f"{n_b} batches of {bsz} clouds, the reference's per-sample numeric work (no file parsing), {dt:.1f} s"}
    return {"input_pipeline": entry}


def large_cloud_measurement(rlg, dev, D):
    """BASELINE configs[4]: ChamferLoss forward + backward, B=64 pairs of N=M=16384 sharded over the ranks (64/n_gpus per
    GPU), the loss scalar all-reduced every step (NCCL)."""
    Bt, n = 64, 16384
    Bl = max(1, Bt // D.world)
    rank = D.dist.get_rank() if D.world > 1 else 0
    g = torch.Generator().manual_seed(1238 + 1000 * rank)
    ring = [(sphere(g, Bl, n).to(dev).requires_grad_(True), sphere(g, Bl, n).to(dev)) for _ in range(max(2, 160 // Bl))]
    crit = rlg.ChamferLoss()
    one = torch.ones((), device=dev)
    out = {}

    def step(k):
        a, b = ring[k % len(ring)]
        a.grad = None
        loss = crit(a, b)
        loss.backward(gradient=one)
        if D.world > 1:
            red = loss.detach().reshape(1).clone()
            D.dist.all_reduce(red)
        out["loss"] = loss

    ms = D.timed(step, 10)
    tf = 8.0 * n * n * Bl * D.world / (ms * 1e-3) / 1e12
    # the backward alone at this size (HBM bound by bytes: 56 per point), float atomics and the reproducible variant
    bwd = {}
    with torch.no_grad():
        saved = [rlg.chamfer_nearest(a.detach(), b)[:4] for a, b in ring]
    gg = torch.full((Bl,), 0.5 / Bl, device=dev)
    for name, det in (("atomics", False), ("deterministic", True)):
        def call(k):
            a, b = ring[k % len(ring)]
            rlg.chamfer_backward(a.detach(), b, *saved[k % len(ring)], gg, gg, deterministic=det)
        t = D.timed(call, 20)
        bwd[name] = {"us": t * 1e3, "algorithmic_gbs": 56.0 * 2 * n * Bl / (t * 1e-3) / 1e9}
    return {"large_cloud": {"backward_alone": bwd, "metric": "chamfer_pairs_per_s", "value": Bl * D.world / (ms * 1e-3), "unit": "pairs/s",
                            "n_gpus": D.world, "scaling": "strong", "pairs_per_gpu": Bl, "ms_per_step": ms,
                            "config": {"workload": "ChamferLoss fwd+bwd, B=64 pairs of N=M=16384 sharded over the GPUs, loss "
                                                   "all-reduce per step (BASELINE configs[4])"},
                            "algorithmic_tflops": tf, "last_loss": float(out["loss"].item())}}


def ae_step_measurement(rlg, dev, D, rank):
    """BASELINE configs[0]: the autoencoder training step (train_rl_gan_net.py:220-249) at the reference's own dims
    (configs/config.yaml: encoder [64,128,128,256,128], latent 128, decoder [256,256,6144], Adam lr 1e-3 wd 1e-5), incomplete
    clouds of 1400 points in, 2048-point reconstruction against the complete cloud.  Per GPU batch 16 (config_quick.yaml) and
    32 (config.yaml), weak scaling; at n_gpus > 1 the flat gradient all-reduce runs every step inside the captured graph.
    Rank 0 also times the same step on stock torch CUDA kernels (reference-shaped modules, cdist Chamfer, eager)."""
    import importlib
    import torch.distributed as dist
    AE = importlib.import_module("gan-rl_3d_b200.ae_step")
    world = D.world
    out = {}
    S = 8
    for bsz in (16, 32):
        torch.manual_seed(0)                                   # same initial weights on every rank
        model = AE.PointCloudAutoencoder().to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True, fused=True)
        gen = torch.Generator(device="cpu").manual_seed(4242 + rank)
        batches = [(sphere(gen, bsz, 1400).to(dev), sphere(gen, bsz, 2048).to(dev)) for _ in range(S)]
        g = AE.AEStepGraph(model, opt, batches, world=world)
        reps = 6
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        ms = D.max_ms(e0.elapsed_time(e1)) / (reps * S)
        coll_us = 0.0
        if world > 1:                                          # the same all-reduce alone, for the collective's share
            flat = torch.zeros(g.grad_bytes // 4, device=dev)
            coll_us = D.timed(lambda k: dist.all_reduce(flat), 20) * 1e3
        in_sync = None
        if world > 1:                                          # the all-reduced gradients must leave every rank with the same weights
            probe = torch.cat([p.detach().reshape(-1)[:64] for p in model.parameters()])
            lo, hi = probe.clone(), probe.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            in_sync = bool(torch.equal(lo, hi))
        entry = {"metric": "ae_train_clouds_per_s", "value": bsz * world / (ms * 1e-3), "unit": "clouds/s", "n_gpus": world,
                 "params_in_sync_across_ranks": in_sync, "overlapped_allreduce_bytes": g.early_bytes if world > 1 else 0,
                 "scaling": "weak", "ms_per_step": ms, "batch_per_gpu": bsz, "steps_per_graph": S,
                 "last_loss": float(g.losses[-1].item()),
                 "collective": {"kind": "none (1 GPU)" if world == 1 else "NCCL all-reduce of the fp32 gradients every step, captured in the "
                                "graph: the decoder's bucket on a side stream under the encoder's backward, the encoder's after it",
                                "bytes": g.grad_bytes if world > 1 else 0, "us_alone": coll_us},
                 "config": {"workload": "AE train step (BASELINE configs[0] at the reference's dims): encoder trunk fwd+bwd with "
                                        "BatchNorm batch statistics and Chamfer fwd+bwd on this library's kernels; global MLP, decoder "
                                        "MLP and Adam on stock torch; incomplete N=1400 -> reconstruction 2048 vs complete 2048"}}
        del g, model, opt
        torch.cuda.empty_cache()
        if rank == 0:
            try:
                torch.manual_seed(0)
                stock = StockAutoencoder().to(dev).train()
                sopt = torch.optim.Adam(stock.parameters(), lr=1e-3, weight_decay=1e-5, fused=True)

                def stock_step(k):
                    x, y = batches[k % S]
                    sopt.zero_grad()
                    dm = torch.cdist(stock(x), y, p=2)
                    loss = torch.mean((torch.mean(torch.min(dm, dim=2)[0], dim=1) + torch.mean(torch.min(dm, dim=1)[0], dim=1)) / 2.0)
                    loss.backward()
                    sopt.step()
                for k in range(3):
                    stock_step(k)
                torch.cuda.synchronize()
                e0.record()
                for k in range(10):
                    stock_step(k)
                e1.record()
                torch.cuda.synchronize()
                sms = e0.elapsed_time(e1) / 10
                entry["torch_cuda"] = {"value": bsz / (sms * 1e-3), "unit": "clouds/s", "ms_per_step": sms,
                                       "what": "the same step on stock torch CUDA kernels (Conv1d/BatchNorm1d/ReLU encoder, cdist->min->mean "
                                               "Chamfer, autograd, Adam), eager, one GPU"}
                del stock, sopt
            except Exception as e:
                entry["torch_cuda"] = {"error": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.empty_cache()
        D.barrier()
        out[f"ae_step_b{bsz}"] = entry
    return out


class StockAutoencoder(torch.nn.Module):
    """The reference's autoencoder restated with stock torch layers (models/autoencoder.py:13-171), for the torch_cuda arm."""

    def __init__(self, dims=(64, 128, 128, 256, 128), latent=128, dec=(256, 256, 6144)):
        super().__init__()
        nn = torch.nn
        seq, c_in = [], 3
        for c in dims:
            seq += [nn.Conv1d(c_in, c, 1), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
            c_in = c
        self.point_mlp = nn.Sequential(*seq)
        self.global_mlp = nn.Sequential(nn.Linear(c_in, latent), nn.BatchNorm1d(latent), nn.ReLU(inplace=True))
        seq, c_in = [], latent
        for c in dec[:-1]:
            seq += [nn.Linear(c_in, c), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
            c_in = c
        seq.append(nn.Linear(c_in, dec[-1]))
        self.mlp = nn.Sequential(*seq)

    def forward(self, x):
        g = self.global_mlp(torch.max(self.point_mlp(x.transpose(2, 1)), dim=2)[0])
        return self.mlp(g).view(x.shape[0], -1, 3)


def torch_cuda_measurement(dev, ring):
    """The kernels to beat: the reference's own op sequence on stock torch CUDA kernels on this GPU.
    Chamfer: torch.cdist -> min x2 -> mean, autograd backward (utils/losses.py:29-37,54-59,75), fp32 (TF32 off, as torch
    defaults).  Encoder: the stock Conv1d/BatchNorm1d/ReLU stack + max + Linear head in eval mode
    (models/autoencoder.py:13-76), fp32 and under bf16 autocast.  Plain torch calls written out here; rank 0 only."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def chamfer_step(a, b):
        a = a.detach().requires_grad_(True)
        dm = torch.cdist(a, b, p=2)
        loss = torch.mean((torch.mean(torch.min(dm, dim=2)[0], dim=1) + torch.mean(torch.min(dm, dim=1)[0], dim=1)) / 2.0)
        loss.backward()
        return loss

    out = {}
    try:
        for k in range(3):
            chamfer_step(*ring[k])
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for k in range(reps):
            chamfer_step(*ring[k % len(ring)])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["chamfer_fwd_bwd"] = {"value": B / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
                                  "what": "torch.cdist -> min x2 -> mean + autograd backward, B=32 N=M=2048, fp32, eager"}
    except Exception as e:
        out["chamfer_fwd_bwd"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()

    def stock_encoder(dims):
        seq, c_in = [], 3
        for c in dims:
            seq += [torch.nn.Conv1d(c_in, c, 1), torch.nn.BatchNorm1d(c), torch.nn.ReLU(inplace=True)]
            c_in = c
        trunk = torch.nn.Sequential(*seq)
        head = torch.nn.Sequential(torch.nn.Linear(c_in, 128), torch.nn.BatchNorm1d(128), torch.nn.ReLU(inplace=True))
        return trunk.eval().to(dev), head.eval().to(dev)

    gen = torch.Generator(device="cpu").manual_seed(77)
    xs = [sphere(gen, ENC_B, ENC_N).to(dev) for _ in range(4)]
    for name, dims in (("encoder", ENC_DIMS), ("encoder_config_dims", CFG_DIMS)):
        torch.manual_seed(0)
        trunk, head = stock_encoder(dims)
        for mode in ("fp32", "bf16_autocast"):
            try:
                def fwd(x):
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                        return head(torch.max(trunk(x.transpose(2, 1)), dim=2)[0])
                for k in range(2):
                    fwd(xs[k])
                torch.cuda.synchronize()
                reps = 8
                e0.record()
                for k in range(reps):
                    fwd(xs[k % len(xs)])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                out[f"{name}_{mode}"] = {"value": ENC_B / (ms * 1e-3), "unit": "clouds/s", "ms_per_step": ms,
                                         "what": f"stock Conv1d/BatchNorm1d/ReLU x{len(dims)} + max + Linear head, eval, "
                                                 f"B=256 N=2048, dims {dims}, {mode}"}
            except Exception as e:
                out[f"{name}_{mode}"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        del trunk, head
        torch.cuda.empty_cache()
    return {"torch_cuda": out}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
