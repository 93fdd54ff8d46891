"""Chamfer backward alone: float-atomics kernel vs the reproducible fixed-point variant, each as a CUDA graph of one call per
ring slot (ring larger than L2), CUDA events around replays.  Diagnostic only (not a benchmark value).
    python tools/bwd_probe.py [B N M]                           default: cfg2 (32 x 2048^2) and cfg5-shaped (64 x 16384^2) clouds
    RLG_EXPERIMENTS_LIB=<variant .so> python tools/bwd_probe.py  the same on an A/B build (e.g. -DRLG_BWD_SCALAR_RED)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

_lib = importlib.import_module("gan-rl_3d_b200._lib")
lib = _lib.load()
dev = torch.device("cuda:0")
print("library:", rlg.library_path())

SHAPES = [tuple(int(v) for v in sys.argv[1:4])] if len(sys.argv) >= 4 else [(32, 2048, 2048), (64, 16384, 16384)]
for B, N, M in SHAPES:
    slot = B * (N + M) * 56
    n_ring = max(2, min(48, (400 << 20) // slot))
    ring = []
    for k in range(n_ring):
        a, b = O.make_clouds(B, N, "sphere", 10 + k).to(dev), O.make_clouds(B, M, "sphere", 90 + k).to(dev)
        d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
        ring.append((a, b, d1, d2, i1, i2, torch.empty_like(a), torch.empty_like(b)))
    g = torch.full((B,), 0.5 / B, device=dev)
    ws = torch.empty(lib.rlg_chamfer_bwd_ws_bytes(B, N, M), dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    for name in ("atomics", "deterministic"):
        def call(slot_):
            a, b, d1, d2, i1, i2, ga, gb = slot_
            args = (a.data_ptr(), b.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(), g.data_ptr(),
                    g.data_ptr(), B, N, M, ga.data_ptr(), gb.data_ptr())
            st = torch.cuda.current_stream().cuda_stream
            rc = (lib.rlg_chamfer_bwd(*args, 0, st) if name == "atomics"
                  else lib.rlg_chamfer_bwd_det(*args, ws.data_ptr(), ws.numel(), 0, st))
            _lib.check("bwd", rc)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            call(ring[0])
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                for slot_ in ring:
                    call(slot_)
        torch.cuda.synchronize()
        for _ in range(3):
            graph.replay()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * n_ring)
        print(f"B={B} N={N} M={M} {name:13s}: {us:8.2f} us per call (memset node(s) included) = "
              f"{slot / us / 1e3:7.1f} GB/s algorithmic (56 B per point), ring of {n_ring}")
    del ring
