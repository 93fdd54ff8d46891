"""Per-kernel timing of one Chamfer training step (forward tile, finalize, backward) with CUDA events around
each C-ABI call, at the headline shape.  Diagnostic only.
    python tools/step_breakdown.py [B] [N] [M]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

_lib = importlib.import_module("gan-rl_3d_b200._lib")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
M = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
lib = _lib.load()
dev = torch.device("cuda:0")
ring = [(O.make_clouds(B, N, "sphere", 10 + k).to(dev), O.make_clouds(B, M, "sphere", 90 + k).to(dev)) for k in range(64)]
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
m1 = torch.empty(B, device=dev); m2 = torch.empty(B, device=dev)
ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev)
ws.fill_(0xFF)
g = torch.full((B,), 0.5 / B, device=dev)
ga = torch.empty(B, N, 3, device=dev); gb = torch.empty(B, M, 3, device=dev)
st = torch.cuda.current_stream().cuda_stream


def fwd(a, b, flags):
    rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                             i2.data_ptr(), m1.data_ptr(), m2.data_ptr(), ws.data_ptr(), ws.numel(), flags, st)
    _lib.check("fwd", rc)


def bwd(a, b):
    rc = lib.rlg_chamfer_bwd(a.data_ptr(), b.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                             g.data_ptr(), g.data_ptr(), B, N, M, ga.data_ptr(), gb.data_ptr(), 0, st)
    _lib.check("bwd", rc)


def timed(fn, reps=200):
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    global st
    with torch.cuda.stream(s):
        st = s.cuda_stream
        for k in range(3):
            fn(*ring[k])
        s.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for k in range(len(ring)):
                fn(*ring[k])
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(1, reps // len(ring))
    e0.record()
    for _ in range(n):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * len(ring))


flop = 8.0 * N * M * B
for name, algo in (("fp32 sweep  ", 0), ("tensor sweep", _lib.CHAMFER_ALGO_TENSOR)):
    ws.fill_(0xFF)
    torch.cuda.synchronize()
    t_tile = timed(lambda a, b: fwd(a, b, _lib.CHAMFER_WS_CLEAN | _lib.CHAMFER_TILE_ONLY | algo))
    ws.fill_(0xFF)
    torch.cuda.synchronize()
    t_fwd = timed(lambda a, b: fwd(a, b, _lib.CHAMFER_WS_CLEAN | algo))
    t_fwd_memset = timed(lambda a, b: fwd(a, b, algo))
    t_bwd = timed(bwd)
    t_all = timed(lambda a, b: (fwd(a, b, _lib.CHAMFER_WS_CLEAN | algo), bwd(a, b)))
    print(f"B={B} N={N} M={M} {name} (graph replay, per call): sweep {t_tile:.2f} us ({flop / t_tile / 1e6:.2f} TFLOP/s)  "
          f"fwd(sweep+finalize) {t_fwd:.2f} us  fwd+memset {t_fwd_memset:.2f} us  bwd {t_bwd:.2f} us  fwd+bwd {t_all:.2f} us")
