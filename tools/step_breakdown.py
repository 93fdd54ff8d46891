"""Per-stage timing of one Chamfer training step with CUDA events around CUDA-graph replays of each C-ABI call,
at the headline shape.  Diagnostic only (not a benchmark value).
    python tools/step_breakdown.py [B] [N] [M] [kind]
    RLG_EXPERIMENTS_LIB=1 python tools/step_breakdown.py ...      also times the first-generation tensor sweep (A/B)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402,F401
from oracle import oracle as O  # noqa: E402

_lib = importlib.import_module("gan-rl_3d_b200._lib")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
M = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
KIND = sys.argv[4] if len(sys.argv) > 4 else "sphere"
EXP = os.environ.get("RLG_EXPERIMENTS_LIB") == "1"
ONLY = sys.argv[5] if len(sys.argv) > 5 else ""          # "tensor": only the tensor-sweep rows (A-B variant runs)
lib = _lib.load()
dev = torch.device("cuda:0")
slot_bytes = (B * N + B * M) * 12
n_ring = max(4, min(64, (300 << 20) // slot_bytes))          # > L2
ring = [(O.make_clouds(B, N, KIND, 10 + k).to(dev), O.make_clouds(B, M, KIND, 90 + k).to(dev)) for k in range(n_ring)]
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
m1 = torch.empty(B, device=dev); m2 = torch.empty(B, device=dev)
loss = torch.empty((), device=dev)
ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev)
ws.fill_(0xFF)
gone = torch.ones((), device=dev)
ga = torch.empty(B, N, 3, device=dev); gb = torch.empty(B, M, 3, device=dev)
st = torch.cuda.current_stream().cuda_stream


def fwd(a, b, flags, zero=False):
    rc = lib.rlg_chamfer_loss_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                                  i2.data_ptr(), m1.data_ptr(), m2.data_ptr(), loss.data_ptr(), 0.5 / B, 0.5 / B,
                                  ga.data_ptr() if zero else None, gb.data_ptr() if zero else None,
                                  ws.data_ptr(), ws.numel(), flags, st)
    _lib.check("fwd", rc)


def bwd(a, b, flags=0):
    rc = lib.rlg_chamfer_loss_bwd(a.data_ptr(), b.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                  gone.data_ptr(), 0.5 / B, 0.5 / B, B, N, M, ga.data_ptr(), gb.data_ptr(), flags, st)
    _lib.check("bwd", rc)


def timed(fn, reps=400, dirty=False):
    """us per call: every ring slot once per CUDA-graph replay, replayed back to back."""
    global st
    ws.fill_(0xFF)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        st = s.cuda_stream
        for k in range(3):
            fn(*ring[k])
            if dirty:
                ws.fill_(0xFF)
        s.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for k in range(len(ring)):
                fn(*ring[k])
                if dirty:
                    ws.fill_(0xFF)          # the measured call leaves the workspace dirty: restore it (memset node)
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


L = _lib
flop = 8.0 * N * M * B
t_fill = timed(lambda a, b: None, dirty=True) if True else 0.0
print(f"B={B} N={N} M={M} {KIND}; ring {n_ring} slots; workspace memset alone {t_fill:.2f} us (subtracted from the 'dirty' rows)")
rows = [("fp32: sweep", L.CHAMFER_WS_CLEAN | L.CHAMFER_TILE_ONLY, True),
        ("fp32: forward (sweep + refinement kernel)", L.CHAMFER_WS_CLEAN, False),
        ("tensor: filter sweep only (diagnostic)", L.CHAMFER_WS_CLEAN | L.CHAMFER_ALGO_TENSOR | L.CHAMFER_FILTER_ONLY, True),
        ("tensor: forward (one fused launch)", L.CHAMFER_WS_CLEAN | L.CHAMFER_ALGO_TENSOR, False),
        ("tensor+track_two: forward", L.CHAMFER_WS_CLEAN | L.CHAMFER_ALGO_TENSOR | L.CHAMFER_TRACK_TWO, False)]
if EXP:
    rows += [("tensor v1 (round 1): sweep", L.CHAMFER_WS_CLEAN | L.X_CHAMFER_TENSOR_V1 | L.CHAMFER_TILE_ONLY, True),
             ("tensor v1 (round 1): forward (sweep + finalize)", L.CHAMFER_WS_CLEAN | L.X_CHAMFER_TENSOR_V1, False)]
if ONLY:
    rows = [r for r in rows if r[0].startswith(ONLY)]
for name, flags, dirty in rows:
    t = timed(lambda a, b: fwd(a, b, flags), dirty=dirty) - (t_fill if dirty else 0.0)
    print(f"{name:52s} {t:8.2f} us   {flop / t / 1e6:7.2f} TFLOP/s algorithmic")
if ONLY:
    sys.exit(0)
t_bwd = timed(lambda a, b: bwd(a, b))
print(f"{'backward (memsets + kernel)':52s} {t_bwd:8.2f} us")
for name, algo in (("fp32", 0), ("tensor", L.CHAMFER_ALGO_TENSOR)):
    t = timed(lambda a, b: (fwd(a, b, L.CHAMFER_WS_CLEAN | algo, zero=True), bwd(a, b, L.CHAMFER_BWD_ACCUMULATE)))
    print(f"{name + ': forward + backward':52s} {t:8.2f} us   {B / t * 1e6:10.0f} pairs/s")
