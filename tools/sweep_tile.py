"""Times the Chamfer tile kernel variants alone (flags bits 8..11) at a given shape, CUDA events, ring of
inputs larger than L2.  python tools/sweep_tile.py [B N M]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402

_lib = importlib.import_module("gan-rl_3d_b200._lib")
lib = _lib.load()
B, N, M = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 2048, 2048)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def sphere(b, n):
    x = torch.randn(b, n, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True)).to(dev)


slots = max(2, min(192, (288 << 20) // ((N + M) * B * 12)))
ring = [(sphere(B, N), sphere(B, M)) for _ in range(slots)]
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev).fill_(0xFF)
stream = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
names = {(0, 0): "filter auto", (0, 1): "filter R16 occ2 pp", (0, 2): "filter R16 occ2 nopp", (0, 3): "filter R16 occ3 pp",
         (0, 4): "filter R16 occ3 nopp", (0, 7): "filter R8 occ3 pp", (0, 8): "filter R8 occ4 nopp", (0, 9): "filter R4 occ4",
         (1, 0): "direct R8 occ3 min2", (0, 10): "KO col scan", (0, 11): "KO row bookkeeping", (0, 12): "KO scan+rows",
         (0, 13): "integer min3 on t", (0, 14): "KO scan+rows+staging", (0, 15): "KO scan+rows+staging, int min"}
if os.environ.get("SWEEP_VARIANTS"):
    keep = {int(v) for v in os.environ["SWEEP_VARIANTS"].split(",")}
    names = {k: v for k, v in names.items() if k[0] == 0 and k[1] in keep}
for (direct, var), name in names.items():
    algo = _lib.CHAMFER_ALGO_DIRECT if direct else 0
    flags = _lib.CHAMFER_WS_CLEAN | _lib.CHAMFER_TILE_ONLY | (var << 8) | algo

    def run(k):
        a, b = ring[k % slots]
        rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                                 i2.data_ptr(), None, None, ws.data_ptr(), ws.numel(), flags, stream)
        assert rc == 0, lib.rlg_last_error()

    for k in range(10):
        run(k)
    torch.cuda.synchronize()
    reps = 200
    e0.record()
    for k in range(reps):
        run(k)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    tf = 8.0 * N * M * B / (us * 1e-6) / 1e12
    print(f"variant {direct}/{var} ({name}): {us:8.2f} us  {tf:6.2f} TFLOP/s  {tf / 74.45 * 100:5.1f}% of FFMA peak")
    if var in (10, 11, 12, 14, 15):
        continue                      # knock-out timing experiments: results are wrong by construction
    # correctness of the variant against the production path
    full = _lib.CHAMFER_WS_CLEAN | (var << 8) | algo
    a, b = ring[0]
    m1 = torch.empty(B, device=dev); m2 = torch.empty(B, device=dev)
    ws.fill_(0xFF)
    rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                             i2.data_ptr(), m1.data_ptr(), m2.data_ptr(), ws.data_ptr(), ws.numel(), full, stream)
    assert rc == 0
    ref = rlg.chamfer_nearest(a, b, simple=True)
    ok = all(torch.equal(x, y) for x, y in zip((d1, d2, i1, i2), ref[:4]))
    print("   matches simple kernel:", ok)
