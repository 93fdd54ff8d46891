"""Tensor-core Chamfer filter (RLG_CHAMFER_ALGO_TENSOR): parity against the simple kernel at several shapes, the filter's
actual error against float64, and the sweep kernel's time.  python tools/try_tcfilter.py"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402

_lib = importlib.import_module("gan-rl_3d_b200._lib")
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def sphere(b, n, scale=1.0, shift=0.0):
    x = torch.randn(b, n, 3, generator=g)
    return ((x / x.norm(dim=2, keepdim=True)) * scale + shift).to(dev)


def check(B, N, M, scale=1.0, shift=0.0, uniform=False):
    if uniform:
        a = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev) * scale + shift
        b = (torch.rand(B, M, 3, generator=g) * 2 - 1).to(dev) * scale + shift
    else:
        a, b = sphere(B, N, scale, shift), sphere(B, M, scale, shift)
    ref = rlg.chamfer_nearest(a, b, simple=True)
    out = rlg.chamfer_nearest(a, b, tensor=True)
    out2 = rlg.chamfer_nearest(a, b, tensor=True)          # second call: workspace hygiene
    torch.cuda.synchronize()
    ok = all(torch.equal(x, y) for x, y in zip(out[:4], ref[:4]))
    ok2 = all(torch.equal(x, y) for x, y in zip(out2[:4], ref[:4]))
    nbad = [(x != y).sum().item() for x, y in zip(out[:4], ref[:4])]
    print(f"B={B} N={N} M={M} scale={scale} shift={shift} uniform={uniform}: match={ok} again={ok2} mismatches={nbad}", flush=True)
    return ok and ok2


def filter_error(B, N, M, scale=1.0, shift=0.0):
    """Runs only the sweep (TILE_ONLY), reads the raw keys back and compares the filter's best value per query with the
    float64 minimum of |x - y|^2, in units of u (a^2 + b^2)."""
    a, b = sphere(B, N, scale, shift), sphere(B, M, scale, shift)
    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev).fill_(0xFF)
    stream = torch.cuda.current_stream().cuda_stream
    flags = _lib.CHAMFER_WS_CLEAN | _lib.CHAMFER_TILE_ONLY | _lib.CHAMFER_ALGO_TENSOR
    rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                             i2.data_ptr(), None, None, ws.data_ptr(), ws.numel(), flags, stream)
    assert rc == 0, lib.rlg_last_error()
    torch.cuda.synchronize()
    keys = ws[: 8 * B * (N + M)].view(torch.int64)
    rowkey = keys[: B * N].view(B, N)
    colkey = keys[B * N:].view(B, M)
    a64, b64 = a.double(), b.double()
    D = ((a64[:, :, None, :] - b64[:, None, :, :]) ** 2).sum(-1)
    u = 2.0 ** -24
    na, nb = (a64 ** 2).sum(-1), (b64 ** 2).sum(-1)
    for name, key, truth, nq, nmax in (("rows", rowkey, D.min(2).values, na, nb.max(1, keepdim=True).values),
                                       ("cols", colkey, D.min(1).values, nb, na.max(1, keepdim=True).values)):
        val = (key >> 32).to(torch.int32).view(torch.float32).double()
        grp = (key & 0xffffffff)
        err = (val - truth).abs() / (u * (nq + nmax))
        truth_grp = (D.argmin(2) if name == "rows" else D.argmin(1)) // 32
        print(f"  filter error {name}: max {err.max().item():.2f} u(a^2+b^2), mean {err.mean().item():.3f}; "
              f"best group == true group for {(grp == truth_grp).double().mean().item() * 100:.3f} %", flush=True)


def time_sweep(B, N, M, tensor):
    slots = max(2, min(192, (288 << 20) // ((N + M) * B * 12)))
    ring = [(sphere(B, N), sphere(B, M)) for _ in range(slots)]
    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=dev).fill_(0xFF)
    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tile_only in (True, False):
        flags = _lib.CHAMFER_WS_CLEAN | (_lib.CHAMFER_TILE_ONLY if tile_only else 0) | (_lib.CHAMFER_ALGO_TENSOR if tensor else 0)

        def run(k):
            a, b = ring[k % slots]
            rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                                     i2.data_ptr(), None, None, ws.data_ptr(), ws.numel(), flags, stream)
            assert rc == 0, lib.rlg_last_error()

        ws.fill_(0xFF)
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
        print(f"B={B} N={N} M={M} {'tensor' if tensor else 'fp32  '} {'sweep only      ' if tile_only else 'sweep + finalize'}: "
              f"{us:8.2f} us  {tf:6.2f} algorithmic TFLOP/s ({tf / 74.45 * 100:5.1f}% of FP32 FFMA peak)", flush=True)


if __name__ == "__main__":
    ok = True
    ok &= check(2, 256, 256)
    filter_error(2, 256, 256)
    ok &= check(4, 2048, 2048)
    filter_error(4, 2048, 2048)
    filter_error(2, 2048, 2048, scale=0.01, shift=5.0)
    ok &= check(3, 1400, 2048)
    ok &= check(2, 100, 37)
    ok &= check(2, 1, 300)
    ok &= check(2, 2048, 2048, scale=0.05, shift=3.0)
    ok &= check(2, 2048, 2048, uniform=True)
    ok &= check(1, 5000, 3000, uniform=True)
    print("ALL MATCH" if ok else "MISMATCH", flush=True)
    for tensor in (False, True):
        time_sweep(32, 2048, 2048, tensor)
    time_sweep(16, 16384, 16384, True)
