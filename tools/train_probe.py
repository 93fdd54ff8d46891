"""Train-mode encoder trunk: error of the B200 kernels and of stock fp32 CUDA against float64, and timings."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(dims, B, N, gscale=1.0):
    torch.manual_seed(0)
    enc = O.RefEncoderPort(3, 128, dims)
    O.randomize_bn(enc, 3)
    enc.train()
    x = O.make_clouds(B, N, "sphere", 1)
    g = gscale * torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(2))
    e64 = copy.deepcopy(enc).double()
    f64 = torch.max(e64.point_mlp(x.double().transpose(2, 1)), dim=2)[0]
    f64.backward(g.double())
    res = {}
    for name in ("ours", "stock"):
        e = copy.deepcopy(enc).to(dev)
        if name == "ours":
            f = rlg.trunk_pool_autograd(e, x.to(dev))
        else:
            f = torch.max(e.point_mlp(x.to(dev).transpose(2, 1)), dim=2)[0]
        f.backward(g.to(dev))
        torch.cuda.synchronize()
        errs = {"pooled": rel(f.detach(), f64.detach())}
        for (n, p), (_, q) in zip(e.point_mlp.named_parameters(), e64.point_mlp.named_parameters()):
            if n.endswith("bias") and int(n.split(".")[0]) % 3 == 0:
                continue
            errs[n] = rel(p.grad, q.grad)
        for (n, b), (_, c) in zip(e.point_mlp.named_buffers(), e64.point_mlp.named_buffers()):
            if not n.endswith("tracked"):
                errs[n] = max(errs.get("buffers", 0.0), rel(b, c))
        res[name] = errs
    print(f"dims {dims} B={B} N={N} gscale={gscale}")
    for n in res["ours"]:
        print(f"   {n:28s} ours {res['ours'][n]:.2e}   stock fp32 {res['stock'][n]:.2e}")

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    e = copy.deepcopy(enc).to(dev)
    xd, gd = x.to(dev), g.to(dev)

    def ours_fb():
        for p in e.parameters():
            p.grad = None
        rlg.trunk_pool_autograd(e, xd).backward(gd)

    def stock_fb():
        for p in e.parameters():
            p.grad = None
        torch.max(e.point_mlp(xd.transpose(2, 1)), dim=2)[0].backward(gd)

    def ours_f():
        with torch.no_grad():
            rlg.trunk_pool_autograd(e, xd)

    def stock_f():
        with torch.no_grad():
            torch.max(e.point_mlp(xd.transpose(2, 1)), dim=2)[0]

    print(f"   fwd: ours {ours_f.__name__ and timed(ours_f):.1f} us, stock {timed(stock_f):.1f} us;  fwd+bwd: ours {timed(ours_fb):.1f} us, stock {timed(stock_fb):.1f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vary":
        run([64, 128, 128, 256, 128], 32, 1400, 1e-6)
        run([64, 128, 128, 256, 128], 16, 1400, 1.0)
        run([64, 128, 128, 256, 128], 32, 700, 1.0)
        run([64, 128, 128], 32, 1400, 1.0)
        sys.exit(0)
    run([64, 128, 128, 256, 128], 2, 200)
    run([64, 128, 128, 256, 128], 16, 1400, 1e-6)
    run([64, 128, 128, 256, 128], 32, 1400)
    run([64, 64], 8, 2048)
